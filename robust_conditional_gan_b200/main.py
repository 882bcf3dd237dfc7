"""Drop-in for mnist/main.py: the same flags (mnist/main.py:12-67), the same construction of DCGAN (:106-131) and the same
train-or-load, then recover_labels control flow (:135-142), on this package's CUDA path.

    python -m robust_conditional_gan_b200.main --algorithm rcgan --alpha 0.3 --disc_type projection --noestimate_confuse \\
        --spectral_norm --max_norm --checkpoint_dir rcgan --epoch 100           # = mnist/run_rcgan.sh

`--model rcgan|rcganu|rcgany|biased|unbiased|ambient` (north_star shorthand) expands to the flag set of the matching run
script.  Extra flags of this implementation: --precision, --max_iters, --synthetic / --synthetic_size (no dataset ships here)."""
import os
import sys
from datetime import datetime

import numpy as np

from . import flags as flags_lib

flags = flags_lib.Flags()
flags.DEFINE_integer("epoch", 5, "Epoch to train [25]")
flags.DEFINE_float("learning_rate", 0.0002, "Learning rate of for adam [0.0002]")
flags.DEFINE_float("beta1", 0.5, "Momentum term of adam [0.5]")
flags.DEFINE_float("train_size", np.inf, "The size of train images [np.inf]")
flags.DEFINE_integer("batch_size", 100, "The size of batch images")
flags.DEFINE_integer("input_height", 108, "The size of image to use (will be center cropped). [108]")
flags.DEFINE_integer("input_width", None, "The size of image to use (will be center cropped). If None, same value as input_height [None]")
flags.DEFINE_integer("output_height", 64, "The size of the output images to produce [64]")
flags.DEFINE_integer("output_width", None, "The size of the output images to produce. If None, same value as output_height [None]")
flags.DEFINE_string("dataset", "mnist", "The name of dataset [mnist]")
flags.DEFINE_string("checkpoint_dir", "rcgan", "Directory name to save the checkpoints [checkpoint]")
flags.DEFINE_string("checkpoint", None, "Directory name to save the checkpoints [checkpoint]")
flags.DEFINE_string("sample_dir", "samples/", "Directory name to save the image samples")
flags.DEFINE_string("data_dir", "../data/", "Root directory of dataset [data]")
flags.DEFINE_string('dir_prefix', None, "dir name prefix")
flags.DEFINE_string('logs_dir', './logs', "logs directory")
flags.DEFINE_boolean('logs_at_ckpt', False, "set logs dir to chechkpoint dir")
flags.DEFINE_string('script_file', None, "script file name for storing script along with results")
flags.DEFINE_boolean("train", False, "True for training, False for testing [False]")
flags.DEFINE_boolean("crop", False, "True for training, False for testing [False]")
flags.DEFINE_boolean("visualize", False, "True for visualizing, False for nothing [False]")
flags.DEFINE_integer("z_dim", 100, "Dimension of input noise Z to the generator")
flags.DEFINE_string("algorithm", "biased", "[biased, unbiased, rcgan, ambient]")
flags.DEFINE_boolean("estimate_confuse", True, "whether to estimate confusion matrix")
flags.DEFINE_float("confuse_multiplier", 10.0, "learning rate multiplier for confusion matrix")
flags.DEFINE_boolean("perm_regularizer", True, "whether to use auxillary permutation regularizer classifier")
flags.DEFINE_float("perm_multiplier", 10.0, "learning rate multiplier for permutation regularizer")
flags.DEFINE_float("alpha", 1.0, "noise in labels")
flags.DEFINE_boolean("confusion_class_depend", False,
                     "whether to generate rows of confusion matrix in a class dependent way or in one coin model way")
flags.DEFINE_string("disc_type", "vanilla", "type of discriminator to use [vanilla, projection]")
flags.DEFINE_string('loss_fn', 'hinge', 'GAN loss function')
flags.DEFINE_boolean("real_match", False, 'whether to match y_gen with y_real in for each batch')
flags.DEFINE_boolean('add_noise', False, 'whether to add noise to both real and fake labels y_real, y_fake')
flags.DEFINE_float("noise_alpha", 0.3, "effective noise in labels")
flags.DEFINE_integer("noise_start", 30, "noise schedule start")
flags.DEFINE_integer("noise_end", 80, "noise schedule end")
flags.DEFINE_boolean('concat_y', False, 'whether to concat y to projection discriminator')
flags.DEFINE_list('concat_y_layers', ['1', ], 'layers of projection discriminator where we want to concat y [1, 2, 3, 4]')
flags.DEFINE_boolean('spectral_norm', True, 'whether to use spectral normalization on conv2d layers of the discriminator')
flags.DEFINE_boolean('max_norm', True, 'whether to use maximum value (clip) normalization on linear layers of discriminator')
flags.DEFINE_integer("recover_epoch", 1000, "Epoch to train [25]")
flags.DEFINE_integer("recover_batch_size", 500, "The size of batch images [64]")
flags.DEFINE_float("recover_learning_rate", 5.e+2, "Learning rate of for adam [0.0002]")
# the run scripts pass --noaux_classifier: accepted (the reference's own flag list does not define it and would reject it)
flags.DEFINE_boolean("aux_classifier", False, "accepted for the run scripts; unused")
# ---- this implementation
flags.DEFINE_string("model", None, "shorthand for a run script's flag set: rcgan | rcganu | rcgany | biased | unbiased | ambient")
flags.DEFINE_string("precision", "bf16", "bf16 (tensor-core path) | fp32 (parity mode)")
flags.DEFINE_integer("max_iters", None, "stop training after this many iterations")
flags.DEFINE_boolean("synthetic", True, "use synthetic 28x28x1 data when the MNIST idx files are absent")
flags.DEFINE_integer("synthetic_size", 70000, "number of synthetic samples")
flags.DEFINE_boolean("recover", True, "run recover_labels after training (mnist/main.py:142)")
FLAGS = flags.FLAGS

# mnist/run_*.sh
MODEL_FLAGS = {
    'rcgan': dict(algorithm='rcgan', disc_type='projection', estimate_confuse=False, add_noise=False, concat_y=False, spectral_norm=True,
                  max_norm=True),
    'rcganu': dict(algorithm='rcgan', disc_type='projection', estimate_confuse=True, add_noise=False, concat_y=False, spectral_norm=True,
                   max_norm=True),
    'rcgany': dict(algorithm='rcgan', disc_type='projection', estimate_confuse=False, add_noise=True, noise_alpha=0.3, noise_start=30,
                   noise_end=80, concat_y=True, concat_y_layers=['1'], spectral_norm=True, max_norm=True),
    'biased': dict(algorithm='biased', disc_type='vanilla', loss_fn='ce', real_match=True, estimate_confuse=False, spectral_norm=False,
                   max_norm=False),
    'ambient': dict(algorithm='ambient', disc_type='vanilla', loss_fn='ce', real_match=True, estimate_confuse=False, spectral_norm=False,
                    max_norm=False),
    'unbiased': dict(algorithm='unbiased', disc_type='projection', estimate_confuse=False, spectral_norm=True, max_norm=True),
}


def configure(FLAGS):
    """mnist/main.py:70-100 (everything before the session) -> FLAGS ready for DCGAN(config=FLAGS)."""
    if FLAGS.model is not None:
        if FLAGS.model not in MODEL_FLAGS:
            raise ValueError('--model must be one of %s' % sorted(MODEL_FLAGS))
        for k, v in MODEL_FLAGS[FLAGS.model].items():
            setattr(FLAGS, k, v)
    FLAGS.concat_y_layers = [int(x) for x in FLAGS.concat_y_layers]
    FLAGS.dir_prefix = '' if FLAGS.dir_prefix is None else FLAGS.dir_prefix + '_'
    if FLAGS.checkpoint is None:
        FLAGS.checkpoint_dir = os.path.join(
            FLAGS.checkpoint_dir, FLAGS.dir_prefix + FLAGS.algorithm + "_" + str(FLAGS.alpha) + "_" + FLAGS.disc_type + "_" +
            datetime.now().strftime("%Y%m%d-%H%M%S"))
    else:
        FLAGS.checkpoint_dir = os.path.join(FLAGS.checkpoint_dir, FLAGS.checkpoint)
    FLAGS.sample_dir = os.path.join(FLAGS.checkpoint_dir, 'samples/')
    FLAGS.input_height = FLAGS.output_height = 28
    if FLAGS.input_width is None:
        FLAGS.input_width = 28
    if FLAGS.output_width is None:
        FLAGS.output_width = 28
    if FLAGS.logs_at_ckpt:
        FLAGS.logs_dir = FLAGS.checkpoint_dir
    FLAGS.dataset = 'mnist'
    return FLAGS


def main(_):
    from . import data
    from .model import DCGAN
    configure(FLAGS)
    print(FLAGS.flag_values_dict())
    os.makedirs(FLAGS.checkpoint_dir, exist_ok=True)
    os.makedirs(FLAGS.sample_dir, exist_ok=True)
    with open(os.path.join(FLAGS.checkpoint_dir, 'command.txt'), 'w') as f:      # utils.dump_script's record of the invocation
        f.write(' '.join(sys.argv) + '\n')
    X, y = data.mnist_data(FLAGS.data_dir, FLAGS.dataset, allow_synthetic=FLAGS.synthetic, n_synthetic=FLAGS.synthetic_size)
    dcgan = DCGAN(
        None, input_width=FLAGS.input_width, input_height=FLAGS.input_height, output_width=FLAGS.output_width,
        output_height=FLAGS.output_height, batch_size=FLAGS.batch_size, sample_num=FLAGS.batch_size, y_dim=10, z_dim=FLAGS.z_dim,
        dataset_name=FLAGS.dataset, crop=FLAGS.crop, checkpoint_dir=FLAGS.checkpoint_dir, data_dir=FLAGS.data_dir,
        algorithm=FLAGS.algorithm, estimate_confuse=FLAGS.estimate_confuse, perm_regularizer=FLAGS.perm_regularizer,
        alpha=FLAGS.alpha, disc_type=FLAGS.disc_type, add_noise=FLAGS.add_noise, noise_alpha=FLAGS.noise_alpha, config=FLAGS,
        precision=FLAGS.precision, data=(X, y))
    if FLAGS.train:
        dcgan.train(FLAGS, max_iters=FLAGS.max_iters)
    else:
        if not dcgan.load(FLAGS.checkpoint_dir)[0]:
            print("[!] Training a model first, then run test mode")
            dcgan.train(FLAGS, max_iters=FLAGS.max_iters)
    if FLAGS.recover:
        dcgan.recover_labels(FLAGS)
    return dcgan


if __name__ == '__main__':
    flags_lib.run(main, flags)
