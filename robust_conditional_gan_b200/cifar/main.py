"""Drop-in for running cifar10/gan_resnet.py: the same flags (gan_resnet.py:38-76), module constants (:141-192) and training
loop (:864-1014: inf_train_gen / inf_train_gen_G, iteration 0 without a G step, N_CRITIC D steps, learning-rate decay,
checkpoint restore / save, sample grids) on this package's CUDA path.

    python -m robust_conditional_gan_b200.cifar.main --algorithm rcgan-u --alpha 0.5 --perm_classifier --confuse_init \\
        --log_file run.log --ngpus 1                                          # = cifar10/run_rcganu.sh

Towers are ranks (launch with torchrun for --ngpus > 1; the reference's in-process 2-tower graph on one GPU is `--towers 2`).
Extra flags of this implementation: --precision, --towers, --data_dir, --synthetic, --rng."""
import logging
import os
import sys
import time
from datetime import datetime

import numpy as np

from .. import flags as flags_lib

flags = flags_lib.Flags()
flags.DEFINE_string("dataset", 'cifar', "Dataset")
flags.DEFINE_string("algorithm", 'rcgan', "Algorithm [rcgan, rcgan-u, biased, unbiased]")
flags.DEFINE_float("alpha", 0.8, "1 - noise level")
flags.DEFINE_string("run", '0', "run name")
flags.DEFINE_string("log_file", None, "logging file")
flags.DEFINE_string("parent_dir", '.', "parent directory for checkpoints")
flags.DEFINE_string("expt_dir", None, "directory for expts")
flags.DEFINE_integer("inception_freq", 2500, "frequncy of inception score calculation")
flags.DEFINE_integer("sample_freq", 2500, "frequncy of dev cost calc. and sample pics")
flags.DEFINE_integer("generated_label_accuracy_freq", 2500, "frequncy of generated label accruacy")
flags.DEFINE_integer("sample_save_freq", 0, "frequncy of saving samples")
flags.DEFINE_integer("batch_size", 64, "batch size")
flags.DEFINE_integer("niters", 50000, "no. of batches")
flags.DEFINE_float("lr", 2.0e-4, "learning rate")
flags.DEFINE_integer("ngpus", 2, "no. of gpus")
flags.DEFINE_boolean("multi_gpu_multi_batch", True, 'whether to multiply batch_size with number of gpus'
                     'and divide nof. iterations by nof. gpus')
flags.DEFINE_boolean("confuse_init", False, "whether to initialize confusion matrix with identity")
flags.DEFINE_float("confuse_init_diag", 0.2, "intial confusion matrix with diagonal entry")
flags.DEFINE_float("confuse_multiplier", 1.0, "learning rate multiplier for learnable confusion matrix ")
flags.DEFINE_boolean("confuse_lr_decay", False, 'whether to decay confusion matrix estimation learning rate')
flags.DEFINE_boolean("perm_classifier", False, 'whether to real fake classifier or not.')
flags.DEFINE_float("perm_multiplier", 1.0, 'whether to real fake classifier or not.')
flags.DEFINE_string("perm_type", 'linear', 'type of real fake classifier to use [linear, 2layer].')
flags.DEFINE_boolean("restore", True, 'whether to restore from past checkpoint')
flags.DEFINE_boolean("perm_gen_label_acc", False, 'whether to calculate generated label accuracy'
                     'by taking min. value over all permutation of labels')
flags.DEFINE_string("log_level", 'info', 'logging level [info, debug]')
# ---- this implementation
flags.DEFINE_string("model", None, "north_star shorthand: rcgan | rcganu | biased | unbiased (sets --algorithm and its run script's flags)")
flags.DEFINE_string("precision", "bf16", "bf16 (tensor-core path) | fp32 (parity mode)")
flags.DEFINE_integer("towers", 1, "towers per process (2 = the reference's single-GPU graph: two towers of batch/2 on one device)")
flags.DEFINE_string("data_dir", '../data/cifar10/cifar-10-batches-py/', "CIFAR-10 python batches")
flags.DEFINE_boolean("synthetic", True, "synthetic 32x32x3 data when the batches are absent")
flags.DEFINE_string("rng", 'device', "device: in-graph noise like the reference; fed: noise drawn on the host and fed")
FLAGS = flags.FLAGS

MODEL_FLAGS = {'rcgan': dict(algorithm='rcgan'), 'rcganu': dict(algorithm='rcgan-u', perm_classifier=True, confuse_init=True),
               'biased': dict(algorithm='biased'), 'unbiased': dict(algorithm='unbiased')}


def configure(FLAGS):
    """gan_resnet.py:78-192: required flags, run directory, effective batch size / iteration count."""
    if FLAGS.model is not None:
        for k, v in MODEL_FLAGS[FLAGS.model].items():
            setattr(FLAGS, k, v)
    if FLAGS.log_file is None:
        raise ValueError('flag log_file is required')
    if FLAGS.ngpus not in [1, 2] and int(os.environ.get('WORLD_SIZE', '1')) != FLAGS.ngpus:
        # the reference stops at 2 GPUs (:184); ranks lift that as long as there is one process per GPU
        raise Exception('ngpus = %d needs a torchrun launch with %d ranks' % (FLAGS.ngpus, FLAGS.ngpus))
    cfg = {}
    cfg['DIR'] = os.path.join(FLAGS.parent_dir, FLAGS.algorithm + '_alpha' + str(FLAGS.alpha) + '_run-' + FLAGS.run + '_' +
                              datetime.now().strftime("%Y%m%d-%H%M%S"))
    if FLAGS.expt_dir is not None:
        cfg['DIR'] = '{}/{}'.format(FLAGS.parent_dir, FLAGS.expt_dir)
    cfg['BATCH_SIZE'], cfg['ITERS'] = FLAGS.batch_size, FLAGS.niters
    if FLAGS.multi_gpu_multi_batch:                         # :190-192
        cfg['BATCH_SIZE'] = FLAGS.batch_size * FLAGS.ngpus
        cfg['ITERS'] = FLAGS.niters // FLAGS.ngpus
    cfg['CHECKPOINT_DIR'] = os.path.join(cfg['DIR'], 'checkpoint')
    return cfg


def train(FLAGS, cfg, max_iters=None):
    import torch
    import torch.distributed as dist
    from .. import checkpoint, data, utils
    from ..feeds import Prefetcher
    from ..sampler import LabelNoiseSampler
    from .gan_resnet import GEN_BS_MULTIPLE, N_CRITIC, RCGANCifar
    rank, world, local = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    os.makedirs(cfg['DIR'], exist_ok=True)
    logging.basicConfig(filename=FLAGS.log_file, level=logging.DEBUG if FLAGS.log_level == 'debug' else logging.INFO,
                        format='%(asctime)s %(levelname)-8s %(message)s', force=True)
    logging.info('alpha = {}'.format(FLAGS.alpha))
    B = cfg['BATCH_SIZE']
    assert B % (world * FLAGS.towers) == 0, 'BATCH_SIZE must divide into ranks x towers'
    n = B // world                                           # this rank's share of every batch (tf.split, :529-538)
    model = RCGANCifar(FLAGS, tower_batch=n // FLAGS.towers, precision=FLAGS.precision, world_size=world, rank=rank, rng=FLAGS.rng,
                       towers=FLAGS.towers)
    smp = LabelNoiseSampler('cuda')
    train_gen, dev_gen = data.load(B, FLAGS.data_dir, model.C_ALPHA, sampler=smp, allow_synthetic=FLAGS.synthetic)
    gen, gen_G = data.inf_train_gen(train_gen), data.inf_train_gen_G(train_gen, GEN_BS_MULTIPLE)
    start = 0
    if FLAGS.restore:                                        # :909-914
        ckpt = checkpoint.latest_checkpoint(cfg['CHECKPOINT_DIR'])
        if ckpt:
            logging.info('restore model from: {}...'.format(ckpt))
            extra = checkpoint.restore(model, ckpt)
            start = int(extra.get('iteration', checkpoint.step_of(ckpt))) + 1
    sh = lambda a, mult=1: a[rank * n * mult:(rank + 1) * n * mult]     # contiguous shard of the global batch
    rs = np.random.RandomState(1234 + rank)
    pf_d, pf_g = Prefetcher(model.d_prog, 'cuda'), Prefetcher(model.g_prog, 'cuda')

    def stage_d():
        _data, _labels, _random_labels, _labels_biased, _labels_inv_weights = next(gen)
        f = dict(all_real_data_int=sh(_data), all_real_labels=sh(_labels), all_random_labels=sh(_random_labels),
                 all_labels_biased=sh(_labels_biased), all_labels_inv_weights=sh(_labels_inv_weights))
        if FLAGS.rng != 'device':
            f.update(noise=rs.randn(n, 128).astype(np.float32), dequant_noise=(rs.rand(n, 3072) / 128).astype(np.float32))
        pf_d.prefetch(**f)

    def stage_g():
        _random_labels_G, _labels_biased_G = next(gen_G)
        f = dict(all_random_labels_G=sh(_random_labels_G, GEN_BS_MULTIPLE), all_labels_biased_G=sh(_labels_biased_G, GEN_BS_MULTIPLE))
        if FLAGS.rng != 'device':
            f.update(noise=rs.randn(GEN_BS_MULTIPLE * n, 128).astype(np.float32))
        pf_g.prefetch(**f)

    stage_g()
    stage_d()
    t0 = time.time()
    iters = cfg['ITERS'] if max_iters is None else min(cfg['ITERS'], start + max_iters)
    for iteration in range(start, iters):
        model.iteration = iteration
        if iteration > 0:                                    # :927-934
            pf_g.commit()
            model.g_step(iteration)
            stage_g()
        for i in range(N_CRITIC):                            # :936-947
            pf_d.commit()
            model.d_step(iteration)
            stage_d()                                        # the next batch is staged and copied while this step runs
        if rank == 0 and (iteration % 100 == 99 or iteration == iters - 1):
            out = model.fetch_losses()
            logging.info('iter %d  d_cost %.5f  g_cost %.5f  (%.1f real img/s)', iteration, out['disc_real_l'] + out['disc_fake_l'],
                         out.get('gen_wgan', float('nan')), (iteration - start + 1) * N_CRITIC * B / (time.time() - t0))
        if rank == 0 and FLAGS.sample_freq and iteration % FLAGS.sample_freq == FLAGS.sample_freq - 1:      # :846-850, 976-979
            labels = np.repeat(np.arange(10), 10)
            samples = model.sample(labels, noise=np.random.RandomState(0).randn(100, 128).astype(np.float32))
            utils.save_images_grid(((samples + 1.) * (255. / 2)).astype('int32'), os.path.join(cfg['DIR'], 'samples_{}.png'.format(iteration)))
        if rank == 0 and ((iteration < 500 and iteration % 100 == 99) or iteration % 1000 == 999 or iteration == iters - 1):   # :1003-1011
            checkpoint.save(model, cfg['CHECKPOINT_DIR'], 'model.ckpt', iteration, max_to_keep=5, extra={'iteration': iteration})
    torch.cuda.synchronize()
    return model


def main(_):
    cfg = configure(FLAGS)
    return train(FLAGS, cfg)


if __name__ == '__main__':
    flags_lib.run(main, flags)
