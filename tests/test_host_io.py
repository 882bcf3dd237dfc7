"""Host-side pieces around the hot path, CPU only: the reference's flag syntax, checkpoints under TF variable
names, PNG / grid writers, metric arithmetic, dataset generators."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from robust_conditional_gan_b200 import checkpoint, data, flags as flags_lib, metrics, utils
from robust_conditional_gan_b200.graph import VariableStore


# ------------------------------------------------------------------------------------------------ flags
def test_mnist_flags_parse_like_the_run_scripts():
    """mnist/run_rcgany.sh's command line through the tf.app.flags stand-in (mnist/main.py:12-67)"""
    from robust_conditional_gan_b200 import main as M
    F = M.flags.FLAGS
    rest = F.parse('--algorithm rcgan --alpha 0.125 --disc_type projection --noestimate_confuse --noaux_classifier --add_noise '
                   '--noise_alpha 0.3 --noise_start 30 --noise_end 80 --concat_y --concat_y_layers 1,2 --spectral_norm --max_norm '
                   '--checkpoint_dir rcgany --script_file run_rcgany.sh --epoch=100'.split())
    assert rest == []
    assert F.algorithm == 'rcgan' and F.alpha == 0.125 and F.estimate_confuse is False and F.add_noise is True
    assert F.concat_y_layers == ['1', '2'] and F.epoch == 100 and F.batch_size == 100 and F.noise_end == 80
    assert F.learning_rate == 0.0002 and F.beta1 == 0.5 and F.recover_learning_rate == 500.0       # reference defaults
    M.configure(F)
    assert F.concat_y_layers == [1, 2] and F.input_height == 28 and F.output_width == 28
    assert F.checkpoint_dir.startswith(os.path.join('rcgany', 'rcgan_0.125_projection_')) and F.sample_dir.endswith('samples/')
    with pytest.raises(ValueError):
        F.parse(['--no_such_flag', '1'])
    # north_star shorthand
    F.parse(['--model', 'rcganu', '--checkpoint', 'x'])
    M.configure(F)
    assert F.algorithm == 'rcgan' and F.estimate_confuse is True and F.disc_type == 'projection'
    assert '__flags' and isinstance(getattr(F, '__flags'), dict)


def test_cifar_flags_and_batch_arithmetic():
    """gan_resnet.py:38-76, 183-192: log_file is mandatory, batch_size is multiplied and niters divided by ngpus"""
    from robust_conditional_gan_b200.cifar import main as M
    F = M.flags.FLAGS
    F.parse('--algorithm rcgan-u --alpha 0.5 --perm_classifier --confuse_init --batch_size 128 --ngpus 2 --niters 1000'.split())
    with pytest.raises(ValueError):
        M.configure(F)
    F.parse(['--log_file', '/tmp/x.log', '--expt_dir', 'e1', '--parent_dir', '/tmp'])
    cfg = M.configure(F)
    assert cfg['BATCH_SIZE'] == 256 and cfg['ITERS'] == 500 and cfg['DIR'] == '/tmp/e1'
    assert cfg['CHECKPOINT_DIR'] == '/tmp/e1/checkpoint'
    assert F.perm_classifier is True and F.confuse_init is True and F.perm_type == 'linear' and F.lr == 2e-4
    F.parse(['--nomulti_gpu_multi_batch'])
    assert M.configure(F)['BATCH_SIZE'] == 128


# ------------------------------------------------------------------------------------------------ checkpoints
def _toy_model(seed):
    g = torch.Generator().manual_seed(seed)
    st = VariableStore(torch.device('cpu'), lambda n: 'd' if n.startswith('discriminator') else ('c' if n == 'confusion_logits' else 'g'))
    st.get('generator/g_h0_lin/Matrix', (110, 64), lambda s: torch.randn(s, generator=g))
    st.get('generator/g_bn0/moving_mean', (64,), lambda s: torch.randn(s, generator=g), trainable=False)
    st.get('discriminator/d_h0_conv/w', (5, 5, 1, 64), lambda s: torch.randn(s, generator=g))
    st.get('discriminator/d_h0_conv/spectral_norm/u', (1, 64), lambda s: torch.randn(s, generator=g), trainable=False)
    st.get('confusion_logits', (10, 10), lambda s: torch.randn(s, generator=g))
    st.finalize()
    for k, grp in st.groups.items():
        grp.m.copy_(torch.randn(grp.numel, generator=g)); grp.v.copy_(torch.rand(grp.numel, generator=g)); grp.t = 7 + len(k)
    return SimpleNamespace(store=st)


def test_checkpoint_round_trip_under_tf_names(tmp_path):
    a, b = _toy_model(1), _toy_model(2)
    d = str(tmp_path / 'ck')
    p = checkpoint.save(a, d, 'DCGAN.model', 502)
    assert os.path.basename(p) == 'DCGAN.model-502' and checkpoint.latest_checkpoint(d) == p and checkpoint.step_of(p) == 502
    with np.load(p + '.npz') as z:
        keys = set(z.files)
    # what tf.train.Saver() writes: the variables (trainable or not) and the Adam slots of the trainable ones
    assert {'generator/g_h0_lin/Matrix', 'generator/g_h0_lin/Matrix/Adam', 'generator/g_h0_lin/Matrix/Adam_1',
            'generator/g_bn0/moving_mean', 'discriminator/d_h0_conv/spectral_norm/u', 'confusion_logits/Adam_1'} <= keys
    assert 'generator/g_bn0/moving_mean/Adam' not in keys
    checkpoint.restore(b, p)
    for n, v in a.store.vars.items():
        assert torch.equal(v.data, b.store.vars[n].data), n
    for k in a.store.groups:
        ga, gb = a.store.groups[k], b.store.groups[k]
        for var in ga.vars:                                   # (the arenas' alignment padding is not part of the checkpoint)
            sl = slice(var.offset, var.offset + var.numel())
            assert torch.equal(ga.m[sl], gb.m[sl]) and torch.equal(ga.v[sl], gb.v[sl]), var.name
        assert ga.t == gb.t
    # max_to_keep pruning and the Saver-style index file
    for step in (1002, 1502, 2002):
        checkpoint.save(a, d, 'DCGAN.model', step, max_to_keep=2)
    assert checkpoint.all_checkpoints(d) == ['DCGAN.model-1502', 'DCGAN.model-2002']
    assert not os.path.exists(os.path.join(d, 'DCGAN.model-502.npz')) and checkpoint.step_of(checkpoint.latest_checkpoint(d)) == 2002
    # weights-only dict (e.g. converted from a TF checkpoint reader): optimizer slots optional, missing variables reported
    sd = {n: v.data.reshape(v.shape).numpy() for n, v in a.store.vars.items() if 'confusion' not in n}
    with pytest.raises(KeyError):
        checkpoint.load_state(b, sd)
    assert checkpoint.load_state(b, sd, strict=False) == ['confusion_logits']
    assert checkpoint.latest_checkpoint(str(tmp_path / 'none')) is None


# ------------------------------------------------------------------------------------------------ writers
def test_png_writers(tmp_path):
    rs = np.random.RandomState(0)
    rgb = rs.randint(0, 256, size=(17, 23, 3)).astype(np.uint8)
    utils.write_png(str(tmp_path / 'a.png'), rgb)
    assert np.array_equal(utils.read_png(str(tmp_path / 'a.png')), rgb)
    grey = rs.randint(0, 256, size=(9, 31)).astype(np.uint8)
    utils.write_png(str(tmp_path / 'b.png'), grey)
    assert np.array_equal(utils.read_png(str(tmp_path / 'b.png')), grey)
    # mnist/utils.py:43-63 merge: row-major grid
    imgs = np.stack([np.full((4, 4, 1), i, dtype=np.float64) for i in range(6)])
    m = utils.merge(imgs, (2, 3))
    assert m.shape == (8, 12) and m[0, 0] == 0 and m[0, 4] == 1 and m[0, 8] == 2 and m[4, 0] == 3 and m[7, 11] == 5
    utils.save_images(imgs / 5. * 2 - 1, (2, 3), str(tmp_path / 'c.png'))          # inverse_transform then min-max bytescale
    c = utils.read_png(str(tmp_path / 'c.png'))
    assert c.shape == (8, 12) and c[0, 0] == 0 and c[7, 11] == 255
    # cifar10/common/misc.py:215-244: 100 samples -> 10 x 10 grid, ints kept as they are
    X = rs.randint(0, 256, size=(100, 32, 32, 3)).astype('int32')
    utils.save_images_grid(X, str(tmp_path / 'd.png'))
    d = utils.read_png(str(tmp_path / 'd.png'))
    assert d.shape == (320, 320, 3) and np.array_equal(d[32:64, 64:96], X[12].astype(np.uint8))


# ------------------------------------------------------------------------------------------------ metrics
def test_inception_score_arithmetic_and_label_accuracy():
    # uniform predictions -> score 1; perfectly confident and balanced over k classes -> score k (inception_score_.py:56-63)
    assert abs(metrics.preds2score(np.full((100, 10), 0.1), splits=5)[0] - 1.0) < 1e-12
    p = np.full((100, 10), 1e-12); p[np.arange(100), np.arange(100) % 10] = 1.0
    p /= p.sum(1, keepdims=True)
    m, s = metrics.preds2score(p, splits=5)
    assert abs(m - 10.0) < 1e-6 and s < 1e-9
    imgs = np.random.RandomState(0).uniform(-1, 1, size=(256, 3, 8, 8))
    sc = metrics.get_inception_score(imgs, lambda b: np.concatenate([b.reshape(len(b), -1)[:, :7] * 5, np.zeros((len(b), 1001))], 1),
                                     splits=2, batch_size=128)
    assert sc[0] > 1.0
    # generated-label accuracy with a learned confusion matrix that swaps classes 0 and 1 (gan_resnet.py:430-440)
    C = np.eye(10); C[[0, 1]] = C[[1, 0]]
    labels = np.array([0, 1, 2, 3])
    assert list(metrics.map_labels_through_confusion(labels, C)) == [1, 0, 2, 3]
    classify = lambda x: np.eye(10)[[1, 0, 2, 4]]
    assert metrics.generated_label_accuracy(None, labels, classify, confusion_matrix=C) == 0.75
    assert metrics.generated_label_accuracy(None, labels, classify) == 0.25
    # mnist/utils.py:273-306 regrouping: [R, 100, ...] with 10 consecutive samples per class inside each 100
    R = 100                                                  # the reference's reshape needs as many batches as the batch size
    samples = np.zeros((R, 100, 28, 28, 1))
    for cls in range(10):
        samples[:, cls * 10:(cls + 1) * 10] = cls
    assert metrics.generated_label_accuracy_mnist(samples, lambda x: x[:, 0, 0, 0].astype(int)) == 1.0
    assert abs(metrics.zero_one_loss(np.eye(10)[[1, 2, 3]], np.eye(10)[[1, 2, 4]]) - 1 / 3) < 1e-12


# ------------------------------------------------------------------------------------------------ data
def test_cifar_generators_follow_the_reference():
    """cifar10.py:20-45 with numpy's own stream (sampler=None) + gan_resnet.py:864-882 generators"""
    C = 0.5 / 9 * np.ones((10, 10)) + (0.5 - 0.5 / 9) * np.eye(10)
    gen = data.cifar_generator(['a'], 64, '', C, sampler=None, n_synthetic=640, seed=547)
    imgs, y = data.synthetic_cifar(640, seed=1)
    # replay with numpy literally
    np.random.seed(547)
    lr = np.random.randint(10, size=640)
    lab, lb = y.copy(), np.zeros(640, dtype=np.int64)
    for i in range(640):
        lab[i] = np.nonzero(np.random.multinomial(1, C[lab[i], :], size=1))[1][0]
        lb[i] = np.nonzero(np.random.multinomial(1, C[lr[i], :], size=1))[1][0]
    batches = list(gen())
    assert len(batches) == 10 and batches[0][0].dtype == np.uint8 and batches[0][0].shape == (64, 3072)
    assert np.array_equal(np.concatenate([b[1] for b in batches]), lab) and np.array_equal(np.concatenate([b[2] for b in batches]), lr)
    assert np.array_equal(np.concatenate([b[3] for b in batches]), lb)
    assert np.allclose(batches[3][4], np.linalg.inv(C)[lab[192:256]])
    # inf_train_gen_G: GEN_BS_MULTIPLE consecutive label batches from its own pass, restarting at the end of an epoch
    gG = data.inf_train_gen_G(gen, 2)
    first = next(gG)
    assert np.array_equal(first[0], lr[:128]) and np.array_equal(first[1], lb[:128])
    for _ in range(4):
        last = next(gG)
    assert np.array_equal(last[0], lr[512:640])
    again = next(gG)
    assert np.array_equal(again[0], lr[:128])
    g = data.inf_train_gen(gen)
    for _ in range(11):
        b = next(g)
    assert np.array_equal(b[1], lab[:64])                    # wrapped around to the first batch


def test_mnist_idx_reader(tmp_path):
    d = tmp_path / 'mnist'
    d.mkdir()
    rs = np.random.RandomState(0)
    tr, te = rs.randint(0, 256, size=(60000, 28, 28), dtype=np.uint8), rs.randint(0, 256, size=(10000, 28, 28), dtype=np.uint8)
    ytr, yte = rs.randint(0, 10, size=60000, dtype=np.uint8), rs.randint(0, 10, size=10000, dtype=np.uint8)
    for name, head, arr in (('train-images-idx3-ubyte', 16, tr), ('t10k-images-idx3-ubyte', 16, te), ('train-labels-idx1-ubyte', 8, ytr),
                            ('t10k-labels-idx1-ubyte', 8, yte)):
        with open(d / name, 'wb') as f:
            f.write(bytes(head) + arr.tobytes())
    X, y = data.mnist_data(str(tmp_path), 'mnist', allow_synthetic=False)
    assert X.shape == (70000, 28, 28, 1) and y.shape == (70000,) and X.max() <= 1.0
    assert np.array_equal((X[60000:, :, :, 0] * 255).round().astype(np.uint8), te) and np.array_equal(y[:60000], ytr)
    with pytest.raises(FileNotFoundError):
        data.mnist_data(str(tmp_path / 'nope'), 'mnist', allow_synthetic=False)
