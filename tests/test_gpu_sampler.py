"""Device label-noise sampler vs numpy's legacy stream (golden vectors + the oracle at full size): bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import sampler as OS
from robust_conditional_gan_b200.sampler import LabelNoiseSampler, class_dependent_confusion, one_coin_confusion

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')


@pytest.mark.parametrize('alpha,real_match', [(0.5, False), (0.3, True), (1.0, False), (0.05, False)])
def test_mnist_sampler_golden(lib, alpha, real_match):
    gold = np.load(os.path.join(GOLD, 'sampler_mnist_a%s_rm%d.npz' % (alpha, int(real_match))))
    s = LabelNoiseSampler('cuda')
    out = s.load_mnist_labels(gold['y_in'], one_coin_confusion(alpha), real_match=real_match, seed=547, shuffle=True)
    for k in ('perm', 'y', 'real', 'gen', 'fake'):
        assert np.array_equal(out[k], gold[k]), k
    z = s.uniform(-1, 1, gold['z'].size).cpu().numpy()
    assert np.array_equal(z, gold['z'].reshape(-1))       # the stream continues into batch_z bit-exactly


def test_mnist_sampler_full_size_vs_oracle(lib, oracle_built):
    """70 000 samples (the reference's train+test set), compared with the C oracle (itself pinned on numpy)."""
    y = np.random.RandomState(3).randint(10, size=70000)
    for C in (one_coin_confusion(0.5), class_dependent_confusion(0.5)):
        ref = OS.mnist_labels_c(y, C, seed=547)
        out = LabelNoiseSampler('cuda').load_mnist_labels(y, C, seed=547)
        for k in ('perm', 'y', 'real', 'gen', 'fake'):
            assert np.array_equal(out[k], ref[k]), k


def test_renoise_and_cifar(lib):
    gold = np.load(os.path.join(GOLD, 'sampler_misc.npz'))
    s = LabelNoiseSampler('cuda')
    s.seed(11)
    r2, f2 = s.renoise_mnist(gold['re_real_in'], gold['re_fake_in'], one_coin_confusion(0.6))
    assert np.array_equal(r2, gold['re_real']) and np.array_equal(f2, gold['re_fake'])
    r2, f2 = s.renoise_mnist(gold['re_real_in'], gold['re_fake_in'], np.eye(10))   # config 3: noise_C = I
    assert np.array_equal(r2, gold['re_real_in']) and np.array_equal(f2, gold['re_fake_in'])
    lab, rnd, biased = LabelNoiseSampler('cuda').cifar_labels(gold['cifar_in'], one_coin_confusion(0.5), seed=547)
    assert np.array_equal(lab, gold['cifar_labels']) and np.array_equal(rnd, gold['cifar_random'])
    assert np.array_equal(biased, gold['cifar_biased'])


def test_integer_threshold_path_equals_literal_double_path(lib, monkeypatch):
    """The sampler decides binomial(1,p) on 53-bit integers against floor(qn*2^53) and falls back to numpy's literal
    double-precision inversion loop only next to U = 1.  With RCGAN_SAMPLER_LITERAL=1 every success takes the literal
    loop: both must reproduce numpy bit for bit (golden vectors), on the one-coin and class-dependent matrices."""
    gold = np.load(os.path.join(GOLD, 'sampler_mnist_a0.5_rm0.npz'))
    y = np.random.RandomState(5).randint(10, size=20000)
    outs = []
    for literal in ('0', '1'):
        monkeypatch.setenv('RCGAN_SAMPLER_LITERAL', literal)
        s = LabelNoiseSampler('cuda')                # fresh table cache: the env var is read when the table is built
        out = s.load_mnist_labels(gold['y_in'], one_coin_confusion(0.5), seed=547)
        for k in ('perm', 'y', 'real', 'gen', 'fake'):
            assert np.array_equal(out[k], gold[k]), (literal, k)
        outs.append(LabelNoiseSampler('cuda').load_mnist_labels(y, class_dependent_confusion(0.4), seed=3))
    for k in ('real', 'gen', 'fake'):
        assert np.array_equal(outs[0][k], outs[1][k]), k
