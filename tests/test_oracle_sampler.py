"""Pins the sampler oracle: numpy's legacy stream (the reference's own dependency) == pure-python restatement
== C restatement == committed golden vectors."""
import os

import numpy as np
import pytest

from oracle import sampler as S

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


@pytest.mark.parametrize('alpha,rm', [(0.5, False), (0.3, True), (1.0, False), (0.05, False)])
def test_golden_matches_numpy_python_and_c(oracle_built, alpha, rm):
    gold = np.load(os.path.join(GOLD, 'sampler_mnist_a%s_rm%d.npz' % (alpha, int(rm))))
    C = S.one_coin_confusion(alpha)
    y_in = gold['y_in']
    a = S.mnist_labels_numpy(y_in, C, real_match=rm, seed=547)
    z = np.random.uniform(-1, 1, [8, 100]).astype(np.float32)
    assert np.array_equal(a['y_real'].argmax(1), gold['real']) and np.array_equal(a['y_fake'].argmax(1), gold['fake'])
    assert np.array_equal(a['y_gen'].argmax(1), gold['gen']) and np.array_equal(a['perm'], gold['perm'])
    assert np.array_equal(z, gold['z'])
    n = 600   # pure python is slow: prefix only, without the shuffle (stream position differs) -> compare to numpy directly
    b = S.mnist_labels_python(y_in[:n], C, real_match=rm, seed=547)
    a2 = S.mnist_labels_numpy(y_in[:n], C, real_match=rm, seed=547)
    assert np.array_equal(b['real'], a2['y_real'].argmax(1)) and np.array_equal(b['fake'], a2['y_fake'].argmax(1))
    assert np.array_equal(b['gen'], a2['y_gen'].argmax(1)) and np.array_equal(b['perm'], a2['perm'])
    c = S.mnist_labels_c(y_in, C, real_match=rm, seed=547)
    for k in ('perm', 'y', 'real', 'gen', 'fake'):
        assert np.array_equal(c[k], gold[k]), k
    # inverse-confusion weights are the rows of C^-1 at the noisy label (mnist/model.py:823)
    assert np.allclose(a['y_real_weights'], np.linalg.inv(C)[gold['real']])


def test_misc_golden(oracle_built):
    gold = np.load(os.path.join(GOLD, 'sampler_misc.npz'))
    lab, rnd, biased = S.cifar_labels_c(gold['cifar_in'], S.one_coin_confusion(0.5), 547)
    assert np.array_equal(lab, gold['cifar_labels']) and np.array_equal(rnd, gold['cifar_random'])
    assert np.array_equal(biased, gold['cifar_biased'])
    eye = np.eye(10)
    np.random.seed(11)
    r, f = S.mnist_renoise_numpy(eye[gold['re_real_in']], eye[gold['re_fake_in']], S.one_coin_confusion(0.6))
    assert np.array_equal(r.argmax(1), gold['re_real']) and np.array_equal(f.argmax(1), gold['re_fake'])


def test_confusion_matrices():
    C = S.one_coin_confusion(0.3)
    assert np.allclose(C.sum(1), 1) and np.allclose(np.diag(C), 0.3) and np.allclose(C[0, 1], 0.7 / 9)
    Cd = S.class_dependent_confusion(0.5)
    assert np.allclose(Cd.sum(1), 1)


def test_empty_and_degenerate_rows(oracle_built):
    C = np.eye(10)      # alpha = 1: p == 0 entries draw nothing, p == 1 draws once (numpy binomial semantics)
    y = np.arange(10).repeat(3)
    a = S.mnist_labels_numpy(y, C, seed=5, shuffle=False)
    c = S.mnist_labels_c(y, C, seed=5, shuffle=False)
    assert np.array_equal(a['y_real'].argmax(1), y) and np.array_equal(c['real'], y)
    assert np.array_equal(a['y_gen'].argmax(1), c['gen']) and np.array_equal(c['fake'], c['gen'])
