"""Generates the sampler golden vectors by running the reference's own dependency (numpy legacy
np.random) through the reference's loops (oracle/sampler.py layer 1).  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', '..'))
from oracle import sampler as S

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    y_in = np.random.RandomState(1).randint(10, size=4000)
    for alpha, rm in [(0.5, False), (0.3, True), (1.0, False), (0.05, False)]:
        a = S.mnist_labels_numpy(y_in, S.one_coin_confusion(alpha), real_match=rm, seed=547)
        z = np.random.uniform(-1, 1, [8, 100]).astype(np.float32)       # batch_z continues the stream
        np.savez_compressed(os.path.join(HERE, 'sampler_mnist_a%s_rm%d.npz' % (alpha, int(rm))), y_in=y_in, perm=a['perm'],
                            y=a['y'], real=a['y_real'].argmax(1), gen=a['y_gen'].argmax(1), fake=a['y_fake'].argmax(1),
                            z=z)
    rs = np.random.RandomState(2)
    re_real_in, re_fake_in = rs.randint(10, size=3000), rs.randint(10, size=3000)
    eye = np.eye(10)
    np.random.seed(11)
    r, f = S.mnist_renoise_numpy(eye[re_real_in], eye[re_fake_in], S.one_coin_confusion(0.6))
    cifar_in = rs.randint(10, size=5000)
    lab, _, rnd, biased = S.cifar_labels_numpy(cifar_in, S.one_coin_confusion(0.5), 547)
    np.savez_compressed(os.path.join(HERE, 'sampler_misc.npz'), re_real_in=re_real_in, re_fake_in=re_fake_in,
                        re_real=r.argmax(1), re_fake=f.argmax(1), cifar_in=cifar_in, cifar_labels=lab, cifar_random=rnd,
                        cifar_biased=biased)


if __name__ == '__main__':
    main()
