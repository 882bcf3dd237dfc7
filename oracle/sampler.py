"""Oracle for the label-noise sampler (test infrastructure only).

Three layers, each checked against the one above in tests/test_oracle_sampler.py:

1. `mnist_labels_numpy` / `cifar_labels_numpy`: the reference's loops verbatim in
   behaviour, calling numpy's frozen legacy `np.random.*` exactly as
   mnist/model.py:795-834, :293-333 and cifar10/common/data/cifar10.py:29-38 do.
   numpy IS the reference's dependency here, so this layer is the pin.
2. `MT19937` + `LegacySampler`: a pure-Python restatement of the numpy legacy
   algorithms those calls reach (init_genrand, random_double, masked-rejection
   randint / shuffle, multinomial -> binomial by inversion).  Slow; small cases.
3. `oracle/sampler_c.c` (built into oracle/_build/liboracle_sampler.so by
   oracle/Makefile): the same algorithms in C, used for full-size (70 000) parity
   and as the timed CPU baseline of the sampler.
"""
import ctypes
import math
import os

import numpy as np


# ----------------------------------------------------------------------------- layer 1
def one_coin_confusion(alpha, y_dim=10):
    """mnist/model.py:809, cifar10/gan_resnet.py:106."""
    return ((1 - alpha) / (y_dim - 1.0)) * np.ones((y_dim, y_dim)) + \
        (alpha - (1 - alpha) / (y_dim - 1.0)) * np.eye(y_dim)


def class_dependent_confusion(alpha):
    """mnist/model.py:811-816 (np.linspace default num=50; first 10 entries used)."""
    C = np.zeros((10, 10))
    mean_diag = np.linspace(0.15, -0.15 + 2 * alpha)
    for i in range(10):
        C[i, :] = (1. - mean_diag[i]) / 9.
        C[i, i] = mean_diag[i]
    return C


def mnist_labels_numpy(y, C, real_match=False, seed=547, shuffle=True):
    """mnist/model.py:795-834.  `y`: int labels BEFORE the shuffle.  Consumes the
    global numpy stream exactly like the reference (seed, shuffle X, seed, shuffle y,
    then per sample multinomial / randint / multinomial).  Returns dict of arrays and
    leaves np.random in the post-loop state (the reference keeps drawing z from it)."""
    y = np.array(y, dtype=np.int64)
    n = len(y)
    perm = np.arange(n)
    if shuffle:
        np.random.seed(seed)
        np.random.shuffle(perm)          # same index stream as shuffle(X) / shuffle(y)
        np.random.seed(seed)
        np.random.shuffle(y)
    else:
        np.random.seed(seed)
    C_inv = np.linalg.inv(C)
    y_real = np.zeros((n, 10)); y_fake = np.zeros((n, 10)); y_gen = np.zeros((n, 10))
    y_actual = np.zeros((n, 10)); y_real_weights = np.zeros((n, 10))
    for i, label in enumerate(y):
        y_actual[i, label] = 1
        y_real[i] = np.random.multinomial(1, C[y[i], :], size=1)
        y_real_weights[i] = C_inv[np.where(y_real[i] == 1)[0], :]
        y_gen_label = np.random.randint(10, size=1)
        y_gen[i, int(y_gen_label[0])] = 1
        if real_match:
            y_gen[i] = y_real[i]
            y_gen_label = np.argmax(y_gen[i])
        else:
            y_gen_label = int(y_gen_label[0])
        y_fake[i] = np.random.multinomial(1, C[int(y_gen_label), :], size=1)
    return dict(perm=perm, y=y, y_actual=y_actual, y_real=y_real, y_gen=y_gen, y_fake=y_fake,
                y_real_weights=y_real_weights)


def mnist_renoise_numpy(y_real_orig, y_fake_orig, noise_C):
    """mnist/model.py:323-333 (per-epoch re-noising under --add_noise); continues the
    global stream."""
    y_real = np.zeros_like(y_real_orig); y_fake = np.zeros_like(y_fake_orig)
    for ii in range(len(y_real_orig)):
        y_real[ii] = np.random.multinomial(1, noise_C[np.argmax(y_real_orig[ii]), :], size=1)
        y_fake[ii] = np.random.multinomial(1, noise_C[np.argmax(y_fake_orig[ii]), :], size=1)
    return y_real, y_fake


def cifar_labels_numpy(labels, C, seed):
    """cifar10/common/data/cifar10.py:29-38 after np.random.seed(seed) (the reference
    sets no seed).  Returns (noisy labels, C_inv rows, random labels, biased labels)."""
    np.random.seed(seed)
    labels = np.array(labels, dtype=np.int64).copy()
    n = len(labels)
    C_inv = np.linalg.inv(C)
    inv_w = np.zeros((n, C.shape[0]))
    labels_random = np.random.randint(C.shape[0], size=n)
    labels_biased = np.zeros(n, dtype=np.int64)
    for i in range(n):
        labels[i] = int(np.where(np.random.multinomial(1, C[labels[i], :], size=1)[0] == 1)[0][0])
        inv_w[i] = C_inv[labels[i], :]
        labels_biased[i] = int(np.where(np.random.multinomial(1, C[labels_random[i], :], size=1)[0] == 1)[0][0])
    return labels, inv_w, labels_random, labels_biased


# ----------------------------------------------------------------------------- layer 2
class MT19937:
    """numpy legacy seeding for an int seed = init_genrand (mt19937_seed)."""
    N, M = 624, 397

    def __init__(self, seed):
        seed &= 0xffffffff
        self.key = [0] * self.N
        for pos in range(self.N):
            self.key[pos] = seed
            seed = (1812433253 * (seed ^ (seed >> 30)) + pos + 1) & 0xffffffff
        self.pos = self.N

    def _gen(self):
        k, N, M = self.key, self.N, self.M
        for i in range(N):
            y = (k[i] & 0x80000000) | (k[(i + 1) % N] & 0x7fffffff)
            k[i] = k[(i + M) % N] ^ (y >> 1) ^ (0x9908b0df if y & 1 else 0)
        self.pos = 0

    def next32(self):
        if self.pos == self.N:
            self._gen()
        y = self.key[self.pos]; self.pos += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9d2c5680
        y ^= (y << 15) & 0xefc60000
        y ^= y >> 18
        return y & 0xffffffff

    def double(self):
        a = self.next32() >> 5
        b = self.next32() >> 6
        return (a * 67108864.0 + b) / 9007199254740992.0


class LegacySampler:
    """Pure-python restatement of the numpy legacy distributions the reference uses."""

    def __init__(self, seed):
        self.rng = MT19937(seed)

    def interval(self, mx):
        """random_interval(max): masked rejection on 32-bit words (max <= 2^32-1)."""
        if mx == 0:
            return 0
        mask = mx
        for s in (1, 2, 4, 8, 16):
            mask |= mask >> s
        while True:
            v = self.rng.next32() & mask
            if v <= mx:
                return v

    def randint(self, high):
        return self.interval(high - 1)

    def shuffle_perm(self, n):
        perm = list(range(n))
        for i in range(n - 1, 0, -1):
            j = self.interval(i)
            perm[i], perm[j] = perm[j], perm[i]
        return perm

    def _inversion(self, n, p):
        q = 1.0 - p
        qn = math.exp(n * math.log(q))
        bound = min(n, n * p + 10.0 * math.sqrt(n * p * q + 1))
        X, px, U = 0, qn, self.rng.double()
        while U > px:
            X += 1
            if X > bound:
                X, px, U = 0, qn, self.rng.double()
            else:
                U -= px
                px = ((n - X + 1) * p * px) / (X * q)
        return X

    def binomial(self, n, p):
        if n == 0 or p == 0.0:
            return 0
        if p <= 0.5:
            return self._inversion(n, p)       # p*n <= 30 always for n=1
        return n - self._inversion(n, 1.0 - p)

    def multinomial1(self, pvals):
        """multinomial(1, pvals): index of the single 1."""
        d = len(pvals)
        Sum, dn = 1.0, 1
        for j in range(d - 1):
            x = self.binomial(dn, pvals[j] / Sum)
            dn -= x
            if dn <= 0:
                return j
            Sum -= pvals[j]
        return d - 1

    def uniform(self, lo, hi):
        return lo + (hi - lo) * self.rng.double()


def mnist_labels_python(y, C, real_match=False, seed=547, shuffle=True):
    y = list(int(v) for v in y)
    n = len(y)
    s = LegacySampler(seed)
    perm = list(range(n))
    if shuffle:
        perm = s.shuffle_perm(n)
        s = LegacySampler(seed)
        p2 = s.shuffle_perm(n)
        y = [y[k] for k in p2]
    lr, lg, lf = [], [], []
    for i in range(n):
        r = s.multinomial1(C[y[i]])
        g = s.randint(10)
        if real_match:
            g = r
        f = s.multinomial1(C[g])
        lr.append(r); lg.append(g); lf.append(f)
    return dict(perm=np.array(perm), y=np.array(y), real=np.array(lr), gen=np.array(lg), fake=np.array(lf),
                sampler=s)


# ----------------------------------------------------------------------------- layer 3
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_build', 'liboracle_sampler.so')
        if not os.path.exists(path):
            raise RuntimeError('oracle C sampler not built: run `make -C oracle` (or __graft_entry__.build())')
        _LIB = ctypes.CDLL(path)
        _LIB.oracle_mnist_labels.restype = ctypes.c_int
        _LIB.oracle_cifar_labels.restype = ctypes.c_int
    return _LIB


def mnist_labels_c(y, C, real_match=False, seed=547, shuffle=True):
    """C restatement, full size.  Returns int arrays (perm, shuffled y, real, gen, fake)."""
    lib = _lib()
    y = np.ascontiguousarray(y, dtype=np.int32).copy()
    n = len(y)
    C = np.ascontiguousarray(C, dtype=np.float64)
    perm = np.zeros(n, dtype=np.int32)
    real = np.zeros(n, dtype=np.int32); gen = np.zeros(n, dtype=np.int32); fake = np.zeros(n, dtype=np.int32)
    P = lambda a, t: a.ctypes.data_as(ctypes.POINTER(t))
    rc = lib.oracle_mnist_labels(ctypes.c_uint32(seed), ctypes.c_int(n), ctypes.c_int(int(shuffle)),
                                 ctypes.c_int(int(real_match)), P(C, ctypes.c_double), P(y, ctypes.c_int32),
                                 P(perm, ctypes.c_int32), P(real, ctypes.c_int32), P(gen, ctypes.c_int32),
                                 P(fake, ctypes.c_int32))
    assert rc == 0
    return dict(perm=perm, y=y, real=real, gen=gen, fake=fake)


def cifar_labels_c(labels, C, seed):
    lib = _lib()
    labels = np.ascontiguousarray(labels, dtype=np.int32).copy()
    n = len(labels)
    C = np.ascontiguousarray(C, dtype=np.float64)
    rnd = np.zeros(n, dtype=np.int32); biased = np.zeros(n, dtype=np.int32)
    P = lambda a, t: a.ctypes.data_as(ctypes.POINTER(t))
    rc = lib.oracle_cifar_labels(ctypes.c_uint32(seed), ctypes.c_int(n), P(C, ctypes.c_double),
                                 P(labels, ctypes.c_int32), P(rnd, ctypes.c_int32), P(biased, ctypes.c_int32))
    assert rc == 0
    return labels, rnd, biased
