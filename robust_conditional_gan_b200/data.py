"""Dataset readers and batch generators on either side of the hot path: the MNIST idx files
(mnist/model.py:770-793) and the CIFAR-10 python batches with their label-noise pass and epoch generators
(cifar10/common/data/cifar10.py:10-48, gan_resnet.py:864-885).  No dataset ships with the repository or is downloadable here:
every reader has a synthetic stand-in of the same shapes / dtypes so the entry scripts run end to end."""
import os
import pickle

import numpy as np


# ------------------------------------------------------------------------------------------------ MNIST
def read_mnist_idx(data_dir):
    """mnist/model.py:770-793: the four idx files -> (X [70000,28,28,1] float in [0,255], y [70000] int), train then test
    (the /255 and the seeded shuffle happen in DCGAN.load_mnist)."""
    def img(name, n):
        return np.fromfile(os.path.join(data_dir, name), dtype=np.uint8)[16:].reshape((n, 28, 28, 1)).astype(np.float64)

    def lab(name, n):
        return np.fromfile(os.path.join(data_dir, name), dtype=np.uint8)[8:].reshape((n,)).astype(np.int64)
    X = np.concatenate((img('train-images-idx3-ubyte', 60000), img('t10k-images-idx3-ubyte', 10000)), axis=0)
    y = np.concatenate((lab('train-labels-idx1-ubyte', 60000), lab('t10k-labels-idx1-ubyte', 10000)), axis=0)
    return X, y


def synthetic_mnist(n=70000, seed=0):
    """same shapes / ranges as read_mnist_idx"""
    rs = np.random.RandomState(seed)
    return rs.randint(0, 256, size=(n, 28, 28, 1)).astype(np.float64), rs.randint(10, size=n).astype(np.int64)


def mnist_data(data_dir, dataset_name='mnist', allow_synthetic=True, n_synthetic=70000):
    """(X in [0,1], y) for DCGAN(data=...): the idx files when present, else the synthetic stand-in."""
    d = os.path.join(data_dir, dataset_name)
    if os.path.exists(os.path.join(d, 'train-images-idx3-ubyte')):
        X, y = read_mnist_idx(d)
    elif allow_synthetic:
        X, y = synthetic_mnist(n_synthetic)
    else:
        raise FileNotFoundError('MNIST idx files not found under %s' % d)
    return X / 255., y


# ------------------------------------------------------------------------------------------------ CIFAR-10
def unpickle(file):
    """cifar10.py:10-17"""
    with open(file, 'rb') as fo:
        d = pickle.load(fo, encoding='bytes')
    return d[b'data'], d[b'labels']


def synthetic_cifar(n, seed=0):
    rs = np.random.RandomState(seed)
    return rs.randint(0, 256, size=(n, 3072)).astype(np.uint8), rs.randint(10, size=n).astype(np.int64)


def cifar_generator(filenames, batch_size, data_dir, C_ALPHA, sampler=None, n_synthetic=None, seed=None):
    """cifar10.py:20-45.  Reads the batch files (uint8 [N,3072] CHW), draws labels_random, then per sample the noisy label, its
    row of C^-1 and the biased label -- the three label draws run on the device (LabelNoiseSampler.cifar_labels, bit-exact with
    the numpy stream from the same state).  Returns get_epoch() yielding (images, labels, labels_random, labels_biased,
    labels_inv_weights) slices; images stay uint8 (the reference hands int arrays to an int32 placeholder).
    sampler None: the numpy loop itself (CPU-only callers); seed: re-seed first (the reference does not seed)."""
    if n_synthetic is not None:
        images, labels = synthetic_cifar(n_synthetic, seed=len(filenames))
    else:
        all_data, all_labels = [], []
        for filename in filenames:
            data, lab = unpickle(os.path.join(data_dir, filename))
            all_data.append(data)
            all_labels.append(lab)
        images = np.concatenate(all_data, axis=0)
        labels = np.concatenate(all_labels, axis=0)
    n = len(labels)
    C_inv = np.linalg.inv(C_ALPHA)
    if sampler is not None:
        labels, labels_random, labels_biased = sampler.cifar_labels(labels, C_ALPHA, seed=seed)
    else:
        if seed is not None:
            np.random.seed(seed)
        labels = np.array(labels, dtype=np.int64)
        labels_random = np.random.randint(10, size=n)
        labels_biased = np.zeros((n,), dtype=np.int64)
        for i in range(n):
            labels[i] = np.nonzero(np.random.multinomial(1, C_ALPHA[labels[i], :], size=1))[1][0]
            labels_biased[i] = np.nonzero(np.random.multinomial(1, C_ALPHA[labels_random[i], :], size=1))[1][0]
    labels_inv_weights = C_inv[labels, :]

    def get_epoch():
        for i in range(int(n / batch_size)):
            s = slice(i * batch_size, (i + 1) * batch_size)
            yield (images[s], labels[s], labels_random[s], labels_biased[s], labels_inv_weights[s])
    return get_epoch


TRAIN_FILES = ['data_batch_1', 'data_batch_2', 'data_batch_3', 'data_batch_4', 'data_batch_5']


def load(batch_size, data_dir, C_ALPHA, sampler=None, allow_synthetic=True):
    """cifar10.py:48-52: (train_gen, dev_gen)."""
    have = os.path.exists(os.path.join(data_dir, 'data_batch_1'))
    if not have and not allow_synthetic:
        raise FileNotFoundError('CIFAR-10 python batches not found under %s' % data_dir)
    return (cifar_generator(TRAIN_FILES, batch_size, data_dir, C_ALPHA, sampler, None if have else 50000),
            cifar_generator(['test_batch'], batch_size, data_dir, C_ALPHA, sampler, None if have else 10000))


def inf_train_gen(train_gen):
    """gan_resnet.py:864-868"""
    while True:
        for batch in train_gen():
            yield batch


def inf_train_gen_G(train_gen, gen_bs_multiple=2):
    """gan_resnet.py:869-882: the generator step's labels = GEN_BS_MULTIPLE consecutive batches of (random, biased) labels from
    its OWN pass over the training generator."""
    _generator = train_gen()
    while True:
        rnd, biased = [], []
        for _ in range(gen_bs_multiple):
            try:
                _, _, r, b, _ = next(_generator)
            except StopIteration:
                _generator = train_gen()
                _, _, r, b, _ = next(_generator)
            rnd.append(r)
            biased.append(b)
        yield (np.concatenate(rnd, axis=0), np.concatenate(biased, axis=0))
