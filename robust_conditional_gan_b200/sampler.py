"""Host side of the label-noise sampler (mnist/model.py:795-834, :293-333; cifar10/common/data/cifar10.py:
29-38): owns the device MT19937 state and calls the sampler kernels of librcgan_b200.so, which reproduce
numpy's legacy RandomState stream bit-exactly."""
import ctypes

import numpy as np
import torch

from . import _C
from ._C import call, stream_ptr


def one_coin_confusion(alpha, y_dim=10):
    """mnist/model.py:809; cifar10/gan_resnet.py:106."""
    return ((1 - alpha) / (y_dim - 1.0)) * np.ones((y_dim, y_dim)) + (alpha - (1 - alpha) / (y_dim - 1.0)) * np.eye(y_dim)


def class_dependent_confusion(alpha):
    """mnist/model.py:811-816."""
    C = np.zeros((10, 10))
    mean_diag = np.linspace(0.15, -0.15 + 2 * alpha)
    for i in range(10):
        C[i, :] = (1. - mean_diag[i]) / 9.
        C[i, i] = mean_diag[i]
    return C


class LabelNoiseSampler:
    def __init__(self, device='cuda'):
        self.device = torch.device(device)
        self.state = torch.zeros(_C.MT_STATE_WORDS, dtype=torch.int32, device=self.device)
        self._tables = {}

    def seed(self, seed):
        """np.random.seed(seed)"""
        call('rcgan_mt_seed', self.state.data_ptr(), int(seed) & 0xffffffff, stream_ptr())

    def table(self, C):
        """binomial-inversion constants of a confusion matrix, computed on the host with libm as numpy does"""
        C = np.ascontiguousarray(C, dtype=np.float64)
        key = C.tobytes()
        if key not in self._tables:
            k = C.shape[0]
            tab = np.zeros(k * (k - 1) * 8, dtype=np.float64)      # RCGAN_SAMPLER_TABLE_DOUBLES(k)
            _C.load().rcgan_sampler_table_host(C.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), k,
                                               tab.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
            self._tables[key] = torch.from_numpy(tab).to(self.device)
        return self._tables[key]

    def shuffle_perm(self, n):
        """the index permutation np.random.shuffle applies to an array of length n"""
        perm = torch.zeros(n, dtype=torch.int32, device=self.device)
        call('rcgan_mt_shuffle_perm', self.state.data_ptr(), perm.data_ptr(), n, stream_ptr())
        return perm

    def load_mnist_labels(self, y, C, real_match=False, seed=547, shuffle=True):
        """mnist/model.py:795-834.  Returns numpy int arrays: perm (apply to X), y (shuffled), real, gen, fake."""
        n = len(y)
        k = C.shape[0]
        self.seed(seed)
        if shuffle:
            perm = self.shuffle_perm(n)
            # the reference re-seeds and shuffles y: the same stream, hence the same permutation and the same final state
            yd = torch.as_tensor(np.asarray(y), dtype=torch.int32, device=self.device)[perm.long()].contiguous()
        else:
            perm = torch.arange(n, dtype=torch.int32, device=self.device)
            yd = torch.as_tensor(np.asarray(y), dtype=torch.int32, device=self.device)
        real = torch.zeros(n, dtype=torch.int32, device=self.device)
        gen = torch.zeros_like(real)
        fake = torch.zeros_like(real)
        call('rcgan_sample_labels_mnist', self.state.data_ptr(), self.table(C).data_ptr(), k, yd.data_ptr(), n,
             int(bool(real_match)), real.data_ptr(), gen.data_ptr(), fake.data_ptr(), stream_ptr())
        return dict(perm=perm.cpu().numpy().astype(np.int64), y=yd.cpu().numpy().astype(np.int64),
                    real=real.cpu().numpy().astype(np.int64), gen=gen.cpu().numpy().astype(np.int64),
                    fake=fake.cpu().numpy().astype(np.int64))

    def renoise_mnist(self, real, fake, noise_C):
        """mnist/model.py:323-333 (continues the stream)."""
        n = len(real)
        r = torch.as_tensor(np.asarray(real), dtype=torch.int32, device=self.device)
        f = torch.as_tensor(np.asarray(fake), dtype=torch.int32, device=self.device)
        r2, f2 = torch.zeros_like(r), torch.zeros_like(f)
        call('rcgan_sample_renoise_mnist', self.state.data_ptr(), self.table(noise_C).data_ptr(), noise_C.shape[0],
             r.data_ptr(), f.data_ptr(), n, r2.data_ptr(), f2.data_ptr(), stream_ptr())
        return r2.cpu().numpy().astype(np.int64), f2.cpu().numpy().astype(np.int64)

    def cifar_labels(self, labels, C, seed=None):
        """cifar10/common/data/cifar10.py:29-38.  Returns (noisy labels, random labels, biased labels)."""
        if seed is not None:
            self.seed(seed)
        n = len(labels)
        lab = torch.as_tensor(np.asarray(labels), dtype=torch.int32, device=self.device).clone()
        rnd, biased = torch.zeros_like(lab), torch.zeros_like(lab)
        call('rcgan_sample_labels_cifar', self.state.data_ptr(), self.table(C).data_ptr(), C.shape[0], lab.data_ptr(), n,
             rnd.data_ptr(), biased.data_ptr(), stream_ptr())
        return (lab.cpu().numpy().astype(np.int64), rnd.cpu().numpy().astype(np.int64),
                biased.cpu().numpy().astype(np.int64))

    def uniform(self, lo, hi, n):
        """np.random.uniform(lo, hi, n).astype(float32), as a device tensor"""
        out = torch.zeros(n, dtype=torch.float32, device=self.device)
        call('rcgan_mt_uniform', self.state.data_ptr(), out.data_ptr(), n, float(lo), float(hi), stream_ptr())
        return out
