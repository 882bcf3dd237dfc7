"""The C-ABI library loads on a CPU-only box and exports every symbol include/rcgan_b200.h declares
(no compute calls without a GPU)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'rcgan_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(rcgan_[a-z0-9_]+)\s*\(', src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from robust_conditional_gan_b200 import _C
    syms = header_symbols()
    assert len(syms) >= 45
    for s in syms:
        assert hasattr(lib, s), 'declared but not exported: ' + s
    assert sorted(_C.EXPORTS) == syms, set(_C.EXPORTS) ^ set(syms)
    assert lib.rcgan_abi_version() == _C.ABI_VERSION
    hdr = open(os.path.join(ROOT, 'include', 'rcgan_b200.h')).read()
    assert int(re.search(r'#define RCGAN_ABI_VERSION (\d+)', hdr).group(1)) == _C.ABI_VERSION


def test_argument_validation_without_gpu(lib):
    from robust_conditional_gan_b200 import _C
    d = _C.ConvDesc(1, 4, 4, 8, 4, 4, 8, 3, 3, 1, 1, 1, 4, 8, _C.F32)   # ldx < cin
    with pytest.raises(_C.RcganError, match='ld smaller'):
        _C.call('rcgan_conv2d_fprop', d, None, None, None, None, None, 0, 0, 0.0, None)
    with pytest.raises(_C.RcganError, match='unsupported'):
        _C.call('rcgan_channel_loss', 1, None, 1, 1, 4, 48, 10, _C.F32, 0, 1.0, None, None, None, 0, None, None, None, None)


def test_sampler_table_host_matches_numpy_thresholds(lib):
    """the libm-dependent inversion constants are produced on the host exactly as numpy computes them"""
    import ctypes
    import math
    from oracle import sampler as S
    C = S.one_coin_confusion(0.3)
    tab = np.zeros(10 * 9 * 8)
    lib.rcgan_sampler_table_host(np.ascontiguousarray(C).ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 10,
                                 tab.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    tab = tab.reshape(10, 9, 8)
    Sum = 1.0
    for j in range(9):
        p = C[0, j] / Sum
        pp = p if p <= 0.5 else 1 - p
        assert tab[0, j, 1] == math.exp(math.log(1 - pp)) and tab[0, j, 3] == p
        assert tab[0, j, 4] == math.floor(math.ldexp(tab[0, j, 1], 53)) and tab[0, j, 5] >= tab[0, j, 4]
        Sum -= C[0, j]


def test_missing_library_fails_loudly(monkeypatch):
    from robust_conditional_gan_b200 import _C
    monkeypatch.setattr(_C, '_lib', None)
    monkeypatch.setattr(_C, 'LIB_PATH', '/nonexistent/librcgan_b200.so')
    with pytest.raises(_C.RcganError, match='no CPU or PyTorch fallback'):
        _C.load()
