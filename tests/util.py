"""Helpers shared by the GPU parity tests."""
import torch

from robust_conditional_gan_b200 import _C


def relerr(a, b):
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


def maxabs(a, b):
    return float((a.detach().double().cpu().reshape(-1) - b.detach().double().cpu().reshape(-1)).abs().max())


TD = {_C.F32: torch.float32, _C.BF16: torch.bfloat16}
TOL = {_C.F32: 2e-5, _C.BF16: 1e-2}   # relative L2; north_star: <=1e-4 fp32, <=1e-2 bf16


def dev(t, dtype=torch.float32):
    return t.to('cuda', dtype).contiguous()


def st():
    return torch.cuda.current_stream().cuda_stream


_KEEP = []


def keep(t):
    """Hold a device temporary alive (a bare `dev(x).data_ptr()` frees the tensor before the kernel runs and the
    caching allocator hands the same block to the next temporary) and return its pointer."""
    _KEEP.append(t)
    if len(_KEEP) > 256:
        import torch
        torch.cuda.synchronize()
        del _KEEP[:128]
    return t.data_ptr()
