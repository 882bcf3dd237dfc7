"""CUDA conv/deconv/linear family vs the oracle (oracle/nn.py) on the same seeded inputs."""
import pytest
import torch

from oracle import nn as O
from robust_conditional_gan_b200 import _C
from robust_conditional_gan_b200._C import ConvDesc, call
from robust_conditional_gan_b200.graph import same_pad
from util import keep, TD, TOL, dev, relerr, st

pytestmark = pytest.mark.gpu

# (n, h, w, cin, cout, k, stride, ldx_pad, ldy_pad)
SHAPES = [
    (3, 28, 28, 1, 64, 5, 2, 0, 0),      # d_h0_conv
    (3, 28, 28, 11, 64, 5, 2, 5, 0),     # d_h0_conv with label concat (ld 16)
    (2, 14, 14, 64, 64, 5, 2, 0, 0),     # d_h1_conv
    (2, 7, 7, 64, 64, 5, 2, 0, 0),       # d_h2_conv
    (5, 4, 4, 64, 64, 5, 2, 0, 0),       # d_h3_conv
    (2, 14, 14, 128, 138, 5, 2, 0, 6),   # g_h2 as the conv it transposes (x: 14x14x128, y: 7x7x138 ld 144)
    (2, 28, 28, 1, 138, 5, 2, 0, 6),     # g_h3
    (2, 8, 8, 128, 128, 3, 1, 0, 0),     # CIFAR D 3x3
    (2, 16, 16, 3, 128, 1, 1, 0, 0),     # CIFAR D shortcut 1x1 cin=3
    (2, 32, 32, 256, 3, 3, 1, 0, 0),     # G.Output cout=3
    (7, 1, 1, 110, 1024, 1, 1, 2, 0),    # g_h0_lin (ld 112)
    (9, 1, 1, 784, 10, 1, 1, 0, 0),      # classifier
    (33, 1, 1, 64, 1, 1, 1, 0, 0),       # d_h4_lin
    (300, 1, 1, 784, 10, 1, 1, 0, 0),    # classifier at a batch that takes the warp-per-row skinny kernel
    (257, 1, 1, 100, 16, 1, 1, 4, 0),    # skinny kernel, N = 16, padded rows
    (512, 1, 1, 3072, 10, 1, 1, 0, 0),   # CIFAR permutation classifier (gan_resnet.py:458-466): skinny kernel, 3 chunks of K
    (259, 1, 1, 2100, 12, 1, 1, 4, 0),   # skinny kernel, ragged last K chunk, ragged last row group
]


def make(shape, dtype, seed=0):
    n, h, w, cin, cout, k, s, px, py = shape
    g = torch.Generator().manual_seed(seed)
    ho, pt = same_pad(h, k, s)
    wo, pl = same_pad(w, k, s)
    ldx, ldy = cin + px, cout + py
    x = torch.randn(n, h, w, cin, generator=g)
    wt = torch.randn(k, k, cin, cout, generator=g) * 0.1
    b = torch.randn(cout, generator=g)
    dy = torch.randn(n, ho, wo, cout, generator=g)
    td = TD[dtype]
    x, dy = x.to(td).float(), dy.to(td).float()      # quantise inputs so both sides see the same values
    xbuf = torch.zeros(n, h, w, ldx); xbuf[..., :cin] = x
    dybuf = torch.zeros(n, ho, wo, ldy); dybuf[..., :cout] = dy
    d = ConvDesc(n, h, w, cin, ho, wo, cout, k, k, s, pt, pl, ldx, ldy, dtype)
    return d, x, wt, b, dy, dev(xbuf, td), dev(dybuf, td), (ho, wo, ldx, ldy)


@pytest.mark.parametrize('dtype', [_C.F32, _C.BF16])
@pytest.mark.parametrize('shape', SHAPES)
def test_fprop(lib, shape, dtype):
    d, x, wt, b, dy, xd, dyd, (ho, wo, ldx, ldy) = make(shape, dtype)
    n, cout = shape[0], shape[4]
    y = torch.full((n, ho, wo, ldy), 7.0, device='cuda', dtype=TD[dtype])
    call('rcgan_conv2d_fprop', d, xd.data_ptr(), keep(dev(wt)), None, keep(dev(b)), y.data_ptr(), dtype, _C.ACT_LRELU,
         0.2, st())
    ref = O.lrelu(O.conv2d(x.double(), wt.double(), shape[6]) + b.double())
    assert relerr(y[..., :cout].float(), ref) < TOL[dtype]
    if ldy > cout:
        assert float((y[..., cout:].float() - 7.0).abs().max()) == 0.0     # padding untouched


@pytest.mark.parametrize('dtype', [_C.F32, _C.BF16])
@pytest.mark.parametrize('shape', SHAPES)
def test_dgrad_and_deconv(lib, shape, dtype):
    d, x, wt, b, dy, xd, dyd, (ho, wo, ldx, ldy) = make(shape, dtype)
    n, h, w, cin, cout, k, s = shape[:7]
    xr = x.double().requires_grad_(True)
    O.conv2d(xr, wt.double(), s).backward(dy.double())
    dx = torch.zeros(n, h, w, ldx, device='cuda', dtype=TD[dtype])
    call('rcgan_conv2d_dgrad', d, dyd.data_ptr(), keep(dev(wt)), None, None, dx.data_ptr(), dtype, _C.ACT_NONE, 0.0, 0, st())
    assert relerr(dx[..., :cin].float(), xr.grad) < TOL[dtype]
    # accumulate
    call('rcgan_conv2d_dgrad', d, dyd.data_ptr(), keep(dev(wt)), None, None, dx.data_ptr(), dtype, _C.ACT_NONE, 0.0, 1, st())
    assert relerr(dx[..., :cin].float(), 2 * xr.grad) < 2 * TOL[dtype]
    # as the forward of deconv2d: conv2d_transpose(dy, w[kh,kw,cin(out),cout(in)]) + bias, sigmoid
    if s == 2:
        bias = torch.randn(cin, generator=torch.Generator().manual_seed(5))
        out = torch.zeros(n, h, w, ldx, device='cuda', dtype=TD[dtype])
        call('rcgan_conv2d_dgrad', d, dyd.data_ptr(), keep(dev(wt)), None, keep(dev(bias)), out.data_ptr(),
             dtype, _C.ACT_SIGMOID, 0.0, 0, st())
        ref = torch.sigmoid(O.conv2d_transpose(dy.double(), wt.double(), (h, w), s) + bias.double())
        assert relerr(out[..., :cin].float(), ref) < TOL[dtype]


@pytest.mark.parametrize('dtype', [_C.F32, _C.BF16])
@pytest.mark.parametrize('shape', SHAPES)
def test_wgrad(lib, shape, dtype):
    d, x, wt, b, dy, xd, dyd, _ = make(shape, dtype)
    s = shape[6]
    wr = wt.double().requires_grad_(True)
    O.conv2d(x.double(), wr, s).backward(dy.double())
    nb = lib.rcgan_conv2d_wgrad_workspace(d)
    ws = torch.zeros(max(nb, 4), dtype=torch.uint8, device='cuda')
    dw = torch.full(wt.shape, 3.0, device='cuda')
    call('rcgan_conv2d_wgrad', d, xd.data_ptr(), dyd.data_ptr(), dw.data_ptr(), 0, ws.data_ptr(), nb, st())
    assert relerr(dw, wr.grad) < TOL[dtype]
    call('rcgan_conv2d_wgrad', d, xd.data_ptr(), dyd.data_ptr(), dw.data_ptr(), 1, ws.data_ptr(), nb, st())
    assert relerr(dw, 2 * wr.grad) < 2 * TOL[dtype]
    db = torch.zeros(shape[4], device='cuda')
    call('rcgan_colsum', dyd.data_ptr(), dyd.numel() // dyd.shape[-1], shape[4], dyd.shape[-1], dtype, db.data_ptr(), 0, st())
    assert relerr(db, dy.double().sum((0, 1, 2))) < TOL[dtype]


@pytest.mark.parametrize('shape', [SHAPES[2], SHAPES[5], SHAPES[10]])
def test_bf16_operands_fp32_output(lib, shape):
    """pre-norm layout: bf16 operands, fp32 stored output (conv -> batch norm); must be MORE accurate than bf16 storage"""
    d, x, wt, b, dy, xd, dyd, (ho, wo, ldx, ldy) = make(shape, _C.BF16)
    n, h, w, cin, cout, k, s = shape[:7]
    y = torch.zeros(n, ho, wo, ldy, device='cuda', dtype=torch.float32)
    call('rcgan_conv2d_fprop', d, xd.data_ptr(), keep(dev(wt)), None, keep(dev(b)), y.data_ptr(), _C.F32, _C.ACT_NONE, 0.0, st())
    ref = O.conv2d(x.double(), wt.double(), s) + b.double()
    assert relerr(y[..., :cout], ref) < 2e-5
    if s == 2:
        out = torch.zeros(n, h, w, ldx, device='cuda', dtype=torch.float32)
        call('rcgan_conv2d_dgrad', d, dyd.data_ptr(), keep(dev(wt)), None, None, out.data_ptr(), _C.F32, _C.ACT_NONE, 0.0, 0, st())
        assert relerr(out[..., :cin], O.conv2d_transpose(dy.double(), wt.double(), (h, w), s)) < 2e-5


def test_wgrad_large_k_splits(lib):
    """split-K path: K = n*ho*wo = 50176 (d_h1 at batch 1024 has 50176 output pixels)"""
    shape = (256, 14, 14, 64, 64, 5, 2, 0, 0)
    d, x, wt, b, dy, xd, dyd, _ = make(shape, _C.F32)
    wr = wt.double().requires_grad_(True)
    O.conv2d(x.double(), wr, 2).backward(dy.double())
    nb = lib.rcgan_conv2d_wgrad_workspace(d)
    ws = torch.zeros(nb, dtype=torch.uint8, device='cuda')
    dw = torch.zeros(wt.shape, device='cuda')
    call('rcgan_conv2d_wgrad', d, xd.data_ptr(), dyd.data_ptr(), dw.data_ptr(), 0, ws.data_ptr(), nb, st())
    assert relerr(dw, wr.grad) < 2e-5


def test_bad_args_fail_loudly(lib):
    d = ConvDesc(1, 4, 4, 8, 4, 4, 8, 3, 3, 1, 1, 1, 4, 8, _C.F32)      # ldx < cin
    with pytest.raises(_C.RcganError, match='ld smaller'):
        call('rcgan_conv2d_fprop', d, None, None, None, None, None, 0, 0, 0.0, st())


# ----------------------------------------------------------------------------- tcgen05 path
TC_SHAPES = [
    (4, 14, 14, 64, 64, 5, 2, 0, 0),      # d_h1_conv
    (4, 7, 7, 64, 64, 5, 2, 0, 0),        # d_h2_conv
    (40, 4, 4, 64, 64, 5, 2, 0, 0),       # d_h3_conv
    (3, 14, 14, 128, 138, 5, 2, 0, 6),    # g_h2 (deconv = dgrad), y side ld 144
    (5, 28, 28, 64, 74, 5, 2, 0, 6),      # odd channel counts, ld 80
    (4, 32, 32, 128, 128, 3, 1, 0, 0),    # CIFAR D 3x3
    (2, 16, 16, 256, 256, 3, 1, 0, 0),    # CIFAR G 3x3
    (2, 8, 8, 1024, 256, 1, 1, 0, 0),     # CIFAR G shortcut 1x1
    (300, 1, 1, 110, 1024, 1, 1, 2, 0),   # g_h0_lin
    (70, 1, 1, 1034, 896, 1, 1, 6, 0),    # g_h1_lin (N reduced)
    (130, 1, 1, 64, 32, 1, 1, 0, 0),      # small N
]


def pack_for(lib, d, wt):
    nb = lib.rcgan_conv_wpack_bytes(d)
    assert nb > 0
    pack = torch.zeros(nb, dtype=torch.uint8, device='cuda')
    wd = dev(wt)
    call('rcgan_conv_wpack', d, wd.data_ptr(), None, pack.data_ptr(), st())
    return pack, wd


@pytest.mark.parametrize('shape', TC_SHAPES)
def test_tc_fprop(lib, shape):
    d, x, wt, b, dy, xd, dyd, (ho, wo, ldx, ldy) = make(shape, _C.BF16)
    n, cout, s = shape[0], shape[4], shape[6]
    assert lib.rcgan_conv_uses_tensor_cores(d, 0) == 1
    pack, wd = pack_for(lib, d, wt)
    bd = dev(b)
    ref = O.lrelu(O.conv2d(x.double(), wt.to(torch.bfloat16).double(), s) + b.double())
    for odt in (_C.BF16, _C.F32):
        y = torch.full((n, ho, wo, ldy), 7.0, device='cuda', dtype=TD[odt])
        call('rcgan_conv2d_fprop', d, xd.data_ptr(), wd.data_ptr(), pack.data_ptr(), bd.data_ptr(), y.data_ptr(), odt,
             _C.ACT_LRELU, 0.2, st())
        torch.cuda.synchronize()
        assert relerr(y[..., :cout].float(), ref) < (6e-3 if odt == _C.BF16 else 2e-5), odt
        if ldy > cout:
            assert float((y[..., cout:].float() - 7.0).abs().max()) == 0.0


@pytest.mark.parametrize('shape', TC_SHAPES)
def test_tc_dgrad_and_deconv(lib, shape):
    d, x, wt, b, dy, xd, dyd, (ho, wo, ldx, ldy) = make(shape, _C.BF16)
    n, h, w, cin, cout, k, s = shape[:7]
    assert lib.rcgan_conv_uses_tensor_cores(d, 1) == 1
    pack, wd = pack_for(lib, d, wt)
    wq = wt.to(torch.bfloat16).double()
    xr = x.double().requires_grad_(True)
    O.conv2d(xr, wq, s).backward(dy.double())
    dx = torch.zeros(n, h, w, ldx, device='cuda', dtype=torch.float32)
    call('rcgan_conv2d_dgrad', d, dyd.data_ptr(), wd.data_ptr(), pack.data_ptr(), None, dx.data_ptr(), _C.F32, _C.ACT_NONE, 0.0,
         0, st())
    torch.cuda.synchronize()
    assert relerr(dx[..., :cin], xr.grad) < 2e-5
    call('rcgan_conv2d_dgrad', d, dyd.data_ptr(), wd.data_ptr(), pack.data_ptr(), None, dx.data_ptr(), _C.F32, _C.ACT_NONE, 0.0,
         1, st())
    assert relerr(dx[..., :cin], 2 * xr.grad) < 2e-5                 # accumulate
    bias = torch.randn(cin, generator=torch.Generator().manual_seed(5))
    out = torch.zeros(n, h, w, ldx, device='cuda', dtype=torch.bfloat16)
    call('rcgan_conv2d_dgrad', d, dyd.data_ptr(), wd.data_ptr(), pack.data_ptr(), keep(dev(bias)), out.data_ptr(), _C.BF16,
         _C.ACT_SIGMOID, 0.0, 0, st())
    if s == 2:
        ref = torch.sigmoid(O.conv2d_transpose(dy.double(), wq, (h, w), s) + bias.double())
    else:
        ref = torch.sigmoid(xr.grad + bias.double())
    assert relerr(out[..., :cin].float(), ref) < 6e-3


def test_tc_full_size_linearity(lib):
    """BASELINE config-2 size (d_h1_conv at batch 1024): conv(a*x1 + x2) == a*conv(x1) + conv(x2) through the
    tcgen05 path (fp32 output), a size-independent property where the oracle would take minutes."""
    shape = (1024, 14, 14, 64, 64, 5, 2, 0, 0)
    n, h, w, cin, cout, k, s = shape[:7]
    g = torch.Generator().manual_seed(9)
    d = ConvDesc(n, h, w, cin, 7, 7, cout, k, k, s, 1, 1, cin, cout, _C.BF16)
    wt = torch.randn(k, k, cin, cout, generator=g) * 0.1
    pack, wd = pack_for(lib, d, wt)
    # power-of-two scale and disjoint-exponent inputs keep a*x1 + x2 exactly representable in bf16
    x1 = torch.randint(-4, 5, (n, h, w, cin), generator=g).float()
    x2 = torch.randint(-4, 5, (n, h, w, cin), generator=g).float()
    outs = []
    for xin in (x1, x2, 2 * x1 + x2):
        y = torch.zeros(n, 7, 7, cout, device='cuda')
        xd = dev(xin, torch.bfloat16)
        call('rcgan_conv2d_fprop', d, xd.data_ptr(), wd.data_ptr(), pack.data_ptr(), None, y.data_ptr(), _C.F32, _C.ACT_NONE, 0.0,
             st())
        torch.cuda.synchronize()
        outs.append(y)
    assert relerr(outs[2], 2 * outs[0] + outs[1]) < 1e-5
    assert float(outs[0].abs().max()) > 0


@pytest.mark.parametrize('shape', TC_SHAPES + [(256, 14, 14, 64, 64, 5, 2, 0, 0), (64, 32, 32, 128, 128, 3, 1, 0, 0)])
def test_tc_wgrad(lib, shape):
    """tcgen05 wgrad (MN-major operands, split-K with fp32 atomics) vs the fp64 oracle on bf16-exact inputs"""
    d, x, wt, b, dy, xd, dyd, _ = make(shape, _C.BF16)
    s = shape[6]
    assert lib.rcgan_conv_uses_tensor_cores(d, 2) == 1
    wr = wt.double().requires_grad_(True)
    O.conv2d(x.double(), wr, s).backward(dy.double())
    nb = lib.rcgan_conv2d_wgrad_workspace(d)
    ws = torch.zeros(max(nb, 4), dtype=torch.uint8, device='cuda')
    dw = torch.full(wt.shape, 3.0, device='cuda')
    call('rcgan_conv2d_wgrad', d, xd.data_ptr(), dyd.data_ptr(), dw.data_ptr(), 0, ws.data_ptr(), nb, st())
    torch.cuda.synchronize()
    assert relerr(dw, wr.grad) < 2e-5
    call('rcgan_conv2d_wgrad', d, xd.data_ptr(), dyd.data_ptr(), dw.data_ptr(), 1, ws.data_ptr(), nb, st())
    assert relerr(dw, 2 * wr.grad) < 2e-5


@pytest.mark.parametrize('dtype', [_C.F32, _C.BF16])
@pytest.mark.parametrize('shape', [(3, 32, 32, 256, 3, 3, 1, 0, 0), (3, 32, 32, 3, 128, 3, 1, 0, 0), (4, 16, 16, 3, 128, 1, 1, 0, 0),
                                   (5, 28, 28, 1, 138, 5, 2, 0, 6), (6, 28, 28, 1, 64, 5, 2, 0, 0)])
def test_narrow_wgrad(lib, shape, dtype):
    """wgrad with <= 4 channels on one side (G.Output, D.Block.1.*, g_h3, d_h0_conv): the CUDA-core split-K path
    (a register-accumulating per-channel variant was measured 2-5x slower: index math per tap dominates)"""
    d, x, wt, b, dy, xd, dyd, _ = make(shape, dtype)
    wr = wt.double().requires_grad_(True)
    O.conv2d(x.double(), wr, shape[6]).backward(dy.double())
    nb = lib.rcgan_conv2d_wgrad_workspace(d)
    ws = torch.zeros(max(nb, 4), dtype=torch.uint8, device='cuda')
    dw = torch.full(wt.shape, 3.0, device='cuda')
    call('rcgan_conv2d_wgrad', d, xd.data_ptr(), dyd.data_ptr(), dw.data_ptr(), 0, ws.data_ptr(), nb, st())
    assert relerr(dw, wr.grad) < 2e-5
    call('rcgan_conv2d_wgrad', d, xd.data_ptr(), dyd.data_ptr(), dw.data_ptr(), 1, ws.data_ptr(), nb, st())
    assert relerr(dw, 2 * wr.grad) < 2e-5


@pytest.mark.parametrize('shape', [(2, 32, 32, 256, 3, 3, 1, 0, 5), (3, 8, 8, 64, 2, 3, 1, 0, 6), (2, 16, 16, 128, 4, 5, 1, 0, 4)])
def test_few_output_channel_backward_as_transposed_patch_gemm(lib, shape):
    """G.Output (256 -> 3, gan_resnet.py:405-407): dL/dx and dL/dw through rcgan_wflip + patch matrix of dL/dy + the
    tensor-core GEMMs, the call sequence ConvOp._backward_transposed issues."""
    dtype = _C.BF16
    d, x, wt, b, dy, xd, dyd, (ho, wo, ldx, ldy) = make(shape, dtype)
    n, h, w, cin, cout, k, s = shape[:7]
    xr, wr = x.double().requires_grad_(True), wt.double().requires_grad_(True)
    O.conv2d(xr, wr, s).backward(dy.double())
    td = ConvDesc(n, ho, wo, cout, h, w, cin, k, k, 1, k - 1 - d.pad_t, k - 1 - d.pad_l, ldy, ldx, dtype)
    kp = k * k * cout
    ldp = (kp + 7) // 8 * 8
    patch = torch.full((n * h * w, ldp), 9.0, device='cuda', dtype=torch.bfloat16)
    g = ConvDesc(n * h * w, 1, 1, kp, 1, 1, cin, 1, 1, 1, 0, 0, ldp, ldx, dtype)
    assert lib.rcgan_conv_uses_tensor_cores(g, 0) and lib.rcgan_conv_uses_tensor_cores(g, 2)
    wdev = dev(wt)
    call('rcgan_im2col', td, dyd.data_ptr(), patch.data_ptr(), ldp, st())
    wflip = torch.zeros(kp * cin, device='cuda')
    call('rcgan_wflip', wdev.data_ptr(), wflip.data_ptr(), k, k, cin, cout, 0, st())
    ref_flip = wt.flip(0, 1).permute(0, 1, 3, 2).reshape(-1)
    assert float((wflip.cpu() - ref_flip).abs().max()) == 0.0
    pack = torch.zeros(lib.rcgan_conv_wpack_bytes(g), dtype=torch.uint8, device='cuda')
    call('rcgan_conv_wpack', g, wflip.data_ptr(), None, pack.data_ptr(), st())
    dx = torch.zeros(n, h, w, ldx, device='cuda', dtype=torch.bfloat16)
    call('rcgan_conv2d_fprop', g, patch.data_ptr(), wflip.data_ptr(), pack.data_ptr(), None, dx.data_ptr(), dtype,
         _C.ACT_NONE, 0.0, st())
    assert relerr(dx[..., :cin].float(), xr.grad) < TOL[dtype]
    nb = lib.rcgan_conv2d_wgrad_workspace(g)
    ws = torch.zeros(max(nb, 4), dtype=torch.uint8, device='cuda')
    dwflip = torch.full((kp * cin,), 3.0, device='cuda')
    call('rcgan_conv2d_wgrad', g, patch.data_ptr(), xd.data_ptr(), dwflip.data_ptr(), 0, ws.data_ptr(), nb, st())
    dw = torch.full(wt.shape, 1.0, device='cuda')
    call('rcgan_wflip', dwflip.data_ptr(), dw.data_ptr(), k, k, cout, cin, 1, st())
    assert relerr(dw - 1.0, wr.grad) < TOL[dtype]


@pytest.mark.parametrize('shape', [SHAPES[6], SHAPES[0], (2, 32, 32, 3, 128, 3, 1, 5, 0), SHAPES[8], (2, 9, 9, 2, 64, 5, 2, 2, 0)])
def test_few_input_channel_dgrad_as_gemm_plus_col2im(lib, shape):
    """g_h3's conv2d_transpose forward / d_h0_conv's input gradient: T = dy * W^T on the tensor cores, then rcgan_col2im
    (the call sequence of nnops.ScatterDgrad), against the oracle's conv2d_transpose and autograd."""
    dtype = _C.BF16
    d, x, wt, b, dy, xd, dyd, (ho, wo, ldx, ldy) = make(shape, dtype)
    n, h, w, cin, cout, k, s = shape[:7]
    xr = x.double().requires_grad_(True)
    O.conv2d(xr, wt.double(), s).backward(dy.double())
    kp = k * k * cin
    ldt = (kp + 7) // 8 * 8
    rows = n * ho * wo
    g = ConvDesc(rows, 1, 1, cout, 1, 1, kp, 1, 1, 1, 0, 0, ldy, ldt, dtype)
    assert lib.rcgan_conv_uses_tensor_cores(g, 0)
    wdev = dev(wt)
    wT = torch.zeros(kp * cout, device='cuda')
    call('rcgan_wflip', wdev.data_ptr(), wT.data_ptr(), 1, 1, kp, cout, 0, st())
    assert float((wT.cpu().reshape(cout, kp) - wt.reshape(kp, cout).t()).abs().max()) == 0.0
    pack = torch.zeros(lib.rcgan_conv_wpack_bytes(g), dtype=torch.uint8, device='cuda')
    call('rcgan_conv_wpack', g, wT.data_ptr(), None, pack.data_ptr(), st())
    T = torch.zeros(rows * ldt, device='cuda')
    call('rcgan_conv2d_fprop', g, dyd.data_ptr(), wT.data_ptr(), pack.data_ptr(), None, T.data_ptr(), _C.F32, _C.ACT_NONE, 0.0, st())
    dx = torch.full((n, h, w, ldx), 7.0, device='cuda', dtype=torch.bfloat16)
    call('rcgan_col2im', d, T.data_ptr(), ldt, None, dx.data_ptr(), dtype, _C.ACT_NONE, 0.0, 0, st())
    assert relerr(dx[..., :cin].float(), xr.grad) < TOL[dtype]
    if ldx > cin:
        assert float((dx[..., cin:].float() - 7.0).abs().max()) == 0.0
    call('rcgan_col2im', d, T.data_ptr(), ldt, None, dx.data_ptr(), dtype, _C.ACT_NONE, 0.0, 1, st())
    assert relerr(dx[..., :cin].float(), 2 * xr.grad) < 2 * TOL[dtype]
    # as a deconv forward with bias + sigmoid into an fp32 image
    bias = torch.randn(cin, generator=torch.Generator().manual_seed(5))
    out = torch.zeros(n, h, w, ldx, device='cuda')
    call('rcgan_col2im', d, T.data_ptr(), ldt, keep(dev(bias)), out.data_ptr(), _C.F32, _C.ACT_SIGMOID, 0.0, 0, st())
    ref = torch.sigmoid(O.conv2d_transpose(dy.double(), wt.double(), (h, w), s) + bias.double())
    assert relerr(out[..., :cin], ref) < TOL[dtype]


@pytest.mark.parametrize('shape', [(2, 32, 32, 256, 3, 3, 1, 0, 5), (3, 8, 8, 64, 2, 3, 1, 0, 6), (2, 16, 16, 128, 4, 5, 1, 0, 4)])
def test_few_output_channel_forward_as_gemm_plus_col2im(lib, shape):
    """G.Output forward (256 -> 3, tanh): T = x * W2 on the tensor cores + rcgan_col2im over the flipped taps, the call
    sequence of ConvOp._forward_scatter, against the oracle conv."""
    dtype = _C.BF16
    d, x, wt, b, dy, xd, dyd, (ho, wo, ldx, ldy) = make(shape, dtype)
    n, h, w, cin, cout, k, s = shape[:7]
    kp = k * k * cout
    ldt = (kp + 7) // 8 * 8
    g = ConvDesc(n * h * w, 1, 1, cin, 1, 1, kp, 1, 1, 1, 0, 0, ldx, ldt, dtype)
    cd = ConvDesc(n, ho, wo, cout, h, w, cin, k, k, 1, k - 1 - d.pad_t, k - 1 - d.pad_l, ldy, ldx, dtype)
    wdev = dev(wt)
    w1 = torch.zeros(kp * cin, device='cuda'); w2 = torch.zeros(kp * cin, device='cuda')
    call('rcgan_wflip', wdev.data_ptr(), w1.data_ptr(), k, k, cin, cout, 0, st())
    call('rcgan_wflip', w1.data_ptr(), w2.data_ptr(), 1, 1, kp, cin, 0, st())
    pack = torch.zeros(lib.rcgan_conv_wpack_bytes(g), dtype=torch.uint8, device='cuda')
    call('rcgan_conv_wpack', g, w2.data_ptr(), None, pack.data_ptr(), st())
    T = torch.zeros(n * h * w * ldt, device='cuda')
    call('rcgan_conv2d_fprop', g, xd.data_ptr(), w2.data_ptr(), pack.data_ptr(), None, T.data_ptr(), _C.F32, _C.ACT_NONE, 0.0, st())
    y = torch.full((n, ho, wo, ldy), 7.0, device='cuda', dtype=torch.bfloat16)
    call('rcgan_col2im', cd, T.data_ptr(), ldt, keep(dev(b)), y.data_ptr(), dtype, _C.ACT_TANH, 0.0, 0, st())
    ref = torch.tanh(O.conv2d(x.double(), wt.double(), s) + b.double())
    assert relerr(y[..., :cout].float(), ref) < TOL[dtype]
    assert float((y[..., cout:].float() - 7.0).abs().max()) == 0.0


@pytest.mark.parametrize('dtype', [_C.F32, _C.BF16])
@pytest.mark.parametrize('shape', [SHAPES[7], (2, 32, 32, 256, 256, 3, 1, 0, 0), (600, 8, 8, 128, 128, 3, 1, 0, 0)])
def test_fprop_with_fused_residual(lib, shape, dtype):
    """rcgan_conv2d_fprop_res == conv + bias + residual (ResidualBlock's shortcut add, gan_resnet.py:328), bit-identical
    to the unfused fprop + add sequence.  (All three shapes are below the persistent kernel's tile threshold and run the
    one-tile kernel; the persistent variants are covered by tests/test_gpu_conv_persist.py.)"""
    d, x, wt, b, dy, xd, dyd, (ho, wo, ldx, ldy) = make(shape, dtype)
    n, cout = shape[0], shape[4]
    res = torch.randn(n, ho, wo, cout, generator=torch.Generator().manual_seed(9)).to(TD[dtype])
    resd = res.cuda()
    wdev, bdev = dev(wt), dev(b)
    pack = None
    if dtype == _C.BF16:
        pack = torch.zeros(lib.rcgan_conv_wpack_bytes(d), dtype=torch.uint8, device='cuda')
        call('rcgan_conv_wpack', d, wdev.data_ptr(), None, pack.data_ptr(), st())
    pp_ = pack.data_ptr() if pack is not None else None
    y = torch.zeros(n, ho, wo, cout, device='cuda', dtype=TD[dtype])
    call('rcgan_conv2d_fprop_res', d, xd.data_ptr(), wdev.data_ptr(), pp_, bdev.data_ptr(), resd.data_ptr(), y.data_ptr(), dtype,
         _C.ACT_NONE, 0.0, st())
    ref = O.conv2d(x.double(), wt.double(), shape[6]) + b.double() + res.double()
    assert relerr(y.float(), ref) < TOL[dtype]
    y2 = torch.zeros_like(y)
    call('rcgan_conv2d_fprop', d, xd.data_ptr(), wdev.data_ptr(), pp_, bdev.data_ptr(), y2.data_ptr(), dtype, _C.ACT_NONE, 0.0, st())
    call('rcgan_add', y2.data_ptr(), resd.data_ptr(), y2.data_ptr(), y2.numel(), dtype, st())
    assert torch.equal(y, y2)


@pytest.mark.parametrize('n,h,w,cin,cout', [(2, 16, 16, 64, 64), (3, 8, 8, 256, 128), (2, 32, 32, 128, 256), (40, 32, 32, 64, 160)])
def test_upsample_conv_folded_forward(lib, n, h, w, cin, cout):
    """rcgan_upconv2d_fprop (4 parity classes x 2x2 pre-summed taps on the small input) == 3x3 SAME conv of the 2x
    nearest-neighbour upsampled input (UpsampleConv, gan_resnet.py:259-272), bias + relu fused."""
    g = torch.Generator().manual_seed(n + h)
    xs = torch.randn(n, h // 2, w // 2, cin, generator=g).bfloat16()
    wt = torch.randn(3, 3, cin, cout, generator=g) * 0.05
    b = torch.randn(cout, generator=g)
    d = ConvDesc(n, h, w, cin, h, w, cout, 3, 3, 1, 1, 1, cin, cout, _C.BF16)
    nb = lib.rcgan_upconv2d_pack_bytes(d)
    assert nb > 0
    wdev, bdev, xd = dev(wt), dev(b), xs.cuda()
    wf = torch.zeros(16 * cin * cout, device='cuda')
    pack = torch.zeros(nb, dtype=torch.uint8, device='cuda')
    call('rcgan_upconv2d_fold', d, wdev.data_ptr(), wf.data_ptr(), pack.data_ptr(), st())
    y = torch.zeros(n, h, w, cout, device='cuda', dtype=torch.bfloat16)
    call('rcgan_upconv2d_fprop', d, xd.data_ptr(), pack.data_ptr(), bdev.data_ptr(), y.data_ptr(), _C.BF16, _C.ACT_RELU, 0.0, st())
    up = xs.double().repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
    ref = torch.relu(O.conv2d(up, wt.double(), 1) + b.double())
    assert relerr(y.float(), ref) < TOL[_C.BF16]
    # unsupported shapes are refused loudly, not silently mis-computed
    bad = ConvDesc(n, h, w, cin, h, w, cout, 5, 5, 1, 2, 2, cin, cout, _C.BF16)
    assert lib.rcgan_upconv2d_pack_bytes(bad) == 0
    with pytest.raises(_C.RcganError):
        call('rcgan_upconv2d_fprop', bad, xd.data_ptr(), pack.data_ptr(), None, y.data_ptr(), _C.BF16, _C.ACT_NONE, 0.0, st())
