"""The persistent tcgen05 conv variants -- the kernels the benchmarks actually run -- against the oracle at sizes that
cross their dispatch thresholds (conv_tc.cu run_tc / run_tc_persist: persistent at >= 3*148 big tiles, 256-row tiles at
>= 2*148), and, forced with RCGAN_TC_PERSIST=2, on the small ragged shapes.  Every case asserts the variant that ran
(rcgan_last_conv_variant), so a change of the dispatch rules cannot silently move these tests onto another kernel.

Reference: oracle/nn.py conv2d / conv2d_transpose in fp32 on bf16-exact operands over the whole tensor (4e-7 from fp64 at
these K) and in fp64 on a sample of images that straddle tile boundaries."""
import pytest
import torch

from oracle import nn as O
from robust_conditional_gan_b200 import _C
from robust_conditional_gan_b200._C import ConvDesc, call
from robust_conditional_gan_b200.graph import same_pad
from util import TD, dev, keep, relerr, st

pytestmark = pytest.mark.gpu
TOL_OUT = {_C.BF16: 6e-3, _C.F32: 2e-5}
NAME = {_C.BF16: 'bf16', _C.F32: 'f32'}

# Every case runs twice: with RCGAN_TC_PAIR=2 on the CTA-pair kernel (cta_group::2) wherever the N tile is a multiple of 32, and
# with RCGAN_TC_PAIR=0 on the one-CTA persistent kernel.  (The default, 1, takes pairs for the 256-wide bf16 tiles only; the
# full-size step tests run that mix.)  The variant strings in the tables name the one-CTA kernel;
# expect() maps them to the pair kernel's instantiation (conv_tc.cu run_tc_persist).
PAIR_OF = {'conv_tc_persist<256,1,3,bf16': 'conv_tc_pair<256,1,4,bf16', 'conv_tc_persist<128,2,3,bf16': 'conv_tc_pair<128,2,4,bf16',
           'conv_tc_persist<128,2,3,f32': 'conv_tc_pair<128,2,3,f32', 'conv_tc_persist<128,1,5,bf16': 'conv_tc_pair<128,1,6,bf16',
           'conv_tc_persist<128,1,4,f32': 'conv_tc_pair<128,1,5,f32'}


@pytest.fixture(autouse=True, params=['pair', 'solo'])
def pair(request, monkeypatch):
    monkeypatch.setenv('RCGAN_TC_PAIR', '2' if request.param == 'pair' else '0')
    return request.param == 'pair'


def expect(variant, ncols, pair, odt=_C.BF16):
    """the variant string of the kernel that must have run: `variant` itself, or its CTA-pair counterpart"""
    if variant is None or not pair or not variant.startswith('conv_tc_persist<'):
        return variant
    cap = 256 if (ncols > 128 and odt == _C.BF16) else 128
    bn_eff = cap if ncols >= cap else (ncols + 15) // 16 * 16
    if bn_eff % 32:
        return variant
    head, tail = variant.rsplit(',', 1)
    return PAIR_OF[head] + ',' + tail


def persistent(v):
    return v.startswith('conv_tc_persist<') or v.startswith('conv_tc_pair<')


def make(shape, seed=0, wscale=0.05):
    n, h, w, cin, cout, k, s, px, py = shape
    g = torch.Generator().manual_seed(seed)
    ho, pt = same_pad(h, k, s)
    wo, pl = same_pad(w, k, s)
    ldx, ldy = cin + px, cout + py
    x = torch.randn(n, h, w, cin, generator=g).bfloat16().float()
    wt = (torch.randn(k, k, cin, cout, generator=g) * wscale).bfloat16().float()      # bf16-exact: the pack rounds nothing
    b = torch.randn(cout, generator=g)
    dy = torch.randn(n, ho, wo, cout, generator=g).bfloat16().float()
    xbuf = torch.zeros(n, h, w, ldx); xbuf[..., :cin] = x
    dybuf = torch.zeros(n, ho, wo, ldy); dybuf[..., :cout] = dy
    d = ConvDesc(n, h, w, cin, ho, wo, cout, k, k, s, pt, pl, ldx, ldy, _C.BF16)
    return d, x, wt, b, dy, dev(xbuf, torch.bfloat16), dev(dybuf, torch.bfloat16), (ho, wo, ldx, ldy)


def pack_for(lib, d, wt):
    nb = lib.rcgan_conv_wpack_bytes(d)
    assert nb > 0
    pack = torch.zeros(nb, dtype=torch.uint8, device='cuda')
    wd = dev(wt)
    call('rcgan_conv_wpack', d, wd.data_ptr(), None, pack.data_ptr(), st())
    return pack, wd


def sample_images(n):
    return sorted(set([0, 1, n // 3, n // 2, n - 2, n - 1]) & set(range(n)))


def check(got, ref32, ref64_fn, n, tol):
    """whole tensor vs the fp32 reference, sampled images vs fp64"""
    assert relerr(got, ref32) < tol
    idx = sample_images(n)
    assert relerr(got[idx], ref64_fn(idx)) < tol


# (shape, output dtype, expected variant): the instantiations of profiles/r1c_*_iteration_summary.csv
FPROP = [
    ((64, 32, 32, 256, 256, 3, 1, 0, 0), _C.BF16, 'conv_tc_persist<256,1,3,bf16,multi=0>'),   # CIFAR G 3x3, 15 % of the iteration
    ((64, 32, 32, 256, 256, 3, 1, 0, 0), _C.F32, 'conv_tc_persist<128,2,3,f32,multi=0>'),     # same layer feeding a norm in fp32
    ((512, 16, 16, 128, 128, 3, 1, 0, 0), _C.BF16, 'conv_tc_persist<128,2,3,bf16,multi=0>'),  # CIFAR D 3x3 at [real; fake]
    ((512, 16, 16, 128, 128, 3, 1, 0, 0), _C.F32, 'conv_tc_persist<128,2,3,f32,multi=0>'),
    ((1000, 12, 12, 128, 128, 3, 1, 0, 0), _C.BF16, 'conv_tc_persist<128,2,3,bf16,multi=0>'),  # ragged M: the last tile has ONE sub-tile
    ((3000, 14, 14, 64, 64, 5, 2, 0, 0), _C.BF16, 'conv_tc_persist<128,2,3,bf16,multi=0>'),     # 5x5 s2, N = 64, ragged M (56 rows)
    ((480, 16, 16, 256, 138, 1, 1, 0, 6), _C.BF16, 'conv_tc_persist<256,1,3,bf16,multi=0>'),  # N = 138 -> one 144-wide tile, ld 144
]


@pytest.mark.parametrize('shape,odt,variant', FPROP)
def test_persistent_fprop(lib, shape, odt, variant, pair):
    d, x, wt, b, dy, xd, dyd, (ho, wo, ldx, ldy) = make(shape)
    n, cout, s = shape[0], shape[4], shape[6]
    pack, wd = pack_for(lib, d, wt)
    bd = dev(b)
    y = torch.full((n, ho, wo, ldy), 7.0, device='cuda', dtype=TD[odt])
    call('rcgan_conv2d_fprop', d, xd.data_ptr(), wd.data_ptr(), pack.data_ptr(), bd.data_ptr(), y.data_ptr(), odt, _C.ACT_LRELU, 0.2, st())
    torch.cuda.synchronize()
    assert _C.last_conv_variant() == expect(variant, cout, pair, odt)
    ref32 = O.lrelu(O.conv2d(x, wt, s) + b)
    check(y[..., :cout].float().cpu(), ref32, lambda i: O.lrelu(O.conv2d(x[i].double(), wt.double(), s) + b.double()), n, TOL_OUT[odt])
    if ldy > cout:
        assert float((y[..., cout:].float() - 7.0).abs().max()) == 0.0      # channel padding untouched


DGRAD = [
    ((64, 32, 32, 256, 256, 3, 1, 0, 0), _C.BF16, 'conv_tc_persist<256,1,3,bf16,multi=0>'),
    ((512, 16, 16, 128, 128, 3, 1, 0, 0), _C.BF16, 'conv_tc_persist<128,2,3,bf16,multi=0>'),
    ((512, 16, 16, 128, 128, 3, 1, 0, 0), _C.F32, 'conv_tc_persist<128,2,3,f32,multi=0>'),
    # g_h2 (conv2d_transpose 7x7x138 -> 14x14x128, k5 s2) at batch 1024: the MNIST roofline kernel, 4 parity classes in one launch
    ((1024, 14, 14, 128, 138, 5, 2, 0, 6), _C.BF16, 'conv_tc_persist<128,2,3,bf16,multi=1>'),
    ((1021, 14, 14, 128, 138, 5, 2, 0, 6), _C.F32, 'conv_tc_persist<128,2,3,f32,multi=1>'),     # ragged: the last tile of every class has one sub-tile
    ((700, 14, 14, 256, 64, 5, 2, 0, 0), _C.BF16, 'conv_tc_persist<256,1,3,bf16,multi=1>'),     # N = cin = 256, multi
    ((2048, 14, 14, 64, 64, 5, 2, 0, 0), _C.BF16, 'conv_tc_persist<128,2,3,bf16,multi=1>'),     # MNIST d_h1's input gradient at 2B
]


@pytest.mark.parametrize('shape,odt,variant', DGRAD)
def test_persistent_dgrad_and_deconv(lib, shape, odt, variant, pair):
    d, x, wt, b, dy, xd, dyd, (ho, wo, ldx, ldy) = make(shape)
    n, h, w, cin, cout, k, s = shape[:7]
    pack, wd = pack_for(lib, d, wt)
    xr = x.clone().requires_grad_(True)
    O.conv2d(xr, wt, s).backward(dy)
    g32 = xr.grad

    def g64(idx):
        xi = x[idx].double().requires_grad_(True)
        O.conv2d(xi, wt.double(), s).backward(dy[idx].double())
        return xi.grad
    dx = torch.zeros(n, h, w, ldx, device='cuda', dtype=TD[odt])
    call('rcgan_conv2d_dgrad', d, dyd.data_ptr(), wd.data_ptr(), pack.data_ptr(), None, dx.data_ptr(), odt, _C.ACT_NONE, 0.0, 0, st())
    torch.cuda.synchronize()
    variant = expect(variant, cin, pair, odt)
    assert _C.last_conv_variant() == variant
    check(dx[..., :cin].float().cpu(), g32, g64, n, TOL_OUT[odt])
    # accumulate=1: += onto a buffer that already holds a gradient (bf16: one more rounding)
    call('rcgan_conv2d_dgrad', d, dyd.data_ptr(), wd.data_ptr(), pack.data_ptr(), None, dx.data_ptr(), odt, _C.ACT_NONE, 0.0, 1, st())
    assert relerr(dx[..., :cin].float().cpu(), 2 * g32) < 2 * TOL_OUT[odt]
    # the same launch as the forward of deconv2d (bias + sigmoid epilogue)
    if s == 2:
        bias = torch.randn(cin, generator=torch.Generator().manual_seed(5))
        out = torch.zeros(n, h, w, ldx, device='cuda', dtype=TD[odt])
        call('rcgan_conv2d_dgrad', d, dyd.data_ptr(), wd.data_ptr(), pack.data_ptr(), keep(dev(bias)), out.data_ptr(), odt,
             _C.ACT_SIGMOID, 0.0, 0, st())
        assert _C.last_conv_variant() == variant
        ref = torch.sigmoid(O.conv2d_transpose(dy, wt, (h, w), s) + bias)
        assert relerr(out[..., :cin].float().cpu(), ref) < TOL_OUT[odt]


# small and ragged shapes pushed through the persistent kernel (RCGAN_TC_PERSIST=2): narrow tiles (N = 32, 74 -> 80, 138 -> 144),
# M far below one tile per SM, K tails (cin = 110 -> two K blocks, the second half empty)
FORCED = [
    (4, 14, 14, 64, 64, 5, 2, 0, 0), (3, 14, 14, 128, 138, 5, 2, 0, 6), (5, 28, 28, 64, 74, 5, 2, 0, 6), (4, 32, 32, 128, 128, 3, 1, 0, 0),
    (2, 16, 16, 256, 256, 3, 1, 0, 0), (2, 8, 8, 1024, 256, 1, 1, 0, 0), (300, 1, 1, 110, 1024, 1, 1, 2, 0), (130, 1, 1, 64, 32, 1, 1, 0, 0),
    (70, 1, 1, 1034, 896, 1, 1, 6, 0),
]


@pytest.mark.parametrize('shape', FORCED)
def test_forced_persistent_small_shapes(lib, shape, monkeypatch):
    monkeypatch.setenv('RCGAN_TC_PERSIST', '2')
    d, x, wt, b, dy, xd, dyd, (ho, wo, ldx, ldy) = make(shape, wscale=0.1)
    n, h, w, cin, cout, k, s = shape[:7]
    pack, wd = pack_for(lib, d, wt)
    bd = dev(b)
    ref = O.lrelu(O.conv2d(x.double(), wt.double(), s) + b.double())
    for odt in (_C.BF16, _C.F32):
        y = torch.full((n, ho, wo, ldy), 7.0, device='cuda', dtype=TD[odt])
        call('rcgan_conv2d_fprop', d, xd.data_ptr(), wd.data_ptr(), pack.data_ptr(), bd.data_ptr(), y.data_ptr(), odt, _C.ACT_LRELU, 0.2, st())
        torch.cuda.synchronize()
        assert persistent(_C.last_conv_variant()), _C.last_conv_variant()
        assert relerr(y[..., :cout].float(), ref) < TOL_OUT[odt]
        if ldy > cout:
            assert float((y[..., cout:].float() - 7.0).abs().max()) == 0.0
    xr = x.double().requires_grad_(True)
    O.conv2d(xr, wt.double(), s).backward(dy.double())
    dx = torch.zeros(n, h, w, ldx, device='cuda', dtype=torch.float32)
    for acc in (0, 1):
        call('rcgan_conv2d_dgrad', d, dyd.data_ptr(), wd.data_ptr(), pack.data_ptr(), None, dx.data_ptr(), _C.F32, _C.ACT_NONE, 0.0, acc, st())
        torch.cuda.synchronize()
        assert persistent(_C.last_conv_variant()), _C.last_conv_variant()
        assert relerr(dx[..., :cin], (1 + acc) * xr.grad) < 2e-5


def test_small_shapes_take_the_one_tile_kernel(lib):
    """below the thresholds the one-tile-per-CTA kernel runs (what tests/test_gpu_conv.py's TC_SHAPES cover)"""
    shape = (4, 32, 32, 128, 128, 3, 1, 0, 0)
    d, x, wt, b, dy, xd, dyd, (ho, wo, ldx, ldy) = make(shape)
    pack, wd = pack_for(lib, d, wt)
    y = torch.zeros(4, ho, wo, ldy, device='cuda', dtype=torch.bfloat16)
    call('rcgan_conv2d_fprop', d, xd.data_ptr(), wd.data_ptr(), pack.data_ptr(), None, y.data_ptr(), _C.BF16, _C.ACT_NONE, 0.0, st())
    assert _C.last_conv_variant() == 'conv_tc<128,3,im2col=1>'


@pytest.mark.parametrize('shape,variant', [((600, 8, 8, 128, 128, 3, 1, 0, 0), 'conv_tc<128,3,im2col=1>'),
                                           ((1800, 8, 8, 128, 128, 3, 1, 0, 0), 'conv_tc_persist<128,2,3,bf16,multi=0>'),
                                           ((64, 32, 32, 256, 256, 3, 1, 0, 0), 'conv_tc_persist<256,1,3,bf16,multi=0>')])
def test_fused_residual_in_both_kernels(lib, shape, variant, pair):
    """rcgan_conv2d_fprop_res (ResidualBlock's shortcut add, gan_resnet.py:328) is bit-identical to fprop + rcgan_add in the
    one-tile kernel and in both persistent tile shapes"""
    d, x, wt, b, dy, xd, dyd, (ho, wo, ldx, ldy) = make(shape)
    n, cout = shape[0], shape[4]
    res = torch.randn(n, ho, wo, cout, generator=torch.Generator().manual_seed(9)).bfloat16()
    resd = res.cuda()
    pack, wd = pack_for(lib, d, wt)
    bd = dev(b)
    y = torch.zeros(n, ho, wo, cout, device='cuda', dtype=torch.bfloat16)
    call('rcgan_conv2d_fprop_res', d, xd.data_ptr(), wd.data_ptr(), pack.data_ptr(), bd.data_ptr(), resd.data_ptr(), y.data_ptr(), _C.BF16,
         _C.ACT_NONE, 0.0, st())
    assert _C.last_conv_variant() == expect(variant, cout, pair)
    ref = O.conv2d(x, wt, shape[6]) + b + res.float()
    assert relerr(y.float().cpu(), ref) < 6e-3
    y2 = torch.zeros_like(y)
    call('rcgan_conv2d_fprop', d, xd.data_ptr(), wd.data_ptr(), pack.data_ptr(), bd.data_ptr(), y2.data_ptr(), _C.BF16, _C.ACT_NONE, 0.0, st())
    call('rcgan_add', y2.data_ptr(), resd.data_ptr(), y2.data_ptr(), y2.numel(), _C.BF16, st())
    assert torch.equal(y, y2)


@pytest.mark.parametrize('n,h,w,cin,cout,variant', [(64, 32, 32, 256, 256, 'conv_tc_persist<256,1,3,bf16,multi=1>'),
                                                    (130, 16, 16, 256, 128, 'conv_tc_persist<128,1,5,bf16,multi=1>'),
                                                    (300, 32, 32, 64, 96, 'conv_tc_persist<128,2,3,bf16,multi=1>')])
def test_persistent_upsample_conv(lib, n, h, w, cin, cout, variant, pair):
    """rcgan_upconv2d_fprop at the benchmark's G.Block sizes: 4 parity problems in one persistent launch vs the oracle's
    3x3 conv of the nearest-neighbour upsampled input (gan_resnet.py:259-272)"""
    g = torch.Generator().manual_seed(n + h)
    xs = torch.randn(n, h // 2, w // 2, cin, generator=g).bfloat16()
    wt = torch.randn(3, 3, cin, cout, generator=g) * 0.05
    b = torch.randn(cout, generator=g)
    d = ConvDesc(n, h, w, cin, h, w, cout, 3, 3, 1, 1, 1, cin, cout, _C.BF16)
    nb = lib.rcgan_upconv2d_pack_bytes(d)
    wdev, bdev, xd = dev(wt), dev(b), xs.cuda()
    wf = torch.zeros(16 * cin * cout, device='cuda')
    pack = torch.zeros(nb, dtype=torch.uint8, device='cuda')
    call('rcgan_upconv2d_fold', d, wdev.data_ptr(), wf.data_ptr(), pack.data_ptr(), st())
    y = torch.zeros(n, h, w, cout, device='cuda', dtype=torch.bfloat16)
    call('rcgan_upconv2d_fprop', d, xd.data_ptr(), pack.data_ptr(), bdev.data_ptr(), y.data_ptr(), _C.BF16, _C.ACT_RELU, 0.0, st())
    torch.cuda.synchronize()
    assert _C.last_conv_variant() == expect(variant, cout, pair)
    up = xs.float().repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
    ref = torch.relu(O.conv2d(up, wt, 1) + b)
    assert relerr(y.float().cpu(), ref) < 1e-2            # the folded taps are pre-summed in fp32, then rounded to bf16 once more


@pytest.mark.parametrize('shape,variant', [((128, 32, 32, 256, 256, 3, 1, 0, 0), 'wgrad_tc<128,im2col=1>'),
                                           ((512, 16, 16, 128, 128, 3, 1, 0, 0), 'wgrad_tc<128,im2col=1>'),     # 18 units: the last pair's peer CTA is empty
                                           ((300, 32, 32, 128, 128, 4, 2, 0, 0), 'wgrad_tc<128,im2col=1>'),     # folded ConvMeanPool, ragged pixel count
                                           ((1024, 14, 14, 128, 138, 5, 2, 0, 6), 'wgrad_tc<128,im2col=1>'),
                                           ((2048, 7, 7, 64, 64, 5, 2, 0, 0), 'wgrad_tc<64,im2col=1>')])
def test_wgrad_at_benchmark_sizes(lib, shape, variant, pair):
    """the tcgen05 wgrad at the benchmarks' contraction lengths (K = n*ho*wo up to 131072 pixels, many split-K slices)"""
    d, x, wt, b, dy, xd, dyd, _ = make(shape)
    s = shape[6]
    wr = wt.clone().requires_grad_(True)
    O.conv2d(x, wr, s).backward(dy)
    nb = lib.rcgan_conv2d_wgrad_workspace(d)
    ws = torch.zeros(max(nb, 4), dtype=torch.uint8, device='cuda')
    dw = torch.full(wt.shape, 3.0, device='cuda')
    call('rcgan_conv2d_wgrad', d, xd.data_ptr(), dyd.data_ptr(), dw.data_ptr(), 0, ws.data_ptr(), nb, st())
    torch.cuda.synchronize()
    if pair and shape[4] % 128 == 0:        # the CTA-pair wgrad: 4 units x 256 (cout % 256 == 0) or 128 channels
        variant = 'wgrad_tc_pair<%d>' % (256 if shape[4] % 256 == 0 else 128)
    assert _C.last_conv_variant() == variant
    assert relerr(dw.cpu(), wr.grad) < 8e-5               # fp32 reference over K ~ 1e5 pixels carries its own ~2e-5
    # fp64 on a slice of the output channels (wgrad cost is linear in cout)
    w64 = wt[..., :16].double().requires_grad_(True)
    O.conv2d(x.double(), w64, s).backward(dy[..., :16].double())
    # fp32 accumulation over K = n*ho*wo up to 131072 pixels: sqrt(K) * 2^-24 ~ 2e-5 is the arithmetic's own floor
    assert relerr(dw[..., :16].cpu(), w64.grad) < 6e-5
    call('rcgan_conv2d_wgrad', d, xd.data_ptr(), dyd.data_ptr(), dw.data_ptr(), 1, ws.data_ptr(), nb, st())
    assert relerr(dw[..., :16].cpu(), 2 * w64.grad) < 6e-5


# ------------------------------------------------------------------------------------------------ fused epilogues
EX_SHAPES = [((600, 8, 8, 128, 128, 3, 1, 0, 0), 'conv_tc<128,3,im2col=1>'),
             ((1800, 8, 8, 128, 128, 3, 1, 0, 0), 'conv_tc_persist<128,2,3,bf16,multi=0>'),
             ((64, 32, 32, 256, 256, 3, 1, 0, 0), 'conv_tc_persist<256,1,3,bf16,multi=0>'),
             ((512, 32, 32, 128, 128, 4, 2, 0, 0), 'conv_tc_persist<128,2,3,bf16,multi=1>'),     # the folded ConvMeanPool's dgrad
             ((37, 16, 16, 64, 72, 3, 1, 0, 0), None)]                      # N = 64 / 72: ragged pieces


@pytest.mark.parametrize('shape,variant', EX_SHAPES)
@pytest.mark.parametrize('act', [_C.ACT_RELU, _C.ACT_LRELU])
def test_dgrad_with_fused_activation_backward(lib, shape, variant, act, pair):
    """rcgan_conv2d_dgrad_ex(mask): dx (=|+=) act'(mask) * dgrad -- bit-identical to rcgan_conv2d_dgrad followed by rcgan_act_bwd
    (the `nonlinearity` in front of a conv, gan_resnet.py:318-325, differentiated in the producing dgrad's epilogue)"""
    d, x, wt, b, dy, xd, dyd, (ho, wo, ldx, ldy) = make(shape)
    n, h, w, cin = shape[:4]
    pack, wd = pack_for(lib, d, wt)
    g = torch.Generator().manual_seed(11)
    mask = torch.randn(n, h, w, ldx, generator=g).bfloat16().cuda()          # the activation's forward output (sign matters only)
    old = torch.randn(n, h, w, ldx, generator=g).bfloat16().cuda()
    tmp = torch.zeros(n, h, w, ldx, device='cuda', dtype=torch.bfloat16)
    call('rcgan_conv2d_dgrad', d, dyd.data_ptr(), wd.data_ptr(), pack.data_ptr(), None, tmp.data_ptr(), _C.BF16, _C.ACT_NONE, 0.0, 0, st())
    rows = n * h * w
    for acc in (0, 1):
        ref = old.clone()
        call('rcgan_act_bwd', tmp.data_ptr(), mask.data_ptr(), ref.data_ptr(), rows, cin, ldx, ldx, ldx, _C.BF16, act, 0.2, acc, st())
        got = old.clone()
        ep = _C.ConvEpilogue(mask=mask.data_ptr(), mask_act=act, mask_leak=0.2)
        import ctypes
        call('rcgan_conv2d_dgrad_ex', d, dyd.data_ptr(), pack.data_ptr(), None, got.data_ptr(), _C.BF16, _C.ACT_NONE, 0.0, acc,
             ctypes.byref(ep), st())
        torch.cuda.synchronize()
        assert variant is None or _C.last_conv_variant() == expect(variant, cin, pair)
        assert torch.equal(got[..., :cin], ref[..., :cin])
    # and against the oracle: relu'(mask) * conv2d gradient
    xr = x.clone().requires_grad_(True)
    O.conv2d(xr, wt, shape[6]).backward(dy)
    m = mask[..., :cin].float().cpu()
    slope = torch.where(m > 0, torch.ones_like(m), torch.full_like(m, 0.0 if act == _C.ACT_RELU else 0.2))
    call('rcgan_conv2d_dgrad_ex', d, dyd.data_ptr(), pack.data_ptr(), None, got.data_ptr(), _C.BF16, _C.ACT_NONE, 0.0, 0, ctypes.byref(ep), st())
    assert relerr(got[..., :cin].float().cpu(), slope * xr.grad) < 6e-3


@pytest.mark.parametrize('shape,variant', [s for s in EX_SHAPES if s[0][6] == 1])
@pytest.mark.parametrize('res_up', [0, 1])
def test_fprop_with_fused_residual_upsampling_and_second_output(lib, shape, variant, res_up, pair):
    """rcgan_conv2d_fprop_ex(res, res_up, out2): y = conv + bias + [upsample2](res), out2 = relu(y) -- ResidualBlock's
    `shortcut + output` with UpsampleConv_1x1's shortcut still at the small resolution (gan_resnet.py:259-272, 305-328) and the
    next block's `nonlinearity(inputs)`; bit-identical to fprop, upsample, add, relu as separate kernels"""
    import ctypes
    d, x, wt, b, dy, xd, dyd, (ho, wo, ldx, ldy) = make(shape)
    n, cout = shape[0], shape[4]
    if res_up and (ho % 2 or wo % 2):
        pytest.skip('odd output size')
    pack, wd = pack_for(lib, d, wt)
    bd = dev(b)
    g = torch.Generator().manual_seed(13)
    rs = (n, ho // 2, wo // 2, ldy) if res_up else (n, ho, wo, ldy)
    res = torch.randn(*rs, generator=g).bfloat16().cuda()
    y0 = torch.zeros(n, ho, wo, ldy, device='cuda', dtype=torch.bfloat16)
    call('rcgan_conv2d_fprop', d, xd.data_ptr(), wd.data_ptr(), pack.data_ptr(), bd.data_ptr(), y0.data_ptr(), _C.BF16, _C.ACT_NONE, 0.0, st())
    full = res.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2) if res_up else res
    ref = (y0.float() + full.float()).bfloat16()
    y = torch.full((n, ho, wo, ldy), 7.0, device='cuda', dtype=torch.bfloat16)
    y2 = torch.full((n, ho, wo, ldy), 7.0, device='cuda', dtype=torch.bfloat16)
    ep = _C.ConvEpilogue(res=res.data_ptr(), res_up=res_up, ld_res=ldy, out2=y2.data_ptr(), out2_act=_C.ACT_RELU)
    call('rcgan_conv2d_fprop_ex', d, xd.data_ptr(), pack.data_ptr(), bd.data_ptr(), y.data_ptr(), _C.BF16, _C.ACT_NONE, 0.0, ctypes.byref(ep), st())
    torch.cuda.synchronize()
    assert variant is None or _C.last_conv_variant() == expect(variant, cout, pair)
    assert torch.equal(y[..., :cout], ref[..., :cout])
    assert torch.equal(y2[..., :cout], torch.relu(ref[..., :cout]))
    if ldy > cout:
        assert float((y[..., cout:].float() - 7.0).abs().max()) == 0.0 and float((y2[..., cout:].float() - 7.0).abs().max()) == 0.0


@pytest.mark.parametrize('shape,kind', [((64, 32, 32, 256, 256, 3, 1, 0, 0), 'fprop'), ((5, 16, 16, 128, 128, 3, 1, 0, 0), 'fprop'),
                                        ((300, 8, 8, 64, 64, 3, 1, 0, 0), 'fprop'), ((130, 16, 16, 256, 256, 4, 2, 0, 0), 'deconv'),
                                        ((3, 8, 8, 128, 256, 4, 2, 0, 0), 'deconv')])
def test_epilogue_column_statistics(lib, shape, kind, pair):
    """rcgan_conv_epilogue.colstats: per-CTA partial sums / sums of squares / row counts of the STORED bf16 output, merged here on
    the host, equal the moments of the tensor (what the following conditional batch norm needs: normalization.py:38-41);
    rcgan_bn_fwd_prestats on them == rcgan_bn_fwd with its own statistics pass.  The persistent kernel is forced for small shapes."""
    import ctypes
    d, x, wt, b, dy, xd, dyd, (ho, wo, ldx, ldy) = make(shape)
    n, h, w, cin, cout = shape[:5]
    pack, wd = pack_for(lib, d, wt)
    N = cout if kind == 'fprop' else cin
    parts = torch.full((lib.rcgan_colstats_floats(N),), 7.0, device='cuda')
    ep = _C.ConvEpilogue(colstats=parts.data_ptr())
    if kind == 'fprop':
        y = torch.zeros(n, ho, wo, cout, device='cuda', dtype=torch.bfloat16)
        call('rcgan_conv2d_fprop_ex', d, xd.data_ptr(), pack.data_ptr(), keep(dev(b)), y.data_ptr(), _C.BF16, _C.ACT_NONE, 0.0,
             ctypes.byref(ep), st())
    else:       # the folded UpsampleConv: conv2d_transpose = dgrad of the 4x4 stride-2 conv, 4 parity problems in one launch
        y = torch.zeros(n, h, w, cin, device='cuda', dtype=torch.bfloat16)
        bias = torch.randn(cin, generator=torch.Generator().manual_seed(2))
        call('rcgan_conv2d_dgrad_ex', d, dyd.data_ptr(), pack.data_ptr(), keep(dev(bias)), y.data_ptr(), _C.BF16, _C.ACT_NONE, 0.0, 0,
             ctypes.byref(ep), st())
    torch.cuda.synchronize()
    assert _C.last_conv_variant().startswith('conv_tc_pair<' if pair else 'conv_tc_persist<')
    P = 148
    sums = parts[:P * 2 * N].reshape(P, 2, N).double().cpu()
    counts = parts[P * 2 * N:].double().cpu()
    live = counts > 0
    rows = y.numel() // N
    assert int(counts.sum()) == rows and float(counts.min()) >= 0
    yf = y.double().cpu().reshape(rows, N)
    assert relerr(sums[live, 0].sum(0), yf.sum(0)) < 1e-5
    assert relerr(sums[live, 1].sum(0), (yf * yf).sum(0)) < 1e-5
    # the batch norm fed with these partials == the batch norm that reads the tensor twice
    g = torch.Generator().manual_seed(4)
    scale, offset = dev(torch.rand(10, N, generator=g) + 0.5), dev(torch.randn(10, N, generator=g))
    spatial = y.shape[1] * y.shape[2]
    lab = torch.randint(0, 10, (n,), generator=g).to(torch.int32).cuda()
    nb = lib.rcgan_bn_workspace(n, spatial, N)
    ws = torch.zeros(max(nb, 256), dtype=torch.uint8, device='cuda')
    o1, o2 = torch.zeros_like(y), torch.zeros_like(y)
    s1, s2 = torch.zeros(2 * N, device='cuda'), torch.zeros(2 * N, device='cuda')
    call('rcgan_bn_fwd', y.data_ptr(), o1.data_ptr(), n, spatial, N, _C.BF16, _C.BF16, scale.data_ptr(), offset.data_ptr(), lab.data_ptr(),
         1e-5, _C.ACT_RELU, 0.0, 1, 0.9, None, None, s1.data_ptr(), ws.data_ptr(), nb, st())
    call('rcgan_bn_fwd_prestats', y.data_ptr(), o2.data_ptr(), n, spatial, N, _C.BF16, _C.BF16, scale.data_ptr(), offset.data_ptr(),
         lab.data_ptr(), 1e-5, _C.ACT_RELU, 0.0, 0.9, None, None, s2.data_ptr(), parts.data_ptr(), st())
    assert relerr(s2[:N], s1[:N]) < 1e-5 and relerr(s2[N:], s1[N:]) < 1e-5
    assert relerr(o2.float(), o1.float()) < 1e-3
