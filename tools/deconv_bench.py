"""Where the time of the folded UpsampleConv (conv2d_transpose 4x4 stride 2 = 4 parity problems in one persistent launch) and
of the fused-epilogue 3x3 conv goes: the launch as is, and with parts of the work skipped (RCGAN_TC_DBG bits of the one-CTA
persistent kernel: 1 no MMA, 2 no TMA loads, 16 no epilogue stores).  Debug helper, not a test.
usage: python tools/deconv_bench.py"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from robust_conditional_gan_b200 import _C
from robust_conditional_gan_b200._C import ConvDesc, call
from robust_conditional_gan_b200.graph import same_pad

lib = _C.load()
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')


def timeit(fn, reps=12):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2] * 1e3


def run(label, fn, flops, modes):
    for pair, dbg in modes:
        os.environ['RCGAN_TC_PAIR'] = str(pair)
        os.environ['RCGAN_TC_DBG'] = str(dbg)
        t = timeit(fn)
        print('%-44s pair=%d dbg=%-2d %-42s %8.1f us %8.1f TFLOP/s' % (label, pair, dbg, _C.last_conv_variant(), t, flops / t / 1e6), flush=True)
    os.environ['RCGAN_TC_DBG'] = '0'


MODES = [(0, 0), (0, 16), (0, 1), (0, 2), (0, 17), (0, 18), (2, 0)]
# the deconv of G's last block: [n, 16, 16, 256] -> [n, 32, 32, 256]; desc = the 4x4 stride-2 conv whose dgrad it is
for n in (512, 256):
    h = w = 32; cin = cout = 256; k = 4; s = 2
    ho, pt = same_pad(h, k, s); wo, pl = same_pad(w, k, s)
    d = ConvDesc(n, h, w, cin, ho, wo, cout, k, k, s, pt, pl, cin, cout, _C.BF16)
    dy = torch.randn(n, ho, wo, cout, device='cuda').bfloat16()
    wt = torch.randn(k, k, cin, cout, device='cuda') * 0.05
    bias = torch.randn(cin, device='cuda')
    pack = torch.zeros(lib.rcgan_conv_wpack_bytes(d), dtype=torch.uint8, device='cuda')
    call('rcgan_conv_wpack', d, wt.data_ptr(), None, pack.data_ptr(), st)
    out = torch.zeros(n, h, w, cin, device='cuda', dtype=torch.bfloat16)
    parts = torch.zeros(lib.rcgan_colstats_floats(cin), device='cuda')
    flops = 2.0 * n * ho * wo * cout * k * k * cin
    run('deconv n%d 16x16x256 -> 32x32x256 bare' % n,
        lambda: call('rcgan_conv2d_dgrad', d, dy.data_ptr(), wt.data_ptr(), pack.data_ptr(), bias.data_ptr(), out.data_ptr(), _C.BF16, 0, 0.0, 0, st),
        flops, MODES)
    ep = _C.ConvEpilogue(colstats=parts.data_ptr())
    run('deconv n%d 16x16x256 -> 32x32x256 colstats' % n,
        lambda: call('rcgan_conv2d_dgrad_ex', d, dy.data_ptr(), pack.data_ptr(), bias.data_ptr(), out.data_ptr(), _C.BF16, 0, 0.0, 0,
                     ctypes.byref(ep), st), flops, [(0, 0), (2, 0)])

# G's last 3x3 conv with its fused epilogue (upsampled shortcut + second relu output + column statistics)
for n in (512, 256):
    h = w = 32; cin = cout = 256; k = 3; s = 1
    ho, pt = same_pad(h, k, s); wo, pl = same_pad(w, k, s)
    d = ConvDesc(n, h, w, cin, ho, wo, cout, k, k, s, pt, pl, cin, cout, _C.BF16)
    x = torch.randn(n, h, w, cin, device='cuda').bfloat16()
    wt = torch.randn(k, k, cin, cout, device='cuda') * 0.05
    b = torch.randn(cout, device='cuda')
    pack = torch.zeros(lib.rcgan_conv_wpack_bytes(d), dtype=torch.uint8, device='cuda')
    call('rcgan_conv_wpack', d, wt.data_ptr(), None, pack.data_ptr(), st)
    y = torch.zeros(n, ho, wo, cout, device='cuda', dtype=torch.bfloat16)
    y2 = torch.zeros_like(y)
    res_s = torch.randn(n, ho // 2, wo // 2, cout, device='cuda').bfloat16()
    parts = torch.zeros(lib.rcgan_colstats_floats(cout), device='cuda')
    flops = 2.0 * n * ho * wo * cout * k * k * cin
    run('conv3x3 n%d 32x32x256 bare' % n,
        lambda: call('rcgan_conv2d_fprop', d, x.data_ptr(), wt.data_ptr(), pack.data_ptr(), b.data_ptr(), y.data_ptr(), _C.BF16, 0, 0.0, st),
        flops, [(0, 0), (0, 16), (0, 1), (0, 2), (2, 0)])
    for name, ep in (('res_up', _C.ConvEpilogue(res=res_s.data_ptr(), res_up=1, ld_res=cout)),
                     ('res_up+out2', _C.ConvEpilogue(res=res_s.data_ptr(), res_up=1, ld_res=cout, out2=y2.data_ptr(), out2_act=_C.ACT_RELU)),
                     ('res_up+colstats', _C.ConvEpilogue(res=res_s.data_ptr(), res_up=1, ld_res=cout, colstats=parts.data_ptr()))):
        run('conv3x3 n%d 32x32x256 %s' % (n, name),
            lambda: call('rcgan_conv2d_fprop_ex', d, x.data_ptr(), pack.data_ptr(), b.data_ptr(), y.data_ptr(), _C.BF16, 0, 0.0,
                         ctypes.byref(ep), st), flops, [(0, 0), (2, 0)])
