"""Cost of the conv epilogue fusions: the same conv launched bare and with each rcgan_conv_epilogue option (CUDA events,
20 launches each, L2 flushed in between).  usage: python tools/epi_bench.py"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
from robust_conditional_gan_b200 import _C
from robust_conditional_gan_b200._C import ConvDesc, call
from robust_conditional_gan_b200.graph import same_pad

lib = _C.load()
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')


def timeit(fn, reps=20):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2] * 1e3


for (n, h, w, cin, cout, k, s) in [(256, 32, 32, 256, 256, 3, 1), (512, 16, 16, 128, 128, 3, 1), (512, 8, 8, 128, 128, 3, 1)]:
    ho, pt = same_pad(h, k, s); wo, pl = same_pad(w, k, s)
    d = ConvDesc(n, h, w, cin, ho, wo, cout, k, k, s, pt, pl, cin, cout, _C.BF16)
    x = torch.randn(n, h, w, cin, device='cuda').bfloat16()
    wt = torch.randn(k, k, cin, cout, device='cuda') * 0.05
    b = torch.randn(cout, device='cuda')
    pack = torch.zeros(lib.rcgan_conv_wpack_bytes(d), dtype=torch.uint8, device='cuda')
    call('rcgan_conv_wpack', d, wt.data_ptr(), None, pack.data_ptr(), st)
    y = torch.zeros(n, ho, wo, cout, device='cuda', dtype=torch.bfloat16)
    y2 = torch.zeros_like(y)
    res = torch.randn(n, ho, wo, cout, device='cuda').bfloat16()
    res_s = torch.randn(n, ho // 2, wo // 2, cout, device='cuda').bfloat16()
    parts = torch.zeros(lib.rcgan_colstats_floats(cout), device='cuda')
    flops = 2.0 * n * ho * wo * cout * k * k * cin
    base = timeit(lambda: call('rcgan_conv2d_fprop', d, x.data_ptr(), wt.data_ptr(), pack.data_ptr(), b.data_ptr(), y.data_ptr(), _C.BF16, 0, 0.0, st))
    var = _C.last_conv_variant()
    rows = [('bare', base)]
    for name, ep in (('res (fprop_res path)', None),
                     ('ex: nothing but out2', _C.ConvEpilogue(out2=y2.data_ptr(), out2_act=_C.ACT_RELU)),
                     ('ex: res', _C.ConvEpilogue(res=res.data_ptr(), ld_res=cout)),
                     ('ex: res_up', _C.ConvEpilogue(res=res_s.data_ptr(), res_up=1, ld_res=cout)),
                     ('ex: colstats', _C.ConvEpilogue(colstats=parts.data_ptr())),
                     ('ex: res_up + colstats', _C.ConvEpilogue(res=res_s.data_ptr(), res_up=1, ld_res=cout, colstats=parts.data_ptr())),
                     ('ex: res + out2', _C.ConvEpilogue(res=res.data_ptr(), ld_res=cout, out2=y2.data_ptr(), out2_act=_C.ACT_RELU))):
        if ep is None:
            t = timeit(lambda: call('rcgan_conv2d_fprop_res', d, x.data_ptr(), wt.data_ptr(), pack.data_ptr(), b.data_ptr(), res.data_ptr(),
                                    y.data_ptr(), _C.BF16, 0, 0.0, st))
        else:
            try:
                t = timeit(lambda: call('rcgan_conv2d_fprop_ex', d, x.data_ptr(), pack.data_ptr(), b.data_ptr(), y.data_ptr(), _C.BF16, 0, 0.0,
                                        ctypes.byref(ep), st))
            except _C.RcganError as e:
                print('   %-26s unsupported (%s)' % (name, e))
                continue
        rows.append((name + ' [' + _C.last_conv_variant() + ']', t))
    print('conv n%d %dx%dx%d -> %d k%d  %s  %.1f GFLOP' % (n, h, w, cin, cout, k, var, flops / 1e9))
    for name, t in rows:
        print('   %-70s %8.1f us  %7.1f TFLOP/s' % (name, t, flops / t / 1e6))
