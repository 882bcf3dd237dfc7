"""Micro-benchmark of single conv launches through the C ABI (debug helper, not a test).
usage: python tools/conv_bench.py  [variant ...]"""
import os, sys, ctypes
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robust_conditional_gan_b200 import _C

SHAPES = [  # n, h, w, cin, cout, k
    (256, 8, 8, 128, 128, 3),
    (256, 16, 16, 128, 128, 3),
    (256, 32, 32, 128, 128, 3),
    (512, 16, 16, 256, 256, 3),
    (512, 32, 32, 256, 256, 1),
    (256, 32, 32, 256, 256, 3),
    (512, 32, 32, 256, 256, 3),
    (1024, 14, 14, 128, 128, 5),
    (200704, 1, 1, 32, 138, 1, 6),
    (524288, 1, 1, 256, 27, 1, 5),
    (200704, 1, 1, 144, 25, 1, 7),
    (1024, 14, 14, 64, 138, 5, 6),
    (1024, 14, 14, 128, 138, 5, 6, 2),   # the conv g_h2 transposes (bench.py's MNIST roofline kernel is its dgrad)
    (512, 16, 16, 128, 128, 3),          # 13: CIFAR D.Block.2 at [real; fake]
    (512, 8, 8, 128, 128, 3),            # 14: CIFAR D.Block.3-6
    (512, 16, 16, 128, 128, 4, 0, 2),    # 15: the folded ConvMeanPool of D.Block.2 (4x4 stride 2)
]


def desc(n, h, w, cin, cout, k, pady=0, stride=1):
    from robust_conditional_gan_b200.graph import same_pad
    d = _C.ConvDesc()
    ho, pt = same_pad(h, k, stride)
    wo, pl = same_pad(w, k, stride)
    d.n, d.h, d.w, d.cin, d.ho, d.wo, d.cout = n, h, w, cin, ho, wo, cout
    d.kh = d.kw = k
    d.stride = stride
    d.pad_t, d.pad_l = pt, pl
    d.ldx, d.ldy, d.dtype = cin, cout + pady, _C.BF16
    return d


def main():
    variants = [v for v in sys.argv[1:]] or ['0', '1', '2']
    dev = torch.device('cuda:0')
    st = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    only = os.environ.get('CONV_BENCH_ONLY')
    for si, shp in enumerate(SHAPES):
        if only is not None and int(only) != si:
            continue
        n, h, w, cin, cout, k = shp[:6]
        pady = shp[6] if len(shp) > 6 else 0
        stride = shp[7] if len(shp) > 7 else 1
        ho, wo = -(-h // stride), -(-w // stride)
        d = desc(*shp)
        x = torch.randn(n, h, w, cin, device=dev).bfloat16()
        wt = torch.randn(k, k, cin, cout, device=dev) * 0.05
        y = torch.empty(n, ho, wo, cout + pady, device=dev, dtype=torch.bfloat16)
        nb = _C.load().rcgan_conv_wpack_bytes(ctypes.byref(d))
        pack = torch.empty(nb, dtype=torch.uint8, device=dev)
        _C.call('rcgan_conv_wpack', ctypes.byref(d), wt.data_ptr(), None, pack.data_ptr(), st)
        ref = None
        for var in variants:
            os.environ['RCGAN_TC_PERSIST'] = var.split(':')[0]
            os.environ['RCGAN_TC_DBG'] = var.split(':')[1] if ':' in var else '0'
            os.environ['RCGAN_TC_PAIR'] = var.split(':')[2] if var.count(':') >= 2 else '1'      # PERSIST:DBG:PAIR
            ts = []
            for it in range(6):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _C.call('rcgan_conv2d_fprop', ctypes.byref(d), x.data_ptr(), wt.data_ptr(), pack.data_ptr(), None, y.data_ptr(),
                        _C.BF16, 0, 0.0, st)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            # back-to-back (warm L2) timing
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for it in range(20):
                _C.call('rcgan_conv2d_fprop', ctypes.byref(d), x.data_ptr(), wt.data_ptr(), pack.data_ptr(), None, y.data_ptr(),
                        _C.BF16, 0, 0.0, st)
            e1.record()
            torch.cuda.synchronize()
            warm = e0.elapsed_time(e1) / 20
            if ref is None:
                ref = y.float().clone()
                err = 0.0
            else:
                err = float((y.float() - ref).abs().max())
            fl = 2.0 * n * ho * wo * cin * cout * k * k
            if os.environ.get('CONV_BENCH_BWD') and var == variants[0]:
                dy = torch.randn(n, ho, wo, cout + pady, device=dev).bfloat16()
                dx = torch.empty(n, h, w, cin, device=dev, dtype=torch.bfloat16)
                dw = torch.empty(k, k, cin, cout, device=dev)
                db = torch.empty(cout, device=dev)
                wsb = _C.load().rcgan_conv2d_wgrad_workspace(ctypes.byref(d))
                ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=dev)
                for nm, fn in (('dgrad', lambda: _C.call('rcgan_conv2d_dgrad', ctypes.byref(d), dy.data_ptr(), wt.data_ptr(), pack.data_ptr(),
                                                          None, dx.data_ptr(), _C.BF16, 0, 0.0, 0, st)),
                               ('wgrad', lambda: _C.call('rcgan_conv2d_wgrad', ctypes.byref(d), x.data_ptr(), dy.data_ptr(), dw.data_ptr(), 0,
                                                          ws.data_ptr(), wsb, st)),
                               ('colsum', lambda: _C.call('rcgan_colsum', dy.data_ptr(), n * ho * wo, cout, cout + pady, _C.BF16, db.data_ptr(), 0, st))):
                  for wv in os.environ.get('WG_WAVES', '2').split(','):
                    os.environ['RCGAN_WG_WAVES_X2'] = wv
                    if nm != 'wgrad' and wv != os.environ.get('WG_WAVES', '2').split(',')[0]:
                        continue
                    fn()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for it in range(20):
                        fn()
                    e1.record()
                    torch.cuda.synchronize()
                    tt = e0.elapsed_time(e1) / 20
                    print(f'{shp} {nm} wv={wv} warm {tt*1e3:8.1f} us  {fl/tt/1e9:7.1f} TF/s', flush=True)
                refw = torch.einsum('nhwc,nhwd->cd', x.float(), dy.float()) if k == 1 else None
                if refw is not None:
                    print('   wgrad relerr', float((dw[0, 0] - refw).abs().max() / refw.abs().max()))
            print(f'{shp} var={var} [{_C.last_conv_variant()}] cold {min(ts[1:])*1e3:8.1f} us  warm {warm*1e3:8.1f} us  {fl/warm/1e9:7.1f} TF/s  maxdiff_vs_var0 {err:.3g}', flush=True)


if __name__ == '__main__':
    main()
