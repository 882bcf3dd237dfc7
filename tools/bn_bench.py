"""Micro-benchmark of the batch-norm entry points (debug helper, not a test)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robust_conditional_gan_b200 import _C

SHAPES = [  # samples, hw, c, n_labels (0 = unconditional)
    (1024, 196, 128, 0),
    (1024, 49, 128, 0),
    (1024, 1, 1024, 0),
    (256, 1024, 256, 10),
    (512, 1024, 256, 10),
    (512, 256, 256, 10),
    (512, 64, 256, 10),
]


def main():
    dev = torch.device('cuda:0')
    st = torch.cuda.current_stream().cuda_stream
    lib = _C.load()
    only = os.environ.get('BN_BENCH_ONLY')
    for si, (n, hw, c, nl) in enumerate(SHAPES):
        if only is not None and int(only) != si:
            continue
        x = torch.randn(n * hw, c, device=dev).bfloat16()
        y = torch.empty_like(x)
        dy = torch.randn(n * hw, c, device=dev).bfloat16()
        dx = torch.empty_like(x)
        L = max(nl, 1)
        scale = torch.rand(L, c, device=dev) + 0.5
        offset = torch.randn(L, c, device=dev)
        dscale, doffset = torch.zeros_like(scale), torch.zeros_like(offset)
        labels = torch.randint(0, L, (n,), device=dev, dtype=torch.int32) if nl else None
        save = torch.zeros(2 * c, device=dev)
        mm, mv = torch.zeros(c, device=dev), torch.ones(c, device=dev)
        wsb = lib.rcgan_bn_workspace(n, hw, c)
        ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)
        lp = labels.data_ptr() if labels is not None else None
        for v in ('4', '8'):
            os.environ['RCGAN_BN_STATS_V'] = v

            def fwd():
                _C.call('rcgan_bn_fwd', x.data_ptr(), y.data_ptr(), n, hw, c, _C.BF16, _C.BF16, scale.data_ptr(), offset.data_ptr(), lp,
                        1e-5, _C.ACT_RELU, 0.0, 1, 0.9, mm.data_ptr(), mv.data_ptr(), save.data_ptr(), ws.data_ptr(), wsb, st)

            def bwd():          # relu mask re-derived from x (5 tensor passes)
                _C.call('rcgan_bn_bwd', dy.data_ptr(), x.data_ptr(), None, dx.data_ptr(), n, hw, c, _C.BF16, _C.BF16,
                        scale.data_ptr(), lp, L, save.data_ptr(), _C.ACT_RELU, 0.0, dscale.data_ptr(), doffset.data_ptr(), 0, 0,
                        ws.data_ptr(), wsb, offset.data_ptr(), st)

            def bwd_y():        # relu mask read back from y (7 tensor passes)
                _C.call('rcgan_bn_bwd', dy.data_ptr(), x.data_ptr(), y.data_ptr(), dx.data_ptr(), n, hw, c, _C.BF16, _C.BF16,
                        scale.data_ptr(), lp, L, save.data_ptr(), _C.ACT_RELU, 0.0, dscale.data_ptr(), doffset.data_ptr(), 0, 0,
                        ws.data_ptr(), wsb, None, st)
            res = []
            for fn in (fwd, bwd, bwd_y):
                fn()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(20):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                res.append(e0.elapsed_time(e1) / 20 * 1e3)
            nbytes = n * hw * c * 2
            print(f'{(n, hw, c, nl)} statsV={v} fwd {res[0]:7.1f} us ({3 * nbytes / res[0] / 1e6:5.2f} TB/s of 3 passes)  '
                  f'bwd(mask from x) {res[1]:7.1f} us ({5 * nbytes / res[1] / 1e6:5.2f} TB/s of 5 passes)  '
                  f'bwd(mask from y) {res[2]:7.1f} us ({7 * nbytes / res[2] / 1e6:5.2f} TB/s of 7 passes)', flush=True)


if __name__ == '__main__':
    main()
