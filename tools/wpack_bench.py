"""Weight-pack kernel timings on the layers that matter (run from the repo root on a B200; RCGAN_WPACK_FLAT=1 = old kernel)."""
import sys, ctypes, torch
sys.path.insert(0, '.')
from robust_conditional_gan_b200 import _C
st = torch.cuda.current_stream().cuda_stream
def desc(n, h, w, cin, cout, k):
    d = _C.ConvDesc(); d.n, d.h, d.w, d.cin, d.ho, d.wo, d.cout = n, h, w, cin, h, w, cout
    d.kh = d.kw = k; d.stride = 1; d.pad_t = d.pad_l = (k - 1) // 2; d.ldx = (cin + 7) // 8 * 8; d.ldy = cout; d.dtype = _C.BF16
    return d
for shp in ((1024, 1, 1, 1034, 6272, 1), (256, 32, 32, 256, 256, 3), (256, 8, 8, 128, 128, 3), (256, 4, 4, 1024, 256, 3)):
    d = desc(*shp)
    n, h, w, cin, cout, k = shp
    wt = torch.randn(k, k, cin, cout, device='cuda')
    nb = _C.load().rcgan_conv_wpack_bytes(ctypes.byref(d))
    pack = torch.empty(nb, dtype=torch.uint8, device='cuda')
    f = lambda: _C.call('rcgan_conv_wpack', ctypes.byref(d), wt.data_ptr(), None, pack.data_ptr(), st)
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    print(shp, 'wpack us', e0.elapsed_time(e1) / 20 * 1e3)
