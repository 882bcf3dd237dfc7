"""Bandwidth-bound kernels (BN / CBN, spectral norm, losses, Adam, elementwise) vs the oracle."""
import ctypes
import math

import numpy as np
import pytest
import torch

from oracle import nn as O
from robust_conditional_gan_b200 import _C
from robust_conditional_gan_b200._C import call
from util import keep, TD, TOL, dev, maxabs, relerr, st

pytestmark = pytest.mark.gpu


def ws_buf(nbytes):
    return torch.zeros(max(int(nbytes), 256), dtype=torch.uint8, device='cuda')


# ----------------------------------------------------------------------------- batch norm
PAIRS = [(_C.F32, _C.F32), (_C.BF16, _C.BF16), (_C.F32, _C.BF16)]


@pytest.mark.parametrize('xdt,dtype', PAIRS)
@pytest.mark.parametrize('samples,hw,c,act', [(16, 1, 1024, 'relu'), (8, 196, 128, 'relu'), (16, 49, 64, 'lrelu'),
                                              (16, 4, 64, 'lrelu'), (5, 1, 74, 'lrelu'), (3, 1, 6272, 'relu'),
                                              (300, 16, 64, 'lrelu')])
def test_bn_fwd_bwd(lib, samples, hw, c, act, xdt, dtype):
    g = torch.Generator().manual_seed(1)
    td = TD[dtype]
    x = (torch.randn(samples, hw, c, generator=g) * 2 + 3).to(TD[xdt]).float()
    dy = torch.randn(samples, hw, c, generator=g).to(td).float()
    gamma = torch.rand(c, generator=g) + 0.5
    beta = torch.randn(c, generator=g)
    mm, mv = torch.randn(c, generator=g), torch.rand(c, generator=g) + 0.5
    actf = {'relu': torch.relu, 'lrelu': O.lrelu}[act]
    xr = x.double().requires_grad_(True); gr = gamma.double().requires_grad_(True); br = beta.double().requires_grad_(True)
    y0, mean, var = O.batch_norm_train(xr, gr, br)
    yr = actf(y0)
    yr.backward(dy.double())
    mm_ref, mv_ref = O.batch_norm_moving_update(mm.double(), mv.double(), mean.detach(), var.detach(), samples * hw)

    xd, dyd = dev(x, TD[xdt]), dev(dy, td)
    y = torch.zeros_like(dyd); dx = torch.zeros_like(dyd)
    save = torch.zeros(2 * c, device='cuda')
    mmd, mvd = dev(mm), dev(mv)
    nb = lib.rcgan_bn_workspace(samples, hw, c)
    ws = ws_buf(nb)
    A = {'relu': _C.ACT_RELU, 'lrelu': _C.ACT_LRELU}[act]
    call('rcgan_bn_fwd', xd.data_ptr(), y.data_ptr(), samples, hw, c, xdt, dtype, keep(dev(gamma)), keep(dev(beta)), None,
         1e-5, A, 0.2, 1, 0.9, mmd.data_ptr(), mvd.data_ptr(), save.data_ptr(), ws.data_ptr(), nb, st())
    assert relerr(y.float(), yr) < TOL[dtype]
    assert relerr(save[:c], mean) < 1e-5 and relerr(save[c:], torch.rsqrt(var + 1e-5)) < 1e-5
    assert relerr(mmd, mm_ref) < 1e-5 and relerr(mvd, mv_ref) < 1e-5
    dg, db = torch.zeros(c, device='cuda'), torch.zeros(c, device='cuda')
    # backward uses the oracle's y to decide the activation mask identically on both sides
    call('rcgan_bn_bwd', dyd.data_ptr(), xd.data_ptr(), y.data_ptr(), dx.data_ptr(), samples, hw, c, xdt, dtype,
         keep(dev(gamma)), None, 1, save.data_ptr(), A, 0.2, dg.data_ptr(), db.data_ptr(), 0, 0, ws.data_ptr(), nb, None, st())
    # the product's form: the relu / lrelu mask re-derived from x (sign of the forward pre-activation) instead of reading y back
    dx2, dg2, db2 = torch.zeros_like(dx), torch.zeros(c, device='cuda'), torch.zeros(c, device='cuda')
    call('rcgan_bn_bwd', dyd.data_ptr(), xd.data_ptr(), None, dx2.data_ptr(), samples, hw, c, xdt, dtype,
         keep(dev(gamma)), None, 1, save.data_ptr(), A, 0.2, dg2.data_ptr(), db2.data_ptr(), 0, 0, ws.data_ptr(), nb, keep(dev(beta)), st())
    assert torch.equal(dx2, dx) and torch.equal(dg2, dg) and torch.equal(db2, db)
    tol = TOL[dtype] * (3 if dtype == _C.BF16 else 1)
    assert relerr(dx.float(), xr.grad) < tol
    assert relerr(dg, gr.grad) < tol and relerr(db, br.grad) < tol
    # inference mode
    call('rcgan_bn_fwd', xd.data_ptr(), y.data_ptr(), samples, hw, c, xdt, dtype, keep(dev(gamma)), keep(dev(beta)), None,
         1e-5, A, 0.2, 0, 0.9, mmd.data_ptr(), mvd.data_ptr(), save.data_ptr(), ws.data_ptr(), nb, st())
    ref = actf(O.batch_norm_infer(x.double(), gamma.double(), beta.double(), mmd.double().cpu(), mvd.double().cpu()))
    assert relerr(y.float(), ref) < TOL[dtype]


@pytest.mark.parametrize('xdt,dtype', PAIRS)
@pytest.mark.parametrize('samples,hw,c,c2', [(8, 196, 128, 10), (16, 1, 1024, 10), (5, 49, 64, 3)])
def test_bn_with_fused_label_concat(lib, samples, hw, c, c2, xdt, dtype):
    """rcgan_bn_fwd_cat / rcgan_bn_bwd_cat: concat([relu(BN(x)), y broadcast]) written by the norm itself (generator,
    mnist/model.py:714-728), padding untouched, backward ignoring the label channels' gradient."""
    g = torch.Generator().manual_seed(2)
    td = TD[dtype]
    ld = (c + c2 + 7) // 8 * 8
    x = (torch.randn(samples, hw, c, generator=g) + 1).to(TD[xdt]).float()
    yb = torch.randn(samples, c2, generator=g)
    dyfull = torch.randn(samples, hw, ld, generator=g).to(td).float()
    gamma = torch.rand(c, generator=g) + 0.5
    beta = torch.randn(c, generator=g)
    xr = x.double().requires_grad_(True); gr = gamma.double().requires_grad_(True); br = beta.double().requires_grad_(True)
    y0, mean, var = O.batch_norm_train(xr, gr, br)
    yr = torch.relu(y0)
    yr.backward(dyfull[..., :c].double())
    xd, dyd = dev(x, TD[xdt]), dev(dyfull, td)
    y = torch.full((samples, hw, ld), 7.0, device='cuda', dtype=td)
    save = torch.zeros(2 * c, device='cuda')
    nb = lib.rcgan_bn_workspace(samples, hw, c)
    ws = ws_buf(nb)
    call('rcgan_bn_fwd_cat', xd.data_ptr(), y.data_ptr(), ld, keep(dev(yb)), c2, samples, hw, c, xdt, dtype, keep(dev(gamma)),
         keep(dev(beta)), None, 1e-5, _C.ACT_RELU, 0.0, 1, 0.9, None, None, save.data_ptr(), ws.data_ptr(), nb, st())
    assert relerr(y[..., :c].float(), yr) < TOL[dtype]
    lab = yb.to(td).float()[:, None, :].expand(samples, hw, c2)
    assert float((y[..., c:c + c2].float().cpu() - lab).abs().max()) == 0.0
    if ld > c + c2:
        assert float((y[..., c + c2:].float() - 7.0).abs().max()) == 0.0        # padding untouched
    dx = torch.zeros(samples, hw, c, device='cuda', dtype=td)
    dg, db = torch.zeros(c, device='cuda'), torch.zeros(c, device='cuda')
    call('rcgan_bn_bwd_cat', dyd.data_ptr(), xd.data_ptr(), y.data_ptr(), ld, dx.data_ptr(), samples, hw, c, xdt, dtype,
         keep(dev(gamma)), None, 1, save.data_ptr(), _C.ACT_RELU, 0.0, dg.data_ptr(), db.data_ptr(), 0, 0, ws.data_ptr(), nb, None, st())
    dx2, dg2, db2 = torch.zeros_like(dx), torch.zeros(c, device='cuda'), torch.zeros(c, device='cuda')
    call('rcgan_bn_bwd_cat', dyd.data_ptr(), xd.data_ptr(), None, ld, dx2.data_ptr(), samples, hw, c, xdt, dtype, keep(dev(gamma)), None, 1,
         save.data_ptr(), _C.ACT_RELU, 0.0, dg2.data_ptr(), db2.data_ptr(), 0, 0, ws.data_ptr(), nb, keep(dev(beta)), st())
    assert torch.equal(dx2, dx) and torch.equal(dg2, dg) and torch.equal(db2, db)
    tol = TOL[dtype] * (3 if dtype == _C.BF16 else 1)
    assert relerr(dx.float(), xr.grad) < tol
    assert relerr(dg, gr.grad) < tol and relerr(db, br.grad) < tol


@pytest.mark.parametrize('xdt,dtype', PAIRS)
@pytest.mark.parametrize('n,hh,c', [(6, 4, 1024), (6, 8, 256), (4, 32, 256), (64, 4, 64)])
def test_cond_batchnorm(lib, n, hh, c, xdt, dtype):
    g = torch.Generator().manual_seed(2)
    td = TD[dtype]
    hw = hh * hh
    x = (torch.randn(n, hh, hh, c, generator=g) + 0.5).to(TD[xdt]).float()
    dy = torch.randn(n, hh, hh, c, generator=g).to(td).float()
    labels = torch.randint(0, 10, (n,), generator=g)
    scale = torch.rand(10, c, generator=g) + 0.5
    offset = torch.randn(10, c, generator=g)
    xr = x.double().requires_grad_(True); sr = scale.double().requires_grad_(True); orf = offset.double().requires_grad_(True)
    yr = torch.relu(O.cond_batchnorm(xr, labels, orf, sr))
    yr.backward(dy.double())
    xd, dyd = dev(x, TD[xdt]), dev(dy, td)
    y = torch.zeros_like(dyd); dx = torch.zeros_like(dyd)
    save = torch.zeros(2 * c, device='cuda')
    lab = labels.to('cuda', torch.int32)
    nb = lib.rcgan_bn_workspace(n, hw, c)
    ws = ws_buf(nb)
    call('rcgan_bn_fwd', xd.data_ptr(), y.data_ptr(), n, hw, c, xdt, dtype, keep(dev(scale)), keep(dev(offset)),
         lab.data_ptr(), 1e-5, _C.ACT_RELU, 0.0, 1, 0.9, None, None, save.data_ptr(), ws.data_ptr(), nb, st())
    assert relerr(y.float(), yr) < TOL[dtype]
    ds, do = torch.ones(10, c, device='cuda'), torch.ones(10, c, device='cuda')
    call('rcgan_bn_bwd', dyd.data_ptr(), xd.data_ptr(), y.data_ptr(), dx.data_ptr(), n, hw, c, xdt, dtype, keep(dev(scale)),
         lab.data_ptr(), 10, save.data_ptr(), _C.ACT_RELU, 0.0, ds.data_ptr(), do.data_ptr(), 0, 0, ws.data_ptr(), nb, None, st())
    dx2, ds2, do2 = torch.zeros_like(dx), torch.ones(10, c, device='cuda'), torch.ones(10, c, device='cuda')
    call('rcgan_bn_bwd', dyd.data_ptr(), xd.data_ptr(), None, dx2.data_ptr(), n, hw, c, xdt, dtype, keep(dev(scale)), lab.data_ptr(), 10,
         save.data_ptr(), _C.ACT_RELU, 0.0, ds2.data_ptr(), do2.data_ptr(), 0, 0, ws.data_ptr(), nb, keep(dev(offset)), st())
    assert torch.equal(dx2, dx) and torch.equal(ds2, ds) and torch.equal(do2, do)
    tol = TOL[dtype] * (3 if dtype == _C.BF16 else 1)
    assert relerr(dx.float(), xr.grad) < tol
    assert relerr(ds, sr.grad) < tol and relerr(do, orf.grad) < tol


# ----------------------------------------------------------------------------- spectral norm
@pytest.mark.parametrize('m,c', [(25, 64), (275, 64), (1600, 64), (1152, 128), (27, 128), (3, 128), (128, 1), (300, 128),
                                 (3072, 10), (9216, 256)])
def test_spectral_norm(lib, m, c):
    g = torch.Generator().manual_seed(3)
    W = torch.randn(m, c, generator=g) * 0.05
    u = torch.randn(1, c, generator=g)
    G = torch.randn(m, c, generator=g)
    Wr = W.double().requires_grad_(True)
    Wb, u_new, sigma = O.spectral_normed_weight(Wr, u.double())
    Wb.backward(G.double())
    Wd, ud, Gd = dev(W), dev(u), dev(G)
    wbar = torch.zeros(m, c, device='cuda'); un = torch.zeros(c, device='cuda')
    save = torch.zeros(lib.rcgan_sn_save_floats(m, c), device='cuda')
    nb = lib.rcgan_sn_workspace(m, c); ws = ws_buf(nb)
    call('rcgan_sn_fwd', Wd.data_ptr(), ud.data_ptr(), m, c, wbar.data_ptr(), un.data_ptr(), save.data_ptr(), ws.data_ptr(),
         nb, st())
    assert abs(float(save[0]) - float(sigma)) / float(sigma) < 1e-5
    assert relerr(wbar, Wb) < 1e-5 and relerr(un, u_new) < 1e-5
    dW = torch.full((m, c), 1.0, device='cuda')
    call('rcgan_sn_bwd', Wd.data_ptr(), ud.data_ptr(), Gd.data_ptr(), m, c, save.data_ptr(), dW.data_ptr(), 0, ws.data_ptr(),
         nb, st())
    assert relerr(dW, Wr.grad) < 2e-5          # full gradient through the power iteration
    const_uv = G.double() / sigma.detach()      # the "u, v constant" shortcut is measurably different
    assert relerr(dW, const_uv) > 1e-4 or m * c < 200


def test_spectral_norm_batched_matches_oracle(lib):
    """rcgan_sn_fwd_batched / rcgan_sn_bwd_batched over weights of different shapes (more than one 32-item launch group)
    against the oracle, item by item; one item accumulates into dW."""
    import ctypes
    shapes = [(1152, 128), (27, 128), (128, 1), (10, 128), (2304, 256), (25, 64)] * 6        # 36 items
    g = torch.Generator().manual_seed(11)
    n = len(shapes)
    Ws = [torch.randn(m, c, generator=g) * 0.05 for m, c in shapes]
    us = [torch.randn(1, c, generator=g) for m, c in shapes]
    Gs = [torch.randn(m, c, generator=g) for m, c in shapes]
    Wd, ud, Gd = [dev(t) for t in Ws], [dev(t) for t in us], [dev(t) for t in Gs]
    wbar = [torch.zeros(m, c, device='cuda') for m, c in shapes]
    un = [torch.zeros(c, device='cuda') for m, c in shapes]
    save = [torch.zeros(lib.rcgan_sn_save_floats(m, c), device='cuda') for m, c in shapes]
    dW = [torch.full((m, c), 1.0, device='cuda') for m, c in shapes]
    PA, IA = ctypes.c_void_p * n, ctypes.c_int * n
    ptrs = lambda ts: PA(*[t.data_ptr() for t in ts])
    ms, cs = IA(*[m for m, c in shapes]), IA(*[c for m, c in shapes])
    acc = IA(*[1 if i == 2 else 0 for i in range(n)])
    nb = lib.rcgan_sn_workspace_batched(n, ms, cs)
    ws = ws_buf(nb)
    call('rcgan_sn_fwd_batched', n, ptrs(Wd), ptrs(ud), ms, cs, ptrs(wbar), ptrs(un), ptrs(save), ws.data_ptr(), nb, st())
    call('rcgan_sn_bwd_batched', n, ptrs(Wd), ptrs(ud), ptrs(Gd), ms, cs, ptrs(save), ptrs(dW), acc, ws.data_ptr(), nb, st())
    for i, (m, c) in enumerate(shapes):
        Wr = Ws[i].double().requires_grad_(True)
        Wb, u_new, sigma = O.spectral_normed_weight(Wr, us[i].double())
        Wb.backward(Gs[i].double())
        assert relerr(wbar[i], Wb) < 1e-5 and relerr(un[i], u_new) < 1e-5, i
        ref = Wr.grad + (1.0 if i == 2 else 0.0)
        assert relerr(dW[i], ref) < 2e-5, i


# ----------------------------------------------------------------------------- losses
def phi_ref(mode, l):
    if mode == _C.HINGE_D_REAL: return torch.relu(1 - l)
    if mode == _C.HINGE_D_FAKE: return torch.relu(1 + l)
    if mode == _C.HINGE_G: return -l
    if mode == _C.CE_D_FAKE: return O.sigmoid_ce(l, torch.zeros_like(l))
    return O.sigmoid_ce(l, torch.ones_like(l))


@pytest.mark.parametrize('mode', [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize('B,d,wkind', [(64, 64, 'onehot'), (1000, 128, 'soft'), (37, 64, 'cinv')])
def test_channel_loss(lib, mode, B, d, wkind):
    g = torch.Generator().manual_seed(4)
    k = 10
    h = torch.randn(B, d, generator=g); psi = torch.randn(B, generator=g); V = torch.randn(k, d, generator=g) * 0.3
    lab = torch.randint(0, k, (B,), generator=g)
    if wkind == 'onehot':
        w = torch.eye(k)[lab]
    elif wkind == 'soft':
        w = torch.softmax(torch.randn(k, k, generator=g), -1)[lab]
    else:
        w = torch.linalg.inv(torch.eye(k) * 0.5 + 0.05)[lab].float()
    hr, pr, Vr, wr = [t.double().requires_grad_(True) for t in (h, psi, V, w)]
    logits = pr[:, None] + hr @ Vr.t()
    L = (phi_ref(mode, logits) * wr).sum(1).mean()
    L.backward()
    loss = torch.zeros(1, device='cuda'); lg = torch.zeros(B, k, device='cuda')
    dh = torch.zeros(B, d, device='cuda'); dpsi = torch.zeros(B, device='cuda')
    dV = torch.zeros(k, d, device='cuda'); dw = torch.zeros(B, k, device='cuda')
    call('rcgan_channel_loss', keep(dev(h)), keep(dev(psi)), keep(dev(V)), keep(dev(w)), B, d, k, _C.F32,
         mode, 1.0 / B, loss.data_ptr(), lg.data_ptr(), dh.data_ptr(), 0, dpsi.data_ptr(), dV.data_ptr(), dw.data_ptr(), st())
    assert abs(float(loss) - float(L)) < 1e-5 * max(1.0, abs(float(L)))
    assert relerr(lg, logits) < 1e-5
    for got, ref in ((dh, hr.grad), (dpsi, pr.grad), (dV, Vr.grad), (dw, wr.grad)):
        assert relerr(got, ref) < 2e-5
    if wkind == 'onehot':   # bit-identical to gathering V[label] (the reference's known-C path)
        single = (psi.cuda() + (h.cuda() * V.cuda()[lab.cuda()]).sum(1))
        assert maxabs(lg[torch.arange(B), lab.cuda()], single) < 1e-5


def test_sigmoid_ce_softmax_gather_logit(lib):
    g = torch.Generator().manual_seed(5)
    B, k = 77, 10
    l = torch.randn(B, k, generator=g) * 3; t = torch.eye(k)[torch.randint(0, k, (B,), generator=g)]
    lr = l.double().requires_grad_(True)
    L = O.sigmoid_ce(lr, t.double()).mean(); L.backward()
    loss = torch.zeros(1, device='cuda'); dl = torch.zeros(B, k, device='cuda')
    call('rcgan_sigmoid_ce', keep(dev(l)), keep(dev(t)), B * k, 1.0 / (B * k), loss.data_ptr(), dl.data_ptr(), st())
    assert abs(float(loss) - float(L)) < 1e-5 and relerr(dl, lr.grad) < 1e-5
    # softmax rows fwd/bwd
    Lg = torch.randn(k, k, generator=g); dC = torch.randn(k, k, generator=g)
    Lr = Lg.double().requires_grad_(True); C = torch.softmax(Lr, -1); C.backward(dC.double())
    Cd = torch.zeros(k, k, device='cuda'); dL = torch.zeros(k, k, device='cuda')
    call('rcgan_softmax_rows_fwd', keep(dev(Lg)), Cd.data_ptr(), k, k, st())
    call('rcgan_softmax_rows_bwd', Cd.data_ptr(), keep(dev(dC)), dL.data_ptr(), k, k, 0, st())
    assert relerr(Cd, C) < 1e-6 and relerr(dL, Lr.grad) < 1e-5
    # gather rows
    y = torch.randint(0, k, (B,), generator=g).to(torch.int32)
    w = torch.zeros(B, k, device='cuda'); dwg = torch.randn(B, k, generator=g)
    call('rcgan_gather_rows_fwd', Cd.data_ptr(), keep(y.cuda()), w.data_ptr(), B, k, st())
    assert maxabs(w, Cd[y.long().cuda()]) == 0.0
    dCd = torch.zeros(k, k, device='cuda')
    call('rcgan_gather_rows_bwd', keep(dev(dwg)), keep(y.cuda()), dCd.data_ptr(), B, k, k, 0, st())
    ref = torch.zeros(k, k, dtype=torch.float64).index_add_(0, y.long(), dwg.double())
    assert relerr(dCd, ref) < 1e-5
    # scalar logit loss
    for mode in range(6):
        x = torch.randn(B, generator=g) * 2; xr = x.double().requires_grad_(True)
        Lm = phi_ref(mode, xr).mean(); Lm.backward()
        loss.zero_(); d1 = torch.zeros(B, device='cuda')
        call('rcgan_logit_loss', keep(dev(x)), B, mode, 1.0 / B, loss.data_ptr(), d1.data_ptr(), st())
        assert abs(float(loss) - float(Lm)) < 1e-5 and relerr(d1, xr.grad) < 1e-5


# ----------------------------------------------------------------------------- adam
def test_adam_tf(lib):
    g = torch.Generator().manual_seed(6)
    n = 100003
    p = torch.randn(n, generator=g); gr = torch.randn(n, generator=g) * 1e-3
    gr[:100] = 1e-9      # tiny gradients: TF's eps-outside form differs from torch.optim.Adam here
    opt = O.TFAdam(['p'], 2e-4, 0.5, clip=())
    P = {'p': p.double()}
    pd, gd = dev(p), dev(gr); m = torch.zeros(n, device='cuda'); v = torch.zeros(n, device='cuda')
    lr_dev = torch.zeros(1, device='cuda')
    lo = (ctypes.c_long * 8)(10, 0, 0, 0, 0, 0, 0, 0); hi = (ctypes.c_long * 8)(5000, 0, 0, 0, 0, 0, 0, 0)
    for t in range(1, 4):
        opt.step(P, {'p': gr.double() * 0.5})
        lr_t = 2e-4 * math.sqrt(1 - 0.999 ** t) / (1 - 0.5 ** t)
        lr_dev.fill_(lr_t)
        call('rcgan_adam_tf', pd.data_ptr(), gd.data_ptr(), m.data_ptr(), v.data_ptr(), n, 0.0, lr_dev.data_ptr(), 0.5, 0.999,
             1e-8, 0.5, lo, hi, 1, st())
    ref = P['p'].clone(); ref[10:5000] = ref[10:5000].clamp(-1, 1)
    # clip applied every step in the kernel vs once here: identical for |p|>1 elements that only move by ~lr
    assert maxabs(pd[5000:], P['p'][5000:]) < 1e-6
    assert float(pd[10:5000].abs().max()) <= 1.0


# ----------------------------------------------------------------------------- elementwise
@pytest.mark.parametrize('dtype', [_C.F32, _C.BF16])
def test_elementwise(lib, dtype):
    g = torch.Generator().manual_seed(7)
    td = TD[dtype]
    n, h, w, c = 3, 8, 8, 24
    x = torch.randn(n, h, w, c, generator=g).to(td).float()
    yb = torch.eye(10)[torch.randint(0, 10, (n,), generator=g)]
    xd = dev(x, td)
    # concat label + slice backward
    out = torch.zeros(n, h, w, 40, device='cuda', dtype=td)
    call('rcgan_concat_label_fwd', xd.data_ptr(), c, keep(dev(yb)), out.data_ptr(), 40, n * h * w, h * w, c, 10, dtype, st())
    ref = O.conv_cond_concat(x, yb)
    assert maxabs(out[..., :34].float(), ref) == 0.0 and float(out[..., 34:].float().abs().max()) == 0.0
    da = torch.zeros_like(xd)
    call('rcgan_slice_bwd', out.data_ptr(), 40, da.data_ptr(), c, n * h * w, c, dtype, 0, st())
    assert maxabs(da.float(), x) == 0.0
    # mean over hw (+relu) fwd/bwd
    for relu in (0, 1):
        xr = x.double().requires_grad_(True)
        m = (torch.relu(xr) if relu else xr).mean((1, 2)); dy = torch.randn(n, c, generator=g).to(td).float()
        m.backward(dy.double())
        y = torch.zeros(n, c, device='cuda', dtype=td); dx = torch.zeros_like(xd)
        call('rcgan_meanhw_fwd', xd.data_ptr(), y.data_ptr(), n, h * w, c, dtype, relu, st())
        call('rcgan_meanhw_bwd', keep(dev(dy, td)), xd.data_ptr(), dx.data_ptr(), n, h * w, c, dtype, relu, 0, st())
        assert relerr(y.float(), m) < TOL[dtype] and relerr(dx.float(), xr.grad) < TOL[dtype]
    # avgpool2 == add_n of strided slices / 4 ; upsample == concat x4 + depth_to_space
    xr = x.double().requires_grad_(True)
    p = (xr[:, ::2, ::2] + xr[:, 1::2, ::2] + xr[:, ::2, 1::2] + xr[:, 1::2, 1::2]) / 4.
    dy = torch.randn(n, h // 2, w // 2, c, generator=g).to(td).float(); p.backward(dy.double())
    y = torch.zeros(n, h // 2, w // 2, c, device='cuda', dtype=td); dx = torch.zeros_like(xd)
    call('rcgan_avgpool2_fwd', xd.data_ptr(), y.data_ptr(), n, h, w, c, dtype, st())
    call('rcgan_avgpool2_bwd', keep(dev(dy, td)), dx.data_ptr(), n, h, w, c, dtype, 0, st())
    assert relerr(y.float(), p) < TOL[dtype] and relerr(dx.float(), xr.grad) < TOL[dtype]
    xr = x.double().requires_grad_(True)
    cat = torch.cat([xr, xr, xr, xr], 3)     # depth_to_space(block 2), NHWC
    up = cat.reshape(n, h, w, 2, 2, c).permute(0, 1, 3, 2, 4, 5).reshape(n, 2 * h, 2 * w, c)
    dy = torch.randn(n, 2 * h, 2 * w, c, generator=g).to(td).float(); up.backward(dy.double())
    y = torch.zeros(n, 2 * h, 2 * w, c, device='cuda', dtype=td); dx = torch.zeros_like(xd)
    call('rcgan_upsample2_fwd', xd.data_ptr(), y.data_ptr(), n, h, w, c, dtype, st())
    call('rcgan_upsample2_bwd', keep(dev(dy, td)), dx.data_ptr(), n, h, w, c, dtype, 0, st())
    assert maxabs(y.float(), up) == 0.0 and relerr(dx.float(), xr.grad) < TOL[dtype]
    # activations
    for act, f in ((_C.ACT_LRELU, O.lrelu), (_C.ACT_SIGMOID, torch.sigmoid), (_C.ACT_TANH, torch.tanh), (_C.ACT_RELU, torch.relu)):
        xr = x.double().requires_grad_(True); yr = f(xr); dy = torch.randn(x.shape, generator=g).to(td).float()
        yr.backward(dy.double())
        y = torch.zeros_like(xd); dx = torch.zeros_like(xd)
        call('rcgan_bias_act_fwd', xd.data_ptr(), None, y.data_ptr(), n * h * w, c, c, c, dtype, act, 0.2, st())
        call('rcgan_act_bwd', keep(dev(dy, td)), y.data_ptr(), dx.data_ptr(), n * h * w, c, c, c, c, dtype, act, 0.2, 0, st())
        assert relerr(y.float(), yr) < TOL[dtype] and relerr(dx.float(), xr.grad) < 3 * TOL[dtype]


def test_preprocess_cifar(lib):
    g = torch.Generator().manual_seed(8)
    n = 5
    raw = torch.randint(0, 256, (n, 3072), generator=g).to(torch.int32)
    noise = torch.rand(n, 3072, generator=g) / 128
    ref = (2 * (raw.float() / 256. - .5) + noise).reshape(n, 3, 32, 32).permute(0, 2, 3, 1)
    out = torch.zeros(n, 32, 32, 3, device='cuda')
    call('rcgan_preprocess_cifar', keep(raw.cuda()), keep(dev(noise)), out.data_ptr(), n, _C.F32, st())
    assert maxabs(out, ref) < 1e-6


# ----------------------------------------------------------------------------- resampling folded into the filter
@pytest.mark.parametrize('mode', [0, 1])
@pytest.mark.parametrize('cin,cout', [(128, 128), (256, 256), (24, 40)])
def test_wfold4_and_adjoint(lib, mode, cin, cout):
    """rcgan_wfold4 == oracle fold4 (ConvMeanPool -> 4x4 s2 conv, UpsampleConv -> 4x4 s2 conv2d_transpose; the conv identities
    themselves are in tests/test_oracle_nn.py), rcgan_wfold4_bwd == its autograd adjoint, overwrite and accumulate"""
    g = torch.Generator().manual_seed(cin + mode)
    w = torch.randn(3, 3, cin, cout, generator=g)
    gw4 = torch.randn((4, 4, cin, cout) if mode == 0 else (4, 4, cout, cin), generator=g)
    wr = w.double().requires_grad_(True)
    ref = O.fold4(wr, mode)
    (ref * gw4.double()).sum().backward()
    wd, gd = dev(w), dev(gw4)
    w4 = torch.zeros(16 * cin * cout, device='cuda')
    call('rcgan_wfold4', wd.data_ptr(), w4.data_ptr(), cin, cout, mode, st())
    assert relerr(w4.reshape(ref.shape), ref) < 1e-6
    dw = torch.full((9 * cin * cout,), 2.0, device='cuda')
    call('rcgan_wfold4_bwd', gd.data_ptr(), dw.data_ptr(), cin, cout, mode, 0, st())
    assert relerr(dw.reshape(w.shape), wr.grad) < 1e-6
    call('rcgan_wfold4_bwd', gd.data_ptr(), dw.data_ptr(), cin, cout, mode, 1, st())
    assert relerr(dw.reshape(w.shape), 2 * wr.grad) < 1e-6


@pytest.mark.parametrize('dtype', [_C.F32, _C.BF16])
@pytest.mark.parametrize('rows,c,ld,act', [(4096, 128, 128, _C.ACT_RELU), (70000, 64, 64, _C.ACT_LRELU), (300, 24, 32, _C.ACT_TANH),
                                            (50, 10, 16, _C.ACT_SIGMOID)])
def test_act_bwd_colsum_equals_the_two_kernels(lib, rows, c, ld, act, dtype):
    """rcgan_act_bwd_colsum == rcgan_act_bwd (in place) followed by rcgan_colsum, bit for bit on dy, to fp32 summation order on db"""
    g = torch.Generator().manual_seed(rows)
    td = TD[dtype]
    dy = torch.randn(rows, ld, generator=g).to(td).cuda()
    y = torch.randn(rows, ld, generator=g).to(td).cuda()
    db0 = torch.randn(c, generator=g).cuda()
    for acc in (0, 1):
        a, b_ = dy.clone(), db0.clone()
        call('rcgan_act_bwd', a.data_ptr(), y.data_ptr(), a.data_ptr(), rows, c, ld, ld, ld, dtype, act, 0.2, 0, st())
        call('rcgan_colsum', a.data_ptr(), rows, c, ld, dtype, b_.data_ptr(), acc, st())
        a2, b2 = dy.clone(), db0.clone()
        call('rcgan_act_bwd_colsum', a2.data_ptr(), y.data_ptr(), rows, c, ld, ld, dtype, act, 0.2, b2.data_ptr(), acc, st())
        assert torch.equal(a2, a)
        assert relerr(b2, b_) < 1e-5
