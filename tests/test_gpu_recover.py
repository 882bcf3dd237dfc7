"""DCGAN.recover_labels (mnist/model.py:494-640): gradient descent on z_recover / y_logit_recover through
gen_sampler, CUDA path vs the oracle restatement (oracle/mnist.py recover_loss / recover_step)."""
import numpy as np
import pytest
import torch

from oracle import mnist as OM
from robust_conditional_gan_b200 import _C
from robust_conditional_gan_b200._C import call
from robust_conditional_gan_b200.model import DCGAN, default_flags
from util import dev, keep, relerr, st

pytestmark = pytest.mark.gpu


def build(R, precision, graph=True):
    kw = dict(algorithm='rcgan', disc_type='projection', estimate_confuse=True)
    flags = default_flags(batch_size=8, alpha=0.5, perm_regularizer=True, **kw)
    ocfg = OM.default_config(batch_size=8, alpha=0.5, perm_regularizer=True, **kw)
    model = DCGAN(batch_size=8, algorithm=flags.algorithm, estimate_confuse=True, perm_regularizer=True, alpha=0.5,
                  disc_type=flags.disc_type, config=flags, precision=precision, use_cuda_graph=graph)
    P = OM.init_params(ocfg, seed=2, dtype=torch.float64)
    g = torch.Generator().manual_seed(7)
    # a "trained" generator: non-trivial moving statistics (gen_sampler normalises with them)
    for n in P:
        if n.startswith('generator/') and n.endswith('moving_mean'):
            P[n] = torch.randn(P[n].shape, generator=g, dtype=torch.float64) * 0.3
        if n.startswith('generator/') and n.endswith('moving_variance'):
            P[n] = torch.rand(P[n].shape, generator=g, dtype=torch.float64) + 0.5
    model.store.load_state_dict(P)
    model.build_recover(R, seed=5)
    z = (torch.rand(R * 10, 100, generator=g, dtype=torch.float64) * 2 - 1) * 0.5
    yl = torch.randn(R, 10, generator=g, dtype=torch.float64)
    actual = torch.rand(R, 28, 28, 1, generator=g, dtype=torch.float64)
    model.r_store.load_state_dict({'z_recover': z, 'y_logit_recover': yl})
    return model, P, ocfg, z, yl, actual


def test_kernels_recover_mse_sgd_and_bn_infer_bwd(lib):
    g = torch.Generator().manual_seed(0)
    R, k, npix = 7, 10, 784
    sample = torch.rand(R * k, npix, generator=g); actual = torch.rand(R, npix, generator=g)
    yrec = torch.softmax(torch.randn(R, k, generator=g), -1)
    s64 = sample.double().requires_grad_(True); y64 = yrec.double().requires_grad_(True)
    sq = ((actual.double().unsqueeze(1) - s64.reshape(R, k, npix)) ** 2).mean(-1)
    loss = (sq * y64).sum(-1).mean()
    loss.backward()
    sd, ad, yd = dev(sample), dev(actual), dev(yrec)
    lacc = torch.zeros(1, device='cuda'); sqd = torch.zeros(R, k, device='cuda')
    ds = torch.zeros(R * k, npix, device='cuda'); dy = torch.zeros(R, k, device='cuda')
    call('rcgan_recover_mse', sd.data_ptr(), ad.data_ptr(), yd.data_ptr(), R, k, npix, lacc.data_ptr(), sqd.data_ptr(),
         ds.data_ptr(), dy.data_ptr(), st())
    assert abs(float(lacc) - float(loss)) < 1e-6 * abs(float(loss))
    assert relerr(sqd, sq.detach()) < 1e-6 and relerr(ds, s64.grad) < 1e-6 and relerr(dy, y64.grad) < 1e-6
    # sgd
    p = torch.randn(1000, generator=g); gr = torch.randn(1000, generator=g)
    pd, gd = dev(p), dev(gr)
    call('rcgan_sgd', pd.data_ptr(), gd.data_ptr(), 1000, 500.0, 0.5, st())
    assert relerr(pd, p.double() - 250.0 * gr.double()) < 1e-6
    # inference-mode BN backward: dx = dy * relu'(y) * scale * rsqrt(var + eps), fp32 and bf16, accumulate
    n, hw, c = 6, 9, 64
    x = torch.randn(n * hw, c, generator=g); dyv = torch.randn(n * hw, c, generator=g)
    scale = torch.rand(c, generator=g) + 0.5; offset = torch.randn(c, generator=g)
    mm = torch.randn(c, generator=g) * 0.2; mv = torch.rand(c, generator=g) + 0.5
    for dtype, td, tol in ((_C.F32, torch.float32, 1e-6), (_C.BF16, torch.bfloat16, 1e-2)):
        xq, dyq = x.to(td), dyv.to(td)
        x64 = xq.double().requires_grad_(True)
        y64 = torch.relu((x64 - mm.double()) * torch.rsqrt(mv.double() + 1e-5) * scale.double() + offset.double())
        y64.backward(dyq.double())
        xd, dyd = xq.cuda(), dyq.cuda()
        yd2 = torch.zeros_like(xd); dxd = torch.ones_like(xd)
        save = torch.zeros(2 * c, device='cuda')
        wsb = max(lib.rcgan_bn_workspace(n, hw, c), 2 * c * 4)
        ws = torch.zeros(wsb, dtype=torch.uint8, device='cuda')
        call('rcgan_bn_fwd', xd.data_ptr(), yd2.data_ptr(), n, hw, c, dtype, dtype, keep(dev(scale)), keep(dev(offset)), None, 1e-5,
             _C.ACT_RELU, 0.0, 0, 0.9, keep(dev(mm)), keep(dev(mv)), save.data_ptr(), ws.data_ptr(), wsb, st())
        assert relerr(yd2.float(), y64.detach()) < tol
        call('rcgan_bn_infer_bwd', dyd.data_ptr(), yd2.data_ptr(), c, dxd.data_ptr(), n, hw, c, dtype, keep(dev(scale)), None,
             save.data_ptr(), _C.ACT_RELU, 0.0, 1, ws.data_ptr(), wsb, st())
        assert relerr(dxd.float() - 1.0, x64.grad) < 2 * tol


@pytest.mark.parametrize('precision,tol_loss,tol_grad', [('fp32', 1e-5, 2e-4), ('bf16', 1e-2, 6e-2)])
def test_recover_step_matches_oracle(lib, precision, tol_loss, tol_grad):
    """loss, dL/dz_recover, dL/dy_logit_recover and the variables after one gradient-descent step (lr 500)."""
    R = 6
    model, P, ocfg, z, yl, actual = build(R, precision, graph=False)
    z1, yl1, loss, (dz, dyl) = OM.recover_step(P, z, yl, actual, ocfg, lr=500.0)
    mse = model.recover_step(actual, learning_rate=500.0)
    assert abs(mse - float(loss)) <= tol_loss * abs(float(loss)), (mse, float(loss))
    gz = model.z_recover.grad_torch().double().cpu(); gy = model.y_logit_recover.grad_torch().double().cpu()
    assert relerr(gz, dz) < tol_grad and relerr(gy, dyl) < tol_grad, (relerr(gz, dz), relerr(gy, dyl))
    assert relerr(model.z_recover.torch().double().cpu(), z1) < tol_grad
    assert relerr(model.y_logit_recover.torch().double().cpu(), yl1) < tol_grad
    # generator weights / moving statistics are untouched by the recovery optimiser
    for n, v in model.store.vars.items():
        assert float((v.torch().double().cpu() - P[n].reshape(v.shape)).abs().max()) < 1e-6, n


def test_recover_loop_decreases_loss_and_graph_equals_eager(lib):
    """50 captured-graph steps == 50 eager steps == the oracle trajectory (fp32); the loss decreases (the reference's
    recovery objective is being minimised), zero_one_loss equals the oracle's."""
    R, steps = 4, 50
    m1, P, ocfg, z, yl, actual = build(R, 'fp32', graph=True)
    m2, _, _, _, _, _ = build(R, 'fp32', graph=False)
    zo, ylo = z, yl
    first = last = None
    for i in range(steps):
        a = m1.recover_step(actual if i == 0 else None, learning_rate=20.0)
        b = m2.recover_step(actual if i == 0 else None, learning_rate=20.0)
        zo, ylo, lo, _ = OM.recover_step(P, zo, ylo, actual, ocfg, lr=20.0)
        assert abs(a - b) < 1e-5 * abs(b) and abs(a - float(lo)) < 2e-4 * abs(float(lo)), (i, a, b, float(lo))
        first = a if first is None else first
        last = a
    assert last < first
    y_actual = torch.eye(10)[torch.arange(R) % 10]
    y_rec, mse, zero_one = m1.recover_labels(actual, y_actual, recover_epoch=1, learning_rate=20.0)
    _, yr_o, _ = OM.recover_loss(P, zo, ylo, actual, ocfg)
    # (one more step was taken by recover_labels; compare against the oracle after the same extra step)
    zo, ylo, _, _ = OM.recover_step(P, zo, ylo, actual, ocfg, lr=20.0)
    _, yr_o, _ = OM.recover_loss(P, zo, ylo, actual, ocfg)
    assert relerr(y_rec.double(), yr_o) < 1e-3
    assert abs(zero_one - float(OM.zero_one_loss(y_actual.double(), yr_o))) < 1e-6
