import sys
import torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from test_gpu_mnist import build, feed
from util import relerr
from robust_conditional_gan_b200 import nnops
for run in ('rcgan', 'rcganu'):
    B = 16
    model, tr, batch = build(run, B, 'fp32', use_graph=False)
    feed(model, batch)
    tr.d_step(batch); model.d_step()
    model._g_body_a()
    torch.cuda.synchronize()
    prog = model.g_prog
    bns = [o for o in prog.ops if isinstance(o, nnops.BatchNormOp) and o.y.shape[-1] == 64]
    print('==', run, 'ws bytes', prog.ws.bytes)
    for i, bn in enumerate(bns):
        x = bn.x.torch().double().reshape(-1, 64); y = bn.y.torch().double().reshape(-1, 64)
        dy = bn.y.grad_torch().double().reshape(-1, 64); dx = bn.x.grad_torch().double().reshape(-1, 64)
        gamma = bn.scale.torch().double().reshape(-1)
        mean = x.mean(0); var = ((x - mean) ** 2).mean(0); istd = torch.rsqrt(var + 1e-5)
        save = bn.save.double()
        g = dy * torch.where(y > 0, torch.ones_like(y), torch.full_like(y, 0.2))
        xh = (x - mean) * istd
        A = (gamma * g).mean(0); Bc = (gamma * g * xh).mean(0)
        exp = istd * (gamma * g - A - xh * Bc)
        print('bn%d rows %d: save mean %.1e istd %.1e | dx vs formula %.2e | |A|/|g| %.2e |B| %.2e acc_x %d need %s' % (
            i + 1, x.shape[0], relerr(save[:64], mean), relerr(save[64:], istd), relerr(dx, exp),
            float(A.norm() / g.abs().mean(0).norm()), float(Bc.norm()), bn.acc_x, bn.need))
