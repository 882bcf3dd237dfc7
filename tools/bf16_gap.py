"""Diagnostic: per-variable relative gradient differences between (product bf16, oracle fp64, oracle with bf16 storage emulation)
for one CIFAR D step + G step at tower batch n, DIM 128.  usage: python tools/bf16_gap.py [n]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
from oracle import cifar as OC, nn as O
import test_gpu_cifar as TC
from util import relerr

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
model, tr, b = TC.build('rcgan', n, 'bf16', 128, perm=False)
P0 = {k: v.clone() for k, v in tr.P.items()}
TC.feed_d(model, b); TC.feed_g(model, b)
tr_e = OC.Trainer({k: v.clone() for k, v in P0.items()}, tr.cfg)
tr.d_step(b, 0)
with O.bf16_storage():
    tr_e.d_step(b, 0)
model.d_step(0)
torch.cuda.synchronize()
print('loss product %.6f plain %.6f emul %.6f' % (sum(model.d_prog.loss_dict(model.d_prog.losses.cpu()).values()),
                                                    float(tr.last['d']['disc_wgan']), float(tr_e.last['d']['disc_wgan'])))
# forward agreement: generated images and discriminator features
nctx = OC.Ctx(P0, False)
fake_p = OC.Generator(nctx, b['noise'], b['labels_random'], 128)
with O.bf16_storage():
    fake_e = OC.Generator(nctx, b['noise'], b['labels_random'], 128)
    h_e, _ = OC.Discriminator(OC.Ctx(P0, False), torch.cat([b['real'], fake_e], 0), 128)
h_p, _ = OC.Discriminator(OC.Ctx(P0, False), torch.cat([b['real'], fake_p], 0), 128)
# layer by layer: the emulation's stored tensors (in order) against the product's op outputs of the same shape (in order)
O._EMUL['trace'] = []
with O.bf16_storage():
    OC.Generator(nctx, b['noise'], b['labels_random'], 128)
trace, O._EMUL['trace'] = O._EMUL['trace'], None
outs = [(type(op).__name__, o) for op in model.d_prog.ops for o in op.outputs if o.dtype == 1 and len(o.shape) >= 2]
j = 0
for i, t in enumerate(trace):
    for k in range(j, min(j + 4, len(outs))):
        o = outs[k][1]
        if o.rows * o.c == t.numel():
            got = o.torch().float().cpu().reshape(t.shape)
            print('  stored #%2d %-18s %-12s relerr %.2e  maxabs %.2e' % (i, tuple(t.shape), outs[k][0], relerr(got, t),
                                                                        float((got.double() - t).abs().max())))
            j = k + 1
            break
    else:
        print('  stored #%2d %-18s (no product tensor: lives in an epilogue)' % (i, tuple(t.shape)))
fk = model.fake_D.torch().float().cpu()
print('fake   prod-plain %.2e prod-emul %.2e emul-plain %.2e  (max abs prod-emul %.2e)' % (
    relerr(fk, fake_p), relerr(fk, fake_e), relerr(fake_e, fake_p), float((fk.double() - fake_e).abs().max())))
hh = model.h_all.torch().float().cpu()
print('h real prod-plain %.2e prod-emul %.2e ; h fake prod-plain %.2e prod-emul %.2e' % (
    relerr(hh[:n], h_p[:n]), relerr(hh[:n], h_e[:n]), relerr(hh[n:], h_p[n:]), relerr(hh[n:], h_e[n:])))
print('%-44s %10s %10s %10s' % ('D variable', 'prod-plain', 'prod-emul', 'emul-plain'))
for v in model.disc_params:
    a, e = tr.last['d_grads'][v.name], tr_e.last['d_grads'][v.name]
    if float(a.norm()) < 1e-10:
        continue
    g = v.grad.reshape(a.shape)
    print('%-44s %10.2e %10.2e %10.2e' % (v.name.split('/', 1)[-1], relerr(g, a), relerr(g, e), relerr(e, a)))
model.store.load_state_dict({k: v.detach() for k, v in tr.P.items()})
tr_e.P = {k: v.detach().clone() for k, v in tr.P.items()}
tr.g_step(b, 1)
with O.bf16_storage():
    tr_e.g_step(b, 1)
model.g_step(1)
torch.cuda.synchronize()
print('%-44s %10s %10s %10s' % ('G variable', 'prod-plain', 'prod-emul', 'emul-plain'))
for v in model.gen_params:
    a, e = tr.last['g_grads'][v.name], tr_e.last['g_grads'][v.name]
    if float(a.norm()) < 1e-10:
        continue
    g = v.grad.reshape(a.shape)
    print('%-44s %10.2e %10.2e %10.2e' % (v.name.split('/', 1)[-1], relerr(g, a), relerr(g, e), relerr(e, a)))
