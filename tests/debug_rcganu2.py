import sys
import torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from test_gpu_mnist import build, feed
from util import relerr
from oracle import mnist as OM, nn as O
from robust_conditional_gan_b200 import nnops
for run in ('rcgan', 'rcganu'):
    B = 16
    model, tr, batch = build(run, B, 'fp32', use_graph=False)
    feed(model, batch)
    tr.d_step(batch); model.d_step()
    torch.cuda.synchronize()
    model.store.load_state_dict({k: v.detach() for k, v in tr.P.items()})
    P = {k: v.detach().clone() for k, v in tr.P.items()}
    model._g_body_a()
    torch.cuda.synchronize()
    prog = model.g_prog
    cfg = tr.cfg
    # oracle side, collapsed formulation with explicit intermediates
    G = OM.generator(P, batch['z'], batch['y_gen'], cfg, True, None).detach().requires_grad_(True)
    n = 'discriminator/'
    sn = cfg.spectral_norm
    h0 = O.lrelu(OM._conv(P, n + 'd_h0_conv', G, sn, None)); h0.retain_grad()
    c1 = OM._conv(P, n + 'd_h1_conv', h0, sn, None); c1.retain_grad()
    h1 = O.lrelu(OM._bn(P, n + 'd_bn1', c1, True, None)); h1.retain_grad()
    c2 = OM._conv(P, n + 'd_h2_conv', h1, sn, None); c2.retain_grad()
    h2 = O.lrelu(OM._bn(P, n + 'd_bn2', c2, True, None)); h2.retain_grad()
    c3 = OM._conv(P, n + 'd_h3_conv', h2, sn, None); c3.retain_grad()
    h3s = O.lrelu(OM._bn(P, n + 'd_bn3', c3, True, None)); h3s.retain_grad()
    h3 = h3s.mean(dim=(1, 2)); h3.retain_grad()
    psi = OM._lin(P, n + 'd_h4_lin', h3)
    V = P[n + 'd_h5_y_lin/Matrix'] + P[n + 'd_h5_y_lin/bias']
    la = psi + h3 @ V.t()
    if run == 'rcganu':
        w = batch['y_gen'] @ torch.softmax(P['confusion_logits'], -1)
    else:
        w = batch['y_fake']
    gl = (-(la) * w).sum(1).mean()
    cl = O.sigmoid_ce(OM.classifier(P, G), batch['y_gen']).mean()
    (gl + cfg.perm_multiplier * cl).backward()
    ops = prog.ops
    convs = [o for o in ops if isinstance(o, nnops.ConvOp) and len(o.y.shape) == 4]
    bns = [o for o in ops if isinstance(o, nnops.BatchNormOp) and o.y.shape[-1] == 64]
    mean = [o for o in ops if isinstance(o, nnops.MeanHWOp)][0]
    loss = [o for o in ops if isinstance(o, nnops.ChannelLossOp)][0]
    print('==', run)
    print('wgt', relerr(loss.wgt.torch(), w), 'logits', relerr(loss.logits.torch(), la))
    print('G fwd', relerr(model.G_gstep.torch(), G), 'G grad', relerr(model.G_gstep.grad_torch(), G.grad))
    for name, t, ref in (('h0', convs[0].y, h0), ('c1', convs[1].y, c1), ('h1', bns[0].y, h1), ('c2', convs[2].y, c2),
                         ('h2', bns[1].y, h2), ('c3', convs[3].y, c3), ('h3s', bns[2].y, h3s), ('h3', mean.y, h3)):
        print('%-4s fwd %.2e  grad %.2e' % (name, relerr(t.torch(), ref), relerr(t.grad_torch(), ref.grad)))
