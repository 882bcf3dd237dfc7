"""End-to-end parity at the BENCHMARK sizes (BASELINE configs[1] and configs[3]): the kernels bench.py times -- the persistent
tcgen05 conv variants, the tcgen05 wgrad at its full contraction length, batch norm / conditional batch norm over the full batch
-- inside whole training steps, against the fp32 oracle on the same weights and inputs.  The oracle needs ~3 s (MNIST, B = 1024)
and ~1 min (CIFAR, tower batch 256: D step + G step at 512 images) on the box's host cores.

Tolerances are the bf16 ones of tests/test_gpu_mnist.py / test_gpu_cifar.py (the benchmark precision); what is new here is the
SIZE: every conv crosses the persistent kernels' dispatch thresholds, which each test asserts from the library's variant log."""
import pytest
import torch

from oracle import cifar as OC
from robust_conditional_gan_b200 import _C
from util import relerr

import test_gpu_cifar as TC
import test_gpu_mnist as TM

pytestmark = pytest.mark.gpu


# biases in front of a batch norm: their true gradient is exactly 0 (the norm subtracts the batch mean); in fp32 both sides hold
# rounding noise of arbitrary direction (the fp64 oracle of the small-batch tests returns < 1e-9 and they are skipped there too)
ZERO_GRAD = ('d_h1_conv/biases', 'd_h2_conv/biases', 'd_h3_conv/biases', 'g_h0_lin/bias', 'g_h1_lin/bias', 'g_h2/biases')


def cosine(a, b):
    a = a.detach().double().cpu().reshape(-1); b = b.detach().double().cpu().reshape(-1)
    return float((a * b).sum() / (a.norm() * b.norm() + 1e-30))


def test_mnist_rcganu_b1024_step_matches_oracle(lib):
    """BASELINE configs[1]: MNIST DCGAN RCGAN-U (learned confusion matrix + permutation regulariser), B = 1024, bf16: one D step and
    one G(+C) step, losses 1e-2, per-variable gradients by direction (cos >= 0.97) and the bf16 bounds of the B = 32 test."""
    B = 1024
    torch.set_num_threads(max(1, torch.get_num_threads()))
    model, tr, batch = TM.build('rcganu', B, 'bf16', use_graph=False)
    P32 = {k: v.float() for k, v in tr.P.items()}
    tr = TM.OM.Trainer(P32, tr.cfg, TM.OS.one_coin_confusion(0.5))          # fp32 oracle: 10 label-wise D calls at B = 1024
    b32 = {k: (v.float() if v.is_floating_point() else v) for k, v in batch.items()}
    TM.feed(model, batch)
    _C.conv_variant_log(reset=True)
    tr.d_step(b32)
    model.d_step()
    torch.cuda.synchronize()
    got = model.d_prog.loss_dict(model.d_prog.losses.cpu())
    for k in ('d_loss_real', 'd_loss_fake', 'class_loss_real'):
        assert abs(got[k] - float(tr.last['d'][k])) < 1e-2, (k, got[k], tr.last['d'][k])
    worst = []
    for v in model.d_vars:
        ref = tr.last['d_grads'][v.name]
        if float(ref.norm()) < 1e-9:
            continue
        if v.name.endswith(ZERO_GRAD):
            assert float(ref.norm()) < 1e-4 and float(v.grad.norm()) < 1e-2, (v.name, float(ref.norm()), float(v.grad.norm()))
            continue
        e, c = relerr(v.grad.reshape(ref.shape), ref), cosine(v.grad, ref)
        worst.append((e, c, v.name))
        head = any(t in v.name for t in ('d_h4_lin', 'd_h5_y_lin', 'classifier', 'bn3/gamma'))
        assert c > 0.97 and e < (5e-2 if head else 0.25), (v.name, e, c)
    tr.g_step(b32)
    model.g_step()
    torch.cuda.synchronize()
    got = model.g_prog.loss_dict(model.g_prog.losses.cpu())
    assert abs(got['g_loss'] - float(tr.last['g']['g_loss'])) < 1e-2
    assert abs(got['class_loss_fake'] - float(tr.last['g']['class_loss_fake'])) < 1e-2
    for v in model.g_vars + model.c_vars:
        ref = tr.last['g_grads'][v.name]
        if float(ref.norm()) < 1e-9:
            continue
        if v.name.endswith(ZERO_GRAD):
            assert float(ref.norm()) < 1e-4 and float(v.grad.norm()) < 1e-2, (v.name, float(ref.norm()), float(v.grad.norm()))
            continue
        e, c = relerr(v.grad.reshape(ref.shape), ref), cosine(v.grad, ref)
        worst.append((e, c, v.name))
        assert c > 0.95 and e < 0.35, (v.name, e, c)
    log = _C.conv_variant_log()
    print('variants', sorted(log)); print('worst', sorted(worst, reverse=True)[:6])
    # g_h2 / d_h1 dgrad, the roofline kernel: the CTA-pair kernel, or the one-CTA persistent kernel under RCGAN_TC_PAIR=0
    assert any(v.startswith(('conv_tc_pair<128,2,4,bf16,multi=1>', 'conv_tc_persist<128,2,3,bf16,multi=1>')) for v in log), log
    assert any(v.startswith(('conv_tc_pair<', 'conv_tc_persist<')) and 'multi=0' in v for v in log), log
    assert any(v.startswith('wgrad_tc<') for v in log), log


def test_cifar_rcgan_b256_steps_match_oracle(lib):
    """BASELINE configs[3]: CIFAR-10 SN-ResNet RCGAN, tower batch 256, bf16, DIM = 128: D step (256 real + 256 generated images)
    and G step (512 generated), losses 2e-2, per-variable gradients 6e-2 (D) / 2e-1 (G, behind 7 conditional-BN backwards)."""
    n = 256
    model, tr, b = TC.build('rcgan', n, 'bf16', 128, perm=False)
    tr = OC.Trainer({k: v.float() for k, v in tr.P.items()}, tr.cfg)
    b32 = {k: (v.float() if v.is_floating_point() else v) for k, v in b.items()}
    TC.feed_d(model, b); TC.feed_g(model, b)
    _C.conv_variant_log(reset=True)
    tr.d_step(b32, 0); model.d_step(0)
    torch.cuda.synchronize()
    got = model.d_prog.loss_dict(model.d_prog.losses.cpu())
    assert abs(got['disc_real_l'] + got['disc_fake_l'] - float(tr.last['d']['disc_wgan'])) < 2e-2
    TC.check_grads(model.disc_params, tr.last['d_grads'], 6e-2, 'rcgan D bf16 n=256')
    dlog = _C.conv_variant_log()
    model.store.load_state_dict({k: v.detach() for k, v in tr.P.items()})
    tr.g_step(b32, 1); model.g_step(1)
    torch.cuda.synchronize()
    got = model.g_prog.loss_dict(model.g_prog.losses.cpu())
    assert abs(got['gen_wgan'] - float(tr.last['g']['gen_wgan'])) < 2e-2
    # every generator conv bias but G.Output's shifts a channel by a constant that the following conditional batch norms remove
    # (a 1x1 shortcut / upsampling keeps it constant): the true gradient is exactly 0, the fp32 oracle holds rounding noise there
    # (the fp64 oracle of tests/test_gpu_cifar.py returns < 1e-10 and check_grads skips them)
    zero = [v for v in model.gen_params if v.name.endswith('/Biases') and 'G.Output' not in v.name]
    for v in zero:
        wg = tr.last['g_grads'][v.name.replace('/Biases', '/Filters')]
        assert float(tr.last['g_grads'][v.name].norm()) < 1e-3 * float(wg.norm()), v.name
        assert float(v.grad.norm()) < 2e-2 * float(wg.norm()), (v.name, float(v.grad.norm()), float(wg.norm()))
    TC.check_grads([v for v in model.gen_params + model.c_params if v not in zero], tr.last['g_grads'], 2e-1, 'rcgan G bf16 n=256')
    glog = _C.conv_variant_log()
    print('D step variants', sorted(dlog)); print('G step variants', sorted(glog))
    for log in (dlog, glog):
        # the 256-channel generator convs and the 128-channel discriminator convs: CTA-pair kernels (one-CTA under RCGAN_TC_PAIR=0)
        assert any(v.startswith(('conv_tc_pair<256,1,4,bf16', 'conv_tc_persist<256,1,3,bf16')) for v in log), log
        assert any(v.startswith(('conv_tc_pair<128,2,4,bf16', 'conv_tc_persist<128,2,3,bf16')) for v in log), log
        assert any(v.startswith(('wgrad_tc_pair<', 'wgrad_tc<128')) for v in log), log
