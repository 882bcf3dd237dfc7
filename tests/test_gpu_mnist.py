"""End-to-end parity of the MNIST DCGAN RCGAN training step (BASELINE configs 1-3 at oracle-sized batches):
the CUDA path through DCGAN.train_iteration vs the oracle Trainer on the same weights, inputs and labels."""
import numpy as np
import pytest
import torch

from oracle import mnist as OM, sampler as OS
from robust_conditional_gan_b200.model import DCGAN, default_flags
from util import relerr

pytestmark = pytest.mark.gpu

RUNS = {
    # run_rcgan.sh / run_rcganu.sh / run_rcgany.sh / run_biased.sh / run_unbiased.sh flag sets
    'rcgan': dict(algorithm='rcgan', disc_type='projection', estimate_confuse=False),
    'rcganu': dict(algorithm='rcgan', disc_type='projection', estimate_confuse=True),
    'rcgany': dict(algorithm='rcgan', disc_type='projection', estimate_confuse=False, concat_y=True, concat_y_layers=[1]),
    'biased': dict(algorithm='biased', disc_type='vanilla', loss_fn='ce', real_match=True, spectral_norm=False,
                   max_norm=False, estimate_confuse=False),
    'ambient': dict(algorithm='ambient', disc_type='vanilla', loss_fn='ce', real_match=True, spectral_norm=False,
                    max_norm=False, estimate_confuse=False),
    'unbiased': dict(algorithm='unbiased', disc_type='projection', estimate_confuse=False),
}


def build(run, B, precision, use_graph=True, pre_norm_fp32=False, seeds=(1, 3)):
    kw = RUNS[run]
    flags = default_flags(batch_size=B, alpha=0.5, perm_regularizer=True, **kw)
    ocfg = OM.default_config(batch_size=B, alpha=0.5, perm_regularizer=True, **kw)
    model = DCGAN(batch_size=B, algorithm=flags.algorithm, estimate_confuse=flags.estimate_confuse, perm_regularizer=True,
                  alpha=0.5, disc_type=flags.disc_type, config=flags, precision=precision, use_cuda_graph=use_graph,
                  pre_norm_fp32=pre_norm_fp32)
    P = OM.init_params(ocfg, seed=seeds[0], dtype=torch.float64)
    assert set(P) == set(model.store.vars), set(P) ^ set(model.store.vars)
    model.store.load_state_dict(P)
    C = OS.one_coin_confusion(0.5)
    tr = OM.Trainer(P, ocfg, C)
    batch = OM.synthetic_batch(B, seed=seeds[1], dtype=torch.float64, C=C, cfg=ocfg)
    return model, tr, batch


def feed(model, batch):
    model.feed(batch_images=batch['x'], batch_z=batch['z'], batch_labels_real=batch['y_real'], batch_labels_gen=batch['y_gen'],
               batch_labels_fake=batch['y_fake'], batch_labels_real_weights=batch['y_real_weights'])


def oracle32_errors(run, B, which, seeds=(1, 3)):
    """What fp32 arithmetic itself achieves: the oracle run in fp32 on CPU vs the fp64 oracle.  The
    product's fp32 error is accepted up to max(1e-4, 4x this): the BN backward cancels catastrophically at
    initialisation and amplifies rounding noise layer after layer (DESIGN.md 'conditioning')."""
    kw = RUNS[run]
    out = {}
    res = {}
    for dt in (torch.float64, torch.float32):
        ocfg = OM.default_config(batch_size=B, alpha=0.5, perm_regularizer=True, **kw)
        P = {k: v.to(dt) for k, v in OM.init_params(ocfg, seed=seeds[0], dtype=torch.float64).items()}
        C = OS.one_coin_confusion(0.5)
        tr = OM.Trainer(P, ocfg, C)
        batch = {k: v.to(dt) for k, v in OM.synthetic_batch(B, seed=seeds[1], dtype=torch.float64, C=C, cfg=ocfg).items()}
        tr.d_step(batch)
        if which == 'g':
            tr.g_step(batch)
        res[dt] = tr.last[which + '_grads']
    for n, ref in res[torch.float64].items():
        if float(ref.norm()) > 1e-12:
            out[n] = relerr(res[torch.float32][n], ref)
    return out


def check_params(model, tr, grads_key, lr_scale):
    """Parameters after Adam.  TF-Adam's first steps are ~lr*g/(|g|+3e-7): elements whose gradient is at fp32 noise
    level legitimately move differently (e.g. biases in front of a batch norm, whose true gradient is exactly 0),
    so: tight bound for all but a tiny fraction of elements, every element inside the Adam trust region."""
    for n, v in model.store.vars.items():
        if n.startswith('discriminator/d_bn') and 'moving' in n:
            continue                            # D's moving stats are never read (not tracked by either side)
        g = tr.last[grads_key].get(n)
        if g is not None and float(g.norm()) < 1e-9:
            continue                            # analytically-zero gradient: the update is pure rounding noise
        diff = (v.data.reshape(-1).double().cpu() - tr.P[n].reshape(-1)).abs()
        info = (n, float(diff.max()), int((diff > 2e-5).sum()), diff.numel())
        assert float(diff.max()) < lr_scale * 2.5 * 2e-4 * (10 if n == 'confusion_logits' else 1), info
        assert float((diff > 2e-5).sum()) <= max(2, 2e-3 * diff.numel()), info


class KinkFlip(AssertionError):
    pass


@pytest.mark.parametrize('run', list(RUNS))
def test_fp32_step_matches_oracle(lib, run):
    """fp32 mode: losses <= 1e-4 relative, per-variable gradients <= 1e-4 relative (or what fp32 itself achieves),
    parameters / u / moving statistics after the D step and after the first G step.

    One documented exception: a LeakyReLU / ReLU pre-activation that lands within fp32 noise of 0 can take the other
    branch than in the fp64 oracle (seen once: |pre-activation| = 1.2e-7 in d_bn1 of the rcganu case, which moves the
    gradient norm by 7e-3 through ONE element).  Such a measure-zero event is retried on other seeds; a real
    discrepancy fails on every seed."""
    last = None
    for seeds in ((1, 3), (2, 5), (4, 9)):
        try:
            _fp32_step(run, seeds)
            return
        except KinkFlip as e:
            last = e
    raise last


def _fp32_step(run, seeds):
    B = 16
    model, tr, batch = build(run, B, 'fp32', use_graph=False, seeds=seeds)
    feed(model, batch)
    # ---- D step
    tr.d_step(batch)
    model.d_step()
    torch.cuda.synchronize()
    got = model.d_prog.loss_dict(model.d_prog.losses.cpu())
    for k in ('d_loss_real', 'd_loss_fake', 'class_loss_real'):
        assert abs(got[k] - float(tr.last['d'][k])) < 1e-4 * max(1.0, abs(float(tr.last['d'][k]))), (k, got[k], tr.last['d'][k])
    cal = oracle32_errors(run, B, 'd', seeds)
    for v in model.d_vars:
        ref = tr.last['d_grads'][v.name]
        if float(ref.norm()) < 1e-12:
            continue
        e = relerr(v.grad.reshape(ref.shape), ref)
        abs_e = float((v.grad.reshape(ref.shape).double().cpu() - ref).norm())
        if not (e < max(1e-4, 4 * cal[v.name]) or abs_e < 5e-7 * ref.numel() ** 0.5):   # abs: sums that cancel (biases)
            raise KinkFlip((v.name, e, cal[v.name], seeds)) if e < 3e-2 else AssertionError((v.name, e, cal[v.name]))
    check_params(model, tr, 'd_grads', 1)
    # ---- first G step (one-step parity from the common post-D-step state)
    tr.g_step(batch)
    model.g_step()
    torch.cuda.synchronize()
    got = model.g_prog.loss_dict(model.g_prog.losses.cpu())
    for k in ('g_loss', 'class_loss_fake'):
        assert abs(got[k] - float(tr.last['g'][k])) < 1e-4 * max(1.0, abs(float(tr.last['g'][k]))), (k, got[k])
    cal = oracle32_errors(run, B, 'g', seeds)
    for v in model.g_vars + model.c_vars:
        ref = tr.last['g_grads'][v.name]
        if float(ref.norm()) < 1e-12:
            continue
        e = relerr(v.grad.reshape(ref.shape), ref)
        abs_e = float((v.grad.reshape(ref.shape).double().cpu() - ref).norm())
        if not (e < max(1e-4, 4 * cal[v.name]) or abs_e < 5e-7 * ref.numel() ** 0.5):
            raise KinkFlip((v.name, e, cal[v.name], seeds)) if e < 3e-2 else AssertionError((v.name, e, cal[v.name]))
    tr.last['all'] = dict(tr.last['d_grads'], **tr.last['g_grads'])
    check_params(model, tr, 'all', 2)
    # ---- second G step on the same z / labels (mnist/model.py:368-371): losses still agree
    tr.g_step(batch)
    model.g_step()
    torch.cuda.synchronize()
    got = model.g_prog.loss_dict(model.g_prog.losses.cpu())
    assert abs(got['g_loss'] - float(tr.last['g']['g_loss'])) < 1e-3


def cosine(a, b):
    a = a.detach().double().cpu().reshape(-1); b = b.detach().double().cpu().reshape(-1)
    return float((a * b).sum() / (a.norm() * b.norm() + 1e-30))


@pytest.mark.parametrize('pre_norm_fp32', [False, True])
@pytest.mark.parametrize('run', ['rcgan', 'rcganu', 'rcgany'])
def test_bf16_step_matches_oracle(lib, run, pre_norm_fp32):
    """bf16 mode (the benchmarked configuration).  Losses <= 1e-2.  Gradients: <= 3e-2 relative for everything that is
    not behind a batch-norm backward, direction (cosine >= 0.97) for the discriminator trunk, whose BN backward at
    this (random-init) point amplifies ANY 4e-3 rounding of its input ~50x -- tests/test_oracle_nn.py shows the fp64
    oracle itself moves by 6-8% when one pre-norm activation is rounded to bf16, so 1e-2 is not attainable there by
    any bf16 implementation (DESIGN.md 'conditioning')."""
    B = 32
    model, tr, batch = build(run, B, 'bf16', use_graph=False, pre_norm_fp32=pre_norm_fp32)
    feed(model, batch)
    tr.d_step(batch)
    model.d_step()
    torch.cuda.synchronize()
    got = model.d_prog.loss_dict(model.d_prog.losses.cpu())
    for k in ('d_loss_real', 'd_loss_fake', 'class_loss_real'):
        assert abs(got[k] - float(tr.last['d'][k])) < 1e-2, (k, got[k], tr.last['d'][k])
    for v in model.d_vars:
        ref = tr.last['d_grads'][v.name]
        if float(ref.norm()) < 1e-9:
            continue
        e, c = relerr(v.grad.reshape(ref.shape), ref), cosine(v.grad, ref)
        head = any(t in v.name for t in ('d_h4_lin', 'd_h5_y_lin', 'classifier', 'bn3/gamma'))
        assert c > 0.97, (v.name, e, c)
        assert e < (5e-2 if head else 0.25), (v.name, e, c)
    tr.g_step(batch)
    model.g_step()
    torch.cuda.synchronize()
    got = model.g_prog.loss_dict(model.g_prog.losses.cpu())
    assert abs(got['g_loss'] - float(tr.last['g']['g_loss'])) < 1e-2
    for v in model.g_vars + model.c_vars:
        ref = tr.last['g_grads'][v.name]
        if float(ref.norm()) < 1e-9:
            continue
        e, c = relerr(v.grad.reshape(ref.shape), ref), cosine(v.grad, ref)
        assert c > 0.95 and e < 0.35, (v.name, e, c)


def test_cuda_graph_iterations_match_eager_and_oracle(lib):
    """3 full iterations (1 D + 2 G each) through captured CUDA graphs == eager launches == oracle trajectory."""
    B = 16
    m1, tr, batch = build('rcganu', B, 'fp32', use_graph=True)
    m2, _, _ = build('rcganu', B, 'fp32', use_graph=False)
    for it in range(3):
        l1 = m1.train_iteration(batch_images=batch['x'], batch_z=batch['z'], batch_labels_real=batch['y_real'],
                                batch_labels_gen=batch['y_gen'], batch_labels_fake=batch['y_fake'],
                                batch_labels_real_weights=batch['y_real_weights'])
        l2 = m2.train_iteration(batch_images=batch['x'], batch_z=batch['z'], batch_labels_real=batch['y_real'],
                                batch_labels_gen=batch['y_gen'], batch_labels_fake=batch['y_fake'],
                                batch_labels_real_weights=batch['y_real_weights'])
        tr.iteration(batch)
        for k in l1:
            assert abs(l1[k] - l2[k]) < 3e-4, (it, k, l1[k], l2[k])
        assert abs(l1['d_loss_real'] - float(tr.last['d']['d_loss_real'])) < 1e-3
        assert abs(l1['g_loss'] - float(tr.last['g']['g_loss'])) < 1e-3
    # graph replay == eager launches up to fp32 atomics ordering (colsum / loss / wgrad reductions), which Adam
    # amplifies only on noise-level gradients
    # (biases in front of a batch norm have an exactly-zero true gradient: TF-Adam turns their rounding noise into +-lr
    # steps, a random walk that the following moving_mean absorbs one to one and training-mode BN cancels exactly)
    for n, v in m1.store.vars.items():
        g = tr.last['g_grads'].get(n, tr.last['d_grads'].get(n))
        if (g is not None and float(g.norm()) < 1e-9) or n.endswith('moving_mean'):
            continue
        diff = (v.data - m2.store.vars[n].data).abs()
        assert float(diff.max()) < 5e-3 and float((diff > 2e-4).sum()) <= max(2, 2e-2 * diff.numel()), (n, float(diff.max()))


def test_gen_sampler_uses_moving_stats(lib):
    B = 8
    model, tr, batch = build('rcgan', B, 'fp32', use_graph=False)
    out = model.sample(batch['z'], batch['y_gen'])
    ref = OM.generator(tr.P, batch['z'], batch['y_gen'], tr.cfg, train=False)
    assert relerr(out, ref) < 1e-4


def test_full_batch_properties(lib):
    """BASELINE config 2 size (B=1024, RCGAN-U, bf16): size-independent properties -- finite losses, hinge bounds,
    softmax rows of the learned confusion matrix sum to 1, max-norm clip respected, u stays unit-norm."""
    B = 1024
    flags = default_flags(batch_size=B, alpha=0.5, **RUNS['rcganu'])
    model = DCGAN(batch_size=B, algorithm='rcgan', estimate_confuse=True, perm_regularizer=True, alpha=0.5,
                  disc_type='projection', config=flags, precision='bf16')
    g = torch.Generator().manual_seed(0)
    lab = lambda: torch.eye(10)[torch.randint(0, 10, (B,), generator=g)]
    for _ in range(3):
        out = model.train_iteration(batch_images=torch.rand(B, 28, 28, 1, generator=g), batch_z=torch.rand(B, 100, generator=g) * 2 - 1,
                                    batch_labels_real=lab(), batch_labels_gen=lab(), batch_labels_fake=lab(),
                                    batch_labels_real_weights=lab())
        assert all(np.isfinite(v) for v in out.values()), out
        assert out['d_loss_real'] >= 0 and out['d_loss_fake'] >= 0
    V = model.store.vars
    for n in ('discriminator/d_h4_lin/Matrix', 'discriminator/d_h5_y_lin/Matrix'):
        assert float(V[n].data.abs().max()) <= 1.0
    for n, v in V.items():
        if n.endswith('spectral_norm/u'):
            assert abs(float(v.data.norm()) - 1.0) < 1e-4, n


def test_train_loop_follows_the_reference_random_stream(lib):
    """DCGAN.train (mnist/model.py:249-372) with --add_noise (run_rcgany.sh): the sample_z draw, the per-epoch re-noising of the
    real / fake labels (:293-333) and every batch_z come from the ONE numpy-legacy stream seeded in load_mnist -- replayed here
    with numpy itself: labels after each epoch's re-noise and every batch_z must be identical."""
    N, B = 200, 32
    rs = np.random.RandomState(0)
    X, y = rs.rand(N, 28, 28, 1).astype(np.float32), rs.randint(10, size=N)
    kw = dict(RUNS['rcgan'])
    flags = default_flags(batch_size=B, alpha=0.3, add_noise=True, noise_alpha=0.2, noise_start=1, noise_end=4, epoch=3,
                          train_size=N, perm_regularizer=True, **kw)
    model = DCGAN(batch_size=B, sample_num=B, algorithm='rcgan', estimate_confuse=False, perm_regularizer=True, alpha=0.3,
                  disc_type='projection', add_noise=True, noise_alpha=0.2, config=flags, precision='fp32', use_cuda_graph=False,
                  data=(X, y))
    seen = []
    model.train_iteration = lambda fetch_losses=False, **f: seen.append({k: torch.as_tensor(v).cpu().numpy().copy() for k, v in f.items()})
    model.train(flags, log_every=10 ** 9)
    # ---- the same sequence with numpy (the reference's own dependency), literally as mnist/model.py does it
    C = OS.one_coin_confusion(0.3)
    ref = OS.mnist_labels_numpy(y, C, real_match=False, seed=547, shuffle=True)      # leaves np.random at the post-sampler state
    sample_z = np.random.uniform(-1, 1, size=(B, 100))
    assert np.array_equal(model.sample_z.cpu().numpy(), sample_z.astype(np.float32))
    y_real, y_fake = ref['y_real'], ref['y_fake']
    it = 0
    for epoch in range(3):
        na, noise_C = model.noise_schedule(epoch)
        y_real, y_fake = OS.mnist_renoise_numpy(y_real, y_fake, noise_C)
        for idx in range(N // B):
            z = np.random.uniform(-1, 1, [B, 100]).astype(np.float32)
            got = seen[it]
            assert np.array_equal(got['batch_z'], z), (epoch, idx)
            assert np.array_equal(got['batch_labels_real'], y_real[idx * B:(idx + 1) * B]), (epoch, idx)
            assert np.array_equal(got['batch_labels_fake'], y_fake[idx * B:(idx + 1) * B]), (epoch, idx)
            assert np.array_equal(got['batch_labels_gen'], ref['y_gen'][idx * B:(idx + 1) * B])
            it += 1
    assert it == len(seen)
    # the schedule itself (:293-321): alpha_start, linear ramp between noise_start and end_epoch, 1.0 afterwards
    a0 = (0.2 - 0.7 / 9) / (0.3 - 0.7 / 9)
    assert abs(model.noise_schedule(0)[0] - a0) < 1e-12 and model.noise_schedule(50)[0] == 1.0
    end = min(4, 1 + (4 - 1) / (0.9 - 0.2) * (0.3 - 0.2))
    assert abs(model.noise_schedule(1)[0] - a0) < 1e-12
    mid = 1 + 0.5 * (end - 1)
    assert model.noise_schedule(int(end) + 1)[0] == 1.0 and a0 < model.noise_schedule(mid)[0] < 1.0


@pytest.mark.parametrize('run', ['rcgan', 'rcganu'])
def test_bf16_error_is_the_storage_quantisation_gap(lib, run):
    """The MNIST counterpart of tests/test_gpu_cifar.py::test_bf16_error_is_the_storage_quantisation_gap: the oracle run plain
    (fp64) and with oracle.nn.bf16_storage() (rounding at the product's bf16 storage points); losses vs the emulation <= 2e-3,
    every per-variable gradient: (product vs fp64) <= 2 x (emulated vs fp64) + 2e-3 (4 x for tensors under 4096 elements, which are
    single draws of the rounding noise) and the mean ratio over the variables < 1.5.  This replaces trust in the loose 25 % /
    35 % bounds of test_bf16_step_matches_oracle: those ARE the storage gap behind three batch-norm backwards, and this test shows
    the kernels add nothing to it."""
    from oracle import nn as O
    B = 32
    model, tr, batch = build(run, B, 'bf16', use_graph=False)
    # the emulation also ARITHMETICS in fp32 like the product (three batch-norm backwards in a row amplify fp32 rounding to the
    # 1e-3 level here: oracle32_errors above), so "storage-only" below means bf16 storage + fp32 arithmetic
    f32 = lambda d: {k: (v.float() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in d.items()}
    tr_e = OM.Trainer(f32({k: v.clone() for k, v in tr.P.items()}), tr.cfg, OS.one_coin_confusion(0.5))
    batch_e = f32(batch)
    feed(model, batch)
    tr.d_step(batch)
    with O.bf16_storage():
        tr_e.d_step(batch_e)
    model.d_step()
    torch.cuda.synchronize()
    got = model.d_prog.loss_dict(model.d_prog.losses.cpu())
    for k in ('d_loss_real', 'd_loss_fake', 'class_loss_real'):
        assert abs(got[k] - float(tr_e.last['d'][k])) < 2e-3, (k, got[k], tr_e.last['d'][k])

    def calibrated(vars_, key, label):
        bad, ratios = [], []
        for v in vars_:
            a, e = tr.last[key][v.name], tr_e.last[key][v.name]
            if float(a.norm()) < 1e-9:
                continue
            g = v.grad.reshape(a.shape)
            pa, ea = relerr(g, a), relerr(e, a)
            ratios.append((pa + 1e-3) / (ea + 1e-3))
            # small tensors (norm tables, the 10x10 confusion logits) are single draws of the rounding noise: wider band
            if pa > (2.0 if a.numel() >= 4096 else 4.0) * ea + 2e-3:
                bad.append((v.name, 'product-fp64 %.1e' % pa, 'storage-only %.1e' % ea, 'product-emulated %.1e' % relerr(g, e)))
        assert not bad, (label, bad[:8])
        assert sum(ratios) / len(ratios) < 1.5, (label, 'mean error ratio product / storage-only', sum(ratios) / len(ratios))
    calibrated(model.d_vars, 'd_grads', run + ' D')
    # common post-D-step state for all three (TF-Adam turns noise-level gradient differences into +-lr parameter differences)
    model.store.load_state_dict({k: v.detach() for k, v in tr.P.items()})
    tr_e.P = f32({k: v.detach().clone() for k, v in tr.P.items()})
    tr.g_step(batch)
    with O.bf16_storage():
        tr_e.g_step(batch_e)
    model.g_step()
    torch.cuda.synchronize()
    got = model.g_prog.loss_dict(model.g_prog.losses.cpu())
    assert abs(got['g_loss'] - float(tr_e.last['g']['g_loss'])) < 2e-3
    calibrated(model.g_vars + model.c_vars, 'g_grads', run + ' G')


def test_bf16_storage_gap_at_a_trained_state(lib):
    """As test_bf16_error_is_the_storage_quantisation_gap, but after 30 training iterations of the product (RCGAN-U, bf16, captured
    graphs): batch-norm tables, spectral-norm vectors, moving statistics and the learned confusion matrix have moved; the product's
    state goes to the oracle by TF variable name and one further D step / G step is compared against the plain and the
    storage-emulating oracle."""
    from oracle import nn as O
    B = 32
    model, tr0, batch0 = build('rcganu', B, 'bf16', use_graph=True)
    C = OS.one_coin_confusion(0.5)
    for it in range(30):
        bt = OM.synthetic_batch(B, seed=50 + it, dtype=torch.float32, C=C, cfg=tr0.cfg)
        out = model.train_iteration(batch_images=bt['x'], batch_z=bt['z'], batch_labels_real=bt['y_real'], batch_labels_gen=bt['y_gen'],
                                    batch_labels_fake=bt['y_fake'], batch_labels_real_weights=bt['y_real_weights'])
    assert all(np.isfinite(v) for v in out.values()), out
    model.use_cuda_graph = False
    P = {k: v.double().cpu() for k, v in model.store.state_dict().items()}
    assert float((P['confusion_logits'] - tr0.P['confusion_logits']).abs().max()) > 1e-3       # the confusion matrix is being learned
    f32 = lambda d: {k: (v.float() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in d.items()}
    tr = OM.Trainer({k: v.clone() for k, v in P.items()}, tr0.cfg, C)
    tr_e = OM.Trainer(f32({k: v.clone() for k, v in P.items()}), tr0.cfg, C)
    batch, batch_e = batch0, f32(batch0)
    feed(model, batch)

    def calibrated(vars_, key, label):
        ratios = []
        for v in vars_:
            a, e = tr.last[key][v.name], tr_e.last[key][v.name]
            if float(a.norm()) < 1e-9:
                continue
            g = v.grad.reshape(a.shape)
            pa, ea = relerr(g, a), relerr(e, a)
            ratios.append((pa + 1e-3) / (ea + 1e-3))
            assert pa <= (2.0 if a.numel() >= 4096 else 4.0) * ea + 2e-3, (label, v.name, pa, ea)
        assert sum(ratios) / len(ratios) < 1.5, (label, sum(ratios) / len(ratios))
    tr.d_step(batch)
    with O.bf16_storage():
        tr_e.d_step(batch_e)
    model.d_step()
    torch.cuda.synchronize()
    got = model.d_prog.loss_dict(model.d_prog.losses.cpu())
    for k in ('d_loss_real', 'd_loss_fake'):
        assert abs(got[k] - float(tr_e.last['d'][k])) < 3e-3, (k, got[k], tr_e.last['d'][k])
    calibrated(model.d_vars, 'd_grads', 'trained D')
    model.store.load_state_dict({k: v.detach() for k, v in tr.P.items()})
    tr_e.P = f32({k: v.detach().clone() for k, v in tr.P.items()})
    tr.g_step(batch)
    with O.bf16_storage():
        tr_e.g_step(batch_e)
    model.g_step()
    torch.cuda.synchronize()
    got = model.g_prog.loss_dict(model.g_prog.losses.cpu())
    assert abs(got['g_loss'] - float(tr_e.last['g']['g_loss'])) < 3e-3
    calibrated(model.g_vars + model.c_vars, 'g_grads', 'trained G')
