"""End-to-end parity of the CIFAR-10 SN-ResNet RCGAN training steps (BASELINE configs 4-5 at oracle-sized batches):
RCGANCifar.d_step / g_step vs the oracle (oracle/cifar.py) on the same weights, inputs and labels."""
import numpy as np
import pytest
import torch

from oracle import cifar as OC
from robust_conditional_gan_b200.cifar.gan_resnet import RCGANCifar, default_flags
from util import relerr

pytestmark = pytest.mark.gpu


def build(alg, n, precision, dim, perm=True, graph=False, seed=1):
    flags = default_flags(algorithm=alg, alpha=0.5, perm_classifier=perm, perm_multiplier=2.0, confuse_init=(alg == 'rcgan-u'))
    ocfg = OC.default_config(algorithm=alg, alpha=0.5, perm_classifier=perm, perm_multiplier=2.0, confuse_init=(alg == 'rcgan-u'),
                             dim=dim)
    model = RCGANCifar(flags, tower_batch=n, precision=precision, dim=dim, use_cuda_graph=graph)
    P = OC.init_params(ocfg, seed=seed, dtype=torch.float64)
    assert set(P) == set(model.store.vars), set(P) ^ set(model.store.vars)
    # break the symmetric initialisation of the conditional-BN tables and biases so their gradients are exercised
    g = torch.Generator().manual_seed(7)
    for k in P:
        if k.endswith('/scale') or k.endswith('/offset') or k.endswith('/Biases') or k.endswith('/b'):
            P[k] = P[k] + 0.1 * torch.randn(P[k].shape, generator=g, dtype=torch.float64)
    model.store.load_state_dict(P)
    tr = OC.Trainer(P, ocfg)
    batch = OC.synthetic_batch(n, seed=3, dtype=torch.float64)
    return model, tr, batch


def feed_d(model, b):
    n = b['raw'].shape[0]
    model.feed(model.d_prog, all_real_data_int=b['raw'], all_real_labels=b['labels'], all_random_labels=b['labels_random'],
               all_labels_biased=b['labels_biased'], all_labels_inv_weights=b['inv_weights'], noise=b['noise'],
               dequant_noise=torch.zeros(n, 3072))


def feed_g(model, b):
    model.feed(model.g_prog, noise=b['noise_G'], all_random_labels_G=b['labels_random_G'], all_labels_biased_G=b['labels_biased_G'])


class RecordRelu:
    """records every tensor the oracle passes to torch.relu (in call order) while active"""

    def __enter__(self):
        self.rec, self.orig = [], OC.torch.relu
        OC.torch.relu = lambda x: (self.rec.append(x.detach()), self.orig(x))[1]
        return self

    def __exit__(self, *a):
        OC.torch.relu = self.orig


def relu_flips(prog, rec):
    """ReLU branches on which the fp32 product and the fp64 oracle disagree: [(oracle |pre-activation|, numel)].
    A pre-activation within fp32 noise of 0 legitimately takes the other branch and moves the gradient norm by
    ~1/sqrt(numel) (measured: one flip at |pre| = 1.6e-6 in a 786k-element CBN output -> 2e-3 on every upstream gradient)."""
    from robust_conditional_gan_b200 import _C, nnops
    masks = []
    for o in prog.ops:
        if isinstance(o, nnops.BatchNormOp) and o.act == _C.ACT_RELU: masks.append(o.y.torch() > 0)
        elif isinstance(o, nnops.ConvOp) and o.act == _C.ACT_RELU: masks.append(o.y.torch() > 0)
        elif isinstance(o, nnops.ConvOp) and o.y2 is not None: masks.append(o.y2.torch() > 0)   # relu(y) as second output
        elif isinstance(o, nnops.ActOp) and o.act == _C.ACT_RELU: masks.append(o.y.torch() > 0)
        elif isinstance(o, nnops.MeanHWOp) and o.relu: masks.append(o.x.torch() > 0)
    assert len(masks) == len(rec), (len(masks), len(rec))
    flips = []
    for pm, ox in zip(masks, rec):
        mism = pm.cpu() != (ox > 0)
        if int(mism.sum()):
            flips.append((float(ox[mism].abs().max()), ox.numel(), int(mism.sum())))
    return flips


def check_grads_or_flips(vars_, ref_grads, tol, label, prog, rec):
    try:
        check_grads(vars_, ref_grads, tol, label)
    except AssertionError as e:
        flips = relu_flips(prog, rec)
        assert flips and all(f[0] < 1e-4 for f in flips), (str(e), flips)      # every excess must be an identified flip
        check_grads(vars_, ref_grads, 3e-2, label + ' (with %d relu flips)' % len(flips))


def check_grads(vars_, ref_grads, tol, label):
    errs = []
    for v in vars_:
        ref = ref_grads[v.name]
        if float(ref.norm()) < 1e-10:
            continue
        errs.append((relerr(v.grad.reshape(ref.shape), ref), v.name.split('/', 1)[-1]))
    errs.sort(reverse=True)
    assert errs[0][0] < tol, (label, [(n_, '%.1e' % e) for e, n_ in errs[:12]])


def _reorder_d(rec, alg):
    """The oracle runs D once on concat[real; fake] (rcgan / biased / unbiased) while the product runs D(real) then D(fake):
    split each recorded discriminator tensor into its two halves, in the product's op order (G first, D(real), D(fake))."""
    n_g = 7                                      # generator relus (3 blocks x 2 + G.OutputNorm) come first in both
    g, d = rec[:n_g], rec[n_g:]
    if alg == 'rcgan-u':
        return rec                               # oracle already runs D(real), D(fake) separately
    half = [t.shape[0] // 2 for t in d]
    return g + [t[:h] for t, h in zip(d, half)] + [t[h:] for t, h in zip(d, half)]


@pytest.mark.parametrize('alg', ['rcgan', 'rcgan-u', 'biased', 'unbiased'])
def test_fp32_steps_match_oracle(lib, alg):
    """fp32 mode at DIM=32, tower batch 6: D step then G step (iteration 1) -- losses 1e-4, gradients 2e-4."""
    model, tr, b = build(alg, 6, 'fp32', 32)
    feed_d(model, b); feed_g(model, b)
    with RecordRelu() as rr:
        tr.d_step(b, 0)
    model.d_step(0)
    torch.cuda.synchronize()
    got = model.d_prog.loss_dict(model.d_prog.losses.cpu())
    ref = tr.last['d']
    assert abs(got['disc_real_l'] + got['disc_fake_l'] - float(ref['disc_wgan'])) < 1e-4 * max(1, abs(float(ref['disc_wgan'])))
    assert abs(got['perm_classifier_real_loss'] - float(ref['perm_real'])) < 1e-4
    check_grads_or_flips(model.disc_params, tr.last['d_grads'], 2e-4, alg + ' D', model.d_prog, _reorder_d(rr.rec, alg))
    for n_, v in model.store.vars.items():
        if n_.endswith('/u'):
            assert relerr(v.data, tr.P[n_]) < 1e-5, n_
    # Adam(beta1=0) moves every weight by ~lr*sign(g) on its first step, so weights whose gradient is at fp32 noise level
    # land 2*lr apart in the two implementations; re-synchronise the state so the G step is compared like for like
    model.store.load_state_dict({k: v.detach() for k, v in tr.P.items()})
    with RecordRelu() as rr:
        tr.g_step(b, 1)
    model.g_step(1)
    torch.cuda.synchronize()
    got = model.g_prog.loss_dict(model.g_prog.losses.cpu())
    assert abs(got['gen_wgan'] - float(tr.last['g']['gen_wgan'])) < 1e-4 * max(1, abs(float(tr.last['g']['gen_wgan'])))
    assert abs(got['perm_classifier_fake_loss'] - float(tr.last['g']['perm_fake'])) < 1e-4
    check_grads_or_flips(model.gen_params + model.c_params, tr.last['g_grads'], 2e-4, alg + ' G', model.g_prog, rr.rec)
    # NO_OPS: the G step leaves the trunk's u alone but D.Embedding_y's u moves (gan_resnet.py:723-736)
    for n_, v in model.store.vars.items():
        if n_.endswith('/u'):
            assert relerr(v.data, tr.P[n_]) < 1e-5, n_


@pytest.mark.parametrize('alg', ['rcgan', 'rcgan-u'])
def test_bf16_steps_match_oracle_full_width(lib, alg):
    """bf16 mode at the real width (DIM=128: every 3x3 / 1x1 conv on tcgen05), tower batch 4."""
    model, tr, b = build(alg, 4, 'bf16', 128, perm=False)
    feed_d(model, b); feed_g(model, b)
    tr.d_step(b, 0); model.d_step(0)
    torch.cuda.synchronize()
    got = model.d_prog.loss_dict(model.d_prog.losses.cpu())
    assert abs(got['disc_real_l'] + got['disc_fake_l'] - float(tr.last['d']['disc_wgan'])) < 2e-2
    check_grads(model.disc_params, tr.last['d_grads'], 6e-2, alg + ' D bf16')
    model.store.load_state_dict({k: v.detach() for k, v in tr.P.items()})
    tr.g_step(b, 1); model.g_step(1)
    torch.cuda.synchronize()
    got = model.g_prog.loss_dict(model.g_prog.losses.cpu())
    assert abs(got['gen_wgan'] - float(tr.last['g']['gen_wgan'])) < 2e-2
    check_grads(model.gen_params + model.c_params, tr.last['g_grads'], 2e-1, alg + ' G bf16')


def test_iteration_schedule_and_graph(lib):
    """gan_resnet.py:919-947: iteration 0 has no G step, later ones G + 5 D; captured graphs replay."""
    model, tr, b = build('rcgan-u', 4, 'fp32', 32, graph=True)
    d = dict(all_real_data_int=b['raw'], all_real_labels=b['labels'], all_random_labels=b['labels_random'],
             all_labels_biased=b['labels_biased'], all_labels_inv_weights=b['inv_weights'], noise=b['noise'],
             dequant_noise=torch.zeros(4, 3072))
    g = dict(noise=b['noise_G'], all_random_labels_G=b['labels_random_G'], all_labels_biased_G=b['labels_biased_G'])
    out0 = model.train_iteration(d, g)
    assert model.groups['g'].t == 0 and model.groups['d'].t == 5
    out1 = model.train_iteration(d, g)
    assert model.groups['g'].t == 1 and model.groups['c'].t == 1 and model.groups['d'].t == 10
    for it in range(2):
        if it > 0:
            tr.g_step(b, it)
        for _ in range(5):
            tr.d_step(b, it)
    assert abs(out1['disc_real_l'] + out1['disc_fake_l'] - float(tr.last['d']['disc_wgan'])) < 2e-2
    assert abs(out1['gen_wgan'] - float(tr.last['g']['gen_wgan'])) < 2e-2
    assert all(np.isfinite(v) for v in out1.values())


def test_two_layer_perm_classifier(lib):
    """--perm_type 2layer (gan_resnet.py:467-480): SN-Linear 3072 -> 128 -> 10; D step (real images) and G step (generated)."""
    flags = default_flags(algorithm='rcgan-u', alpha=0.5, perm_classifier=True, perm_multiplier=2.0, confuse_init=True, perm_type='2layer')
    ocfg = OC.default_config(algorithm='rcgan-u', alpha=0.5, perm_classifier=True, perm_multiplier=2.0, confuse_init=True, dim=32,
                             perm_type='2layer')
    model = RCGANCifar(flags, tower_batch=6, precision='fp32', dim=32, use_cuda_graph=False)
    P = OC.init_params(ocfg, seed=1, dtype=torch.float64)
    assert set(P) == set(model.store.vars), set(P) ^ set(model.store.vars)
    model.store.load_state_dict(P)
    tr = OC.Trainer(P, ocfg)
    b = OC.synthetic_batch(6, seed=3, dtype=torch.float64)
    feed_d(model, b); feed_g(model, b)
    tr.d_step(b, 0); model.d_step(0)
    torch.cuda.synchronize()
    got = model.d_prog.loss_dict(model.d_prog.losses.cpu())
    assert abs(got['perm_classifier_real_loss'] - float(tr.last['d']['perm_real'])) < 1e-4
    pc = [v for v in model.disc_params if 'perm_classifier' in v.name]
    assert len(pc) == 4
    check_grads(pc, tr.last['d_grads'], 2e-4, '2layer perm classifier')
    model.store.load_state_dict({k: v.detach() for k, v in tr.P.items()})
    tr.g_step(b, 1); model.g_step(1)
    torch.cuda.synchronize()
    got = model.g_prog.loss_dict(model.g_prog.losses.cpu())
    assert abs(got['perm_classifier_fake_loss'] - float(tr.last['g']['perm_fake'])) < 1e-4
    check_grads(model.gen_params + model.c_params, tr.last['g_grads'], 5e-4, '2layer G')


@pytest.mark.parametrize('precision,dim,tol', [('fp32', 32, 2e-3), ('bf16', 128, 2e-1)])
def test_two_towers_in_one_process(lib, precision, dim, tol):
    """The reference's single-GPU graph (DEVICES = [gpu0, gpu0], gan_resnet.py:183-192): two towers of BATCH_SIZE/2 whose costs are
    averaged -- conditional-BN statistics per tower.  One process with towers=2 must equal the oracle's two-tower cost and differ
    from one big tower."""
    n = 4
    flags = default_flags(algorithm='rcgan', alpha=0.5)
    ocfg = OC.default_config(algorithm='rcgan', alpha=0.5, dim=dim)
    model = RCGANCifar(flags, tower_batch=n, precision=precision, dim=dim, use_cuda_graph=False, towers=2)
    P = OC.init_params(ocfg, seed=1, dtype=torch.float64)
    g = torch.Generator().manual_seed(7)
    for k in P:
        if k.endswith('/scale') or k.endswith('/offset'):
            P[k] = P[k] + 0.1 * torch.randn(P[k].shape, generator=g, dtype=torch.float64)
    model.store.load_state_dict(P)
    full = OC.synthetic_batch(2 * n, seed=3, dtype=torch.float64)
    feed_g(model, full)
    model.g_step(1)
    torch.cuda.synchronize()
    tower = lambda r: {k: (v[r * (v.shape[0] // 2):(r + 1) * (v.shape[0] // 2)] if torch.is_tensor(v) else v) for k, v in full.items()}
    tr = OC.Trainer(P, ocfg)
    tr._req(tr.gn)
    cost = sum(OC.gen_cost(tr.P, tower(r), ocfg)[0]['gen_cost'] for r in range(2)) / 2
    ref = dict(zip(tr.gn, torch.autograd.grad(cost, [tr.P[k] for k in tr.gn], allow_unused=True)))
    got = model.g_prog.loss_dict(model.g_prog.losses.cpu())
    assert abs(got['gen_wgan'] - float(cost)) < (1e-4 if precision == 'fp32' else 2e-2)
    ref = {k: (v if v is not None else torch.zeros_like(tr.P[k])) for k, v in ref.items()}
    check_grads(model.gen_params, ref, tol, 'two towers ' + precision)
    if precision == 'fp32':
        tr._req(tr.gn)
        big = OC.gen_cost(tr.P, full, ocfg)[0]['gen_cost']
        gb = dict(zip(tr.gn, torch.autograd.grad(big, [tr.P[k] for k in tr.gn], allow_unused=True)))
        k = 'Generator/G.Block.1.Conv1/Filters'
        assert relerr(gb[k], ref[k]) > 2e-2            # per-tower statistics matter (10x the tolerance above)


def test_generator_sampler(lib):
    """RCGANCifar.sample == the oracle's Generator on the same noise / labels (gan_resnet.py:820-829 fixed_noise_samples)"""
    model, tr, b = build('rcgan', 4, 'fp32', 32, perm=False)
    labels = np.repeat(np.arange(10), 2)
    noise = torch.randn(20, 128, generator=torch.Generator().manual_seed(5), dtype=torch.float64)
    got = model.sample(labels, noise=noise)
    ref = OC.Generator(OC.Ctx(tr.P, False), noise, torch.as_tensor(labels), 32)
    assert got.shape == (20, 32, 32, 3) and relerr(torch.as_tensor(got), ref) < 1e-4


def _calibrated(vars_, got_of, plain, emul, label, slack=2.0, floor=2e-3):
    """product-vs-fp64 error of every variable's gradient against what bf16 STORAGE alone costs (emulated-vs-fp64 oracle)"""
    rows = []
    for v in vars_:
        a, e = plain[v.name], emul[v.name]
        if float(a.norm()) < 1e-10:
            continue
        g = got_of(v).reshape(a.shape)
        rows.append((relerr(g, a), relerr(e, a), relerr(g, e), v.name.split('/', 1)[-1]))
    bad = [r for r in rows if r[0] > slack * r[1] + floor]
    assert not bad, (label, [(n_, 'product-fp64 %.1e' % pa, 'storage-only %.1e' % ea, 'product-emulated %.1e' % pe) for pa, ea, pe, n_ in bad[:8]])
    mean_ratio = sum((r[0] + 1e-3) / (r[1] + 1e-3) for r in rows) / len(rows)
    assert mean_ratio < 1.5, (label, 'mean error ratio product / storage-only', mean_ratio)
    return rows


@pytest.mark.parametrize('alg', ['rcgan', 'rcgan-u'])
def test_bf16_error_is_the_storage_quantisation_gap(lib, alg):
    """What separates a bf16 step from the fp64 oracle must be bf16 STORAGE, not the kernels.  The oracle is run twice: plain
    fp64, and with O.bf16_storage() -- rounding to bf16 at exactly the product's storage points (activations, activation
    gradients, weight packs incl. the folded filters) with fp64 arithmetic in between.
      * losses: product vs the storage-emulating oracle <= 1e-3 (north_star's loss bar; measured 5e-5);
      * generated images: product vs emulated <= 1e-2 (measured 7.5e-3 after 21 stored layers; 2e-5 after the first);
      * every per-variable gradient: (product vs fp64) <= 2 x (emulated vs fp64) + 2e-3 -- a kernel defect of the size VERDICT r1
        worried about (10 % in a bf16-only path) sits 3-5x above the 1.3-3e-2 the discriminator's storage rounding costs.
    Element-exact agreement with the emulation is not attainable end to end: one bf16 rounding that lands on the other side of
    a tie (fp32 vs fp64 accumulation, 1 element in 1e4) perturbs everything downstream by a bf16 ulp, which re-randomises the
    later roundings -- the layer-by-layer growth 2e-5 -> 7.5e-3 is printed by tools/bf16_gap.py (profiles/r2_bf16_gap.txt)."""
    from oracle import nn as O
    model, tr, b = build(alg, 4, 'bf16', 128, perm=(alg == 'rcgan-u'))
    P0 = {k: v.clone() for k, v in tr.P.items()}
    tr_e = OC.Trainer({k: v.clone() for k, v in P0.items()}, tr.cfg)
    feed_d(model, b); feed_g(model, b)
    tr.d_step(b, 0)
    with O.bf16_storage():
        tr_e.d_step(b, 0)
        fake_e = OC.Generator(OC.Ctx(P0, False), b['noise'], b['labels_random'], 128)
    model.d_step(0)
    torch.cuda.synchronize()
    got = model.d_prog.loss_dict(model.d_prog.losses.cpu())
    assert abs(got['disc_real_l'] + got['disc_fake_l'] - float(tr_e.last['d']['disc_wgan'])) < 1e-3
    assert relerr(model.fake_D.torch().float().cpu(), fake_e) < 1e-2
    _calibrated(model.disc_params, lambda v: v.grad, tr.last['d_grads'], tr_e.last['d_grads'], alg + ' D')
    model.store.load_state_dict({k: v.detach() for k, v in tr.P.items()})
    tr_e.P = {k: v.detach().clone() for k, v in tr.P.items()}
    tr.g_step(b, 1)
    with O.bf16_storage():
        tr_e.g_step(b, 1)
    model.g_step(1)
    torch.cuda.synchronize()
    got = model.g_prog.loss_dict(model.g_prog.losses.cpu())
    assert abs(got['gen_wgan'] - float(tr_e.last['g']['gen_wgan'])) < 2e-3
    _calibrated(model.gen_params + model.c_params, lambda v: v.grad, tr.last['g_grads'], tr_e.last['g_grads'], alg + ' G')




def test_bf16_storage_gap_at_a_trained_state(lib):
    """The same calibrated comparison after 40 training iterations of the product itself (bf16, captured graphs): weights,
    spectral-norm vectors and conditional-BN tables are no longer at their symmetric initial values, the hinge losses are partly
    saturated.  The product's state is handed to the oracle (TF variable names) and ONE further D step and G step are compared."""
    from oracle import nn as O
    n = 4
    flags = default_flags(algorithm='rcgan-u', alpha=0.5, perm_classifier=True, perm_multiplier=2.0, confuse_init=True, lr=2e-4)
    ocfg = OC.default_config(algorithm='rcgan-u', alpha=0.5, perm_classifier=True, perm_multiplier=2.0, confuse_init=True, dim=128)
    model = RCGANCifar(flags, tower_batch=n, precision='bf16', dim=128, use_cuda_graph=True)
    model.store.load_state_dict(OC.init_params(ocfg, seed=1, dtype=torch.float64))
    rs = np.random.RandomState(0)
    for it in range(40):
        bt = OC.synthetic_batch(n, seed=100 + it, dtype=torch.float32)
        d = dict(all_real_data_int=bt['raw'], all_real_labels=bt['labels'], all_random_labels=bt['labels_random'],
                 all_labels_biased=bt['labels_biased'], all_labels_inv_weights=bt['inv_weights'], noise=bt['noise'],
                 dequant_noise=torch.as_tensor(rs.rand(n, 3072) / 128, dtype=torch.float32))
        g = dict(noise=bt['noise_G'], all_random_labels_G=bt['labels_random_G'], all_labels_biased_G=bt['labels_biased_G'])
        out = model.train_iteration(d, g)
    assert all(np.isfinite(v) for v in out.values()), out
    model.use_cuda_graph = False
    P = {k: v.double().cpu() for k, v in model.store.state_dict().items()}
    moved = max(float((P[k] - v).abs().max()) for k, v in OC.init_params(ocfg, seed=1, dtype=torch.float64).items() if k.endswith('Filters'))
    assert moved > 1e-3                                              # it did train (Adam: ~lr per step per weight)
    b = OC.synthetic_batch(n, seed=3, dtype=torch.float64)
    tr = OC.Trainer({k: v.clone() for k, v in P.items()}, ocfg)
    tr_e = OC.Trainer({k: v.clone() for k, v in P.items()}, ocfg)
    feed_d(model, b); feed_g(model, b)
    tr.d_step(b, 41)
    with O.bf16_storage():
        tr_e.d_step(b, 41)
    model.d_step(41)
    torch.cuda.synchronize()
    got = model.d_prog.loss_dict(model.d_prog.losses.cpu())
    assert abs(got['disc_real_l'] + got['disc_fake_l'] - float(tr_e.last['d']['disc_wgan'])) < 2e-3
    _calibrated(model.disc_params, lambda v: v.grad, tr.last['d_grads'], tr_e.last['d_grads'], 'trained D')
    model.store.load_state_dict({k: v.detach() for k, v in tr.P.items()})
    tr_e.P = {k: v.detach().clone() for k, v in tr.P.items()}
    tr.g_step(b, 41)
    with O.bf16_storage():
        tr_e.g_step(b, 41)
    model.g_step(41)
    torch.cuda.synchronize()
    got = model.g_prog.loss_dict(model.g_prog.losses.cpu())
    assert abs(got['gen_wgan'] - float(tr_e.last['g']['gen_wgan'])) < 3e-3
    _calibrated(model.gen_params + model.c_params, lambda v: v.grad, tr.last['g_grads'], tr_e.last['g_grads'], 'trained G')
