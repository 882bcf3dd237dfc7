"""Oracle restatement of the CIFAR-10 SN-ResNet RCGAN graph and training step
(test infrastructure only -- see oracle/__init__.py; parity unpinned for floats).

Follows cifar10/gan_resnet.py:199-421 (networks), :458-490 (perm classifier), :498-817 (losses, optimizers) and
cifar10/common/ops/{conv2d,linear,normalization,embedding,sn}.py.  One tower (the reference's per-device graph);
the product maps towers to ranks.  Parameters live in a flat dict keyed by the reference's TF variable names.
"""
from types import SimpleNamespace

import numpy as np
import torch

from . import nn as O

DIM_G = DIM_D = Z_DIM = 128
VOCAB, EMB = 10, 300


def default_config(**kw):
    """cifar10/gan_resnet.py:40-76 flag defaults + :140-176 constants (hot-path subset)."""
    cfg = SimpleNamespace(algorithm='rcgan', alpha=0.8, batch_size=64, lr=2e-4, confuse_init=False, confuse_init_diag=0.2,
                          confuse_multiplier=1.0, confuse_lr_decay=False, perm_classifier=False, perm_multiplier=1.0,
                          niters=50000, n_critic=5, gen_bs_multiple=2, dim=128)
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def _uniform(g, stdev, shape, dtype):
    lim = stdev * np.sqrt(3)
    return ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * lim).to(dtype)


def init_params(cfg, seed=0, dtype=torch.float64):
    """Shapes / names / initialisers of SURVEY 8a + appendix C (conv2d.py:83-127, linear.py:53-80, embedding.py:29-43,
    normalization.py:49-52, sn.py:35)."""
    g = torch.Generator().manual_seed(seed)
    P = {}
    dim = cfg.dim

    def conv(name, k, cin, cout, he, sn):
        fan_in, fan_out = cin * k * k, cout * k * k
        std = np.sqrt((4. if he else 2.) / (fan_in + fan_out))
        P[name + '/Filters'] = _uniform(g, std, (k, k, cin, cout), dtype)
        P[name + '/Biases'] = torch.zeros(cout, dtype=dtype)
        if sn:
            t = torch.empty((1, cout), dtype=torch.float64)
            torch.nn.init.trunc_normal_(t, 0, 1, -2, 2, generator=g)
            P[name + '/filters/spectral_norm/u'] = t.to(dtype)

    def lin(name, cin, cout, sn):
        P[name + '/W'] = _uniform(g, np.sqrt(2. / (cin + cout)), (cin, cout), dtype)
        P[name + '/b'] = torch.zeros(cout, dtype=dtype)
        if sn:
            t = torch.empty((1, cout), dtype=torch.float64)
            torch.nn.init.trunc_normal_(t, 0, 1, -2, 2, generator=g)
            P[name + '/spectral_norm/u'] = t.to(dtype)

    def cbn(name, c):
        P[name + '/CondBatchNorm/offset'] = torch.zeros(VOCAB, c, dtype=dtype)
        P[name + '/CondBatchNorm/scale'] = torch.ones(VOCAB, c, dtype=dtype)

    G = 'Generator/'
    lin(G + 'G.Input', Z_DIM, 4 * 4 * dim * 8, False)
    for i, (ci, co) in enumerate([(dim * 8, dim * 2), (dim * 2, dim * 2), (dim * 2, dim * 2)], 1):
        b = G + 'G.Block.%d' % i
        conv(b + '.Shortcut', 1, ci, co, False, False)
        cbn(b + '.N1', ci)
        conv(b + '.Conv1', 3, ci, co, True, False)
        cbn(b + '.N2', co)
        conv(b + '.Conv2', 3, co, co, True, False)
    cbn(G + 'G.OutputNorm', dim * 2)
    conv(G + 'G.Output', 3, dim * 2, 3, False, False)
    D = 'Discriminator/'
    conv(D + 'D.Block.1.Shortcut', 1, 3, dim, False, True)
    conv(D + 'D.Block.1.Conv1', 3, 3, dim, True, True)
    conv(D + 'D.Block.1.Conv2', 3, dim, dim, True, True)
    conv(D + 'D.Block.2.Shortcut', 1, dim, dim, False, True)
    for i in range(2, 7):
        conv(D + 'D.Block.%d.Conv1' % i, 3, dim, dim, True, True)
        conv(D + 'D.Block.%d.Conv2' % i, 3, dim, dim, True, True)
    lin(D + 'D.Output', dim, 1, True)
    P[D + 'Embedding.Label/embedding_map'] = ((torch.rand((VOCAB, EMB), generator=g, dtype=torch.float64) * 2 - 1) * 0.08).to(dtype)
    lin(D + 'D.Embedding_y', EMB, dim, True)
    if cfg.perm_classifier:
        if getattr(cfg, 'perm_type', 'linear') == '2layer':          # gan_resnet.py:467-480
            lin(D + 'D.d_perm_classifier_h1', 3072, 128, True)
            lin(D + 'D.d_perm_classifier_h2', 128, VOCAB, True)
        else:
            lin(D + 'D.d_perm_classifier_h1', 3072, VOCAB, True)
    if cfg.algorithm == 'rcgan-u':
        if cfg.confuse_init:
            aa = 7.0 if cfg.confuse_init_diag > 0.99 else np.log(VOCAB * cfg.confuse_init_diag / (1. - cfg.confuse_init_diag))
            aa = min(7.0, aa)
            ci = (0 - aa / VOCAB) * np.ones([VOCAB, VOCAB], dtype=np.float32)
            np.fill_diagonal(ci, aa - aa / VOCAB)
            P['confusion_logits'] = torch.as_tensor(ci, dtype=dtype)
        else:
            P['confusion_logits'] = _uniform(g, np.sqrt(2. / (2 * VOCAB)), (VOCAB, VOCAB), dtype)
    return P


class Ctx:
    """update_collection semantics: u_new is recorded unless NO_OPS (sn.py:51-71)."""

    def __init__(self, P, update=True):
        self.P, self.update, self.u_new = P, update, {}


def _w(ctx, name, key, sn, uname):
    w = ctx.P[name + key]
    if sn:
        w, u_new, _ = O.spectral_normed_weight(w, ctx.P[name + uname])
        if ctx.update:
            ctx.u_new[name + uname] = u_new
    return w


def Conv2D(ctx, x, name, sn, fold=None):
    """conv2d.py:169-216: SN under scope `filters`, stride-1 SAME conv, + Biases.
    Under O.bf16_storage() the weight is rounded as the product's tensor-core pack is; fold ('pool' | 'up', emulation only): the
    caller's 2x resampling folded into a 4x4 stride-2 filter exactly as the product does (O.fold4) -- the fold is computed from
    the fp32 weight and the FOLDED filter is what gets rounded."""
    w = _w(ctx, name, '/Filters', sn, '/filters/spectral_norm/u')
    b = ctx.P[name + '/Biases']
    if fold == 'pool':
        return O.conv2d(x, O.qw(O.fold4(w, 0)), 2) + b
    if fold == 'up':
        return O.conv2d_transpose(x, O.qw(O.fold4(w, 1)), (2 * x.shape[1], 2 * x.shape[2]), 2) + b
    return O.conv2d(x, O.qw(w), 1) + b


def Linear(ctx, x, name, sn, tc=False):
    """linear.py:161-180.  tc: a linear the product runs on the tensor cores (bf16 weight pack) in bf16 mode."""
    w = _w(ctx, name, '/W', sn, '/spectral_norm/u')
    return x @ (O.qw(w) if tc else w) + ctx.P[name + '/b']


def mean_pool(x):
    """gan_resnet.py:239-240."""
    return (x[:, ::2, ::2] + x[:, 1::2, ::2] + x[:, ::2, 1::2] + x[:, 1::2, 1::2]) / 4.


def upsample(x):
    """gan_resnet.py:263-264: concat x4 + depth_to_space == nearest neighbour."""
    return x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)


def Normalize(ctx, name, x, labels):
    """gan_resnet.py:207-228: conditional BN in G, identity in D."""
    if 'G.' in name:
        return O.cond_batchnorm(x, labels, ctx.P[name + '/CondBatchNorm/offset'], ctx.P[name + '/CondBatchNorm/scale'])
    return x


def ResidualBlock(ctx, x, name, sn, resample, labels, cin, cout):
    """gan_resnet.py:275-328."""
    if O.emulating():
        return _ResidualBlock_bf16(ctx, x, name, sn, resample, labels, cin, cout)
    if resample == 'down':
        conv1 = lambda t: Conv2D(ctx, t, name + '.Conv1', sn)
        conv2 = lambda t: mean_pool(Conv2D(ctx, t, name + '.Conv2', sn))
        short = lambda t: mean_pool(Conv2D(ctx, t, name + '.Shortcut', sn))
    elif resample == 'up':
        conv1 = lambda t: Conv2D(ctx, upsample(t), name + '.Conv1', sn)
        conv2 = lambda t: Conv2D(ctx, t, name + '.Conv2', sn)
        short = lambda t: Conv2D(ctx, upsample(t), name + '.Shortcut', sn)
    else:
        conv1 = lambda t: Conv2D(ctx, t, name + '.Conv1', sn)
        conv2 = lambda t: Conv2D(ctx, t, name + '.Conv2', sn)
        short = lambda t: Conv2D(ctx, t, name + '.Shortcut', sn)
    shortcut = x if (cin == cout and resample is None) else short(x)
    out = torch.relu(Normalize(ctx, name + '.N1', x, labels))
    out = conv1(out)
    out = torch.relu(Normalize(ctx, name + '.N2', out, labels))
    out = conv2(out)
    return shortcut + out


def _ResidualBlock_bf16(ctx, x, name, sn, resample, labels, cin, cout):
    """The same block with the product's dataflow and bf16 storage points (robust_conditional_gan_b200/cifar/gan_resnet.py
    ResidualBlock): 1x1 shortcuts on the small grid, ConvMeanPool / UpsampleConv as folded 4x4 stride-2 filters, the shortcut added
    in Conv2's epilogue (the conv result is rounded when it is staged, the sum once more).  Mathematically identical to the block
    above; only the rounding points differ from a plain cast-everything-to-bf16."""
    q = O.q
    if resample == 'up':
        short = q(Conv2D(ctx, x, name + '.Shortcut', sn))                      # on the small grid; upsampled inside the add
    elif resample == 'down':
        short = q(Conv2D(ctx, q(mean_pool(x)), name + '.Shortcut', sn))
    else:
        short = x if cin == cout else q(Conv2D(ctx, x, name + '.Shortcut', sn))
    out = q(torch.relu(Normalize(ctx, name + '.N1', x, labels)))
    out = Conv2D(ctx, out, name + '.Conv1', sn, fold='up' if resample == 'up' else None)
    out = q(torch.relu(Normalize(ctx, name + '.N2', q(out), labels))) if 'G.' in name else q(torch.relu(out))
    out = q(Conv2D(ctx, out, name + '.Conv2', sn, fold='pool' if resample == 'down' else None))
    return q((upsample(short) if resample == 'up' else short) + out)


def Generator(ctx, noise, labels, dim=128):
    """gan_resnet.py:356-371.  Returns NHWC [n,32,32,3]."""
    n = 'Generator/'
    q = O.q
    out = q(Linear(ctx, q(noise), n + 'G.Input', False, tc=True)).reshape(-1, 4, 4, dim * 8)
    out = ResidualBlock(ctx, out, n + 'G.Block.1', False, 'up', labels, dim * 8, dim * 2)
    out = ResidualBlock(ctx, out, n + 'G.Block.2', False, 'up', labels, dim * 2, dim * 2)
    out = ResidualBlock(ctx, out, n + 'G.Block.3', False, 'up', labels, dim * 2, dim * 2)
    out = q(torch.relu(Normalize(ctx, n + 'G.OutputNorm', out, labels)))
    return O.stored_act(Conv2D(ctx, out, n + 'G.Output', False), 'tanh')


def Discriminator(ctx, x, dim=128):
    """gan_resnet.py:331-353, 374-412.  x NHWC [N,32,32,3] -> (output [N,dim], output_wgan [N])."""
    n = 'Discriminator/'
    q = O.q
    if O.emulating():
        # the product's dataflow / storage points (see _ResidualBlock_bf16): image cast to bf16, pooled shortcut input stored, Conv2 +
        # mean-pool as one folded stride-2 conv with the shortcut added in its epilogue; the [N, dim] head runs in fp32
        x = q(x)
        short = q(Conv2D(ctx, q(mean_pool(x)), n + 'D.Block.1.Shortcut', True))
        out = q(torch.relu(Conv2D(ctx, x, n + 'D.Block.1.Conv1', True)))
        out = q(short + q(Conv2D(ctx, out, n + 'D.Block.1.Conv2', True, fold='pool')))
    else:
        short = Conv2D(ctx, mean_pool(x), n + 'D.Block.1.Shortcut', True)
        out = Conv2D(ctx, x, n + 'D.Block.1.Conv1', True)
        out = mean_pool(Conv2D(ctx, torch.relu(out), n + 'D.Block.1.Conv2', True))
        out = short + out
    out = ResidualBlock(ctx, out, n + 'D.Block.2', True, 'down', None, dim, dim)
    for i in range(3, 7):
        out = ResidualBlock(ctx, out, n + 'D.Block.%d' % i, True, None, None, dim, dim)
    out = q(torch.relu(out).mean(dim=(1, 2)))
    return out, Linear(ctx, out, n + 'D.Output', True).reshape(-1)


def Discriminator_projection(ctx, labels):
    """gan_resnet.py:414-421: embed_y -> SN-Linear(300 -> dim)."""
    n = 'Discriminator/'
    return Linear(ctx, ctx.P[n + 'Embedding.Label/embedding_map'][labels], n + 'D.Embedding_y', True)


def perm_classifier(ctx, x):
    """gan_resnet.py:456-483: perm_type 'linear' (one SN-Linear) or '2layer' (two, no nonlinearity between them) -- told apart
    by the presence of the second layer's weights."""
    h = Linear(ctx, x.reshape(x.shape[0], -1), 'Discriminator/D.d_perm_classifier_h1', True)
    if 'Discriminator/D.d_perm_classifier_h2/W' in ctx.P:
        h = Linear(ctx, h, 'Discriminator/D.d_perm_classifier_h2', True)
    return h


def confusion_matrix(P, cfg, dtype):
    if cfg.algorithm == 'rcgan-u':
        return torch.softmax(P['confusion_logits'], -1)
    a = cfg.alpha
    return torch.as_tensor(((1 - a) / 9.0) * np.ones((10, 10)) + (a - (1 - a) / 9.0) * np.eye(10), dtype=dtype)


def disc_cost(P, batch, cfg):
    """One tower of gan_resnet.py:557-695.  batch: real [n,32,32,3] NHWC float, labels (noisy real), labels_random,
    labels_biased, inv_weights [n,10], noise [n,128]."""
    dt = batch['real'].dtype
    ctx = Ctx(P, True)
    nctx = Ctx(P, False)
    n = batch['real'].shape[0]
    fake = Generator(nctx, batch['noise'], batch['labels_random'], cfg.dim)
    eye = torch.eye(VOCAB, dtype=dt)
    V = Discriminator_projection(ctx, torch.arange(VOCAB))
    if cfg.algorithm == 'rcgan-u':
        h_r, psi_r = Discriminator(ctx, batch['real'], cfg.dim)
        disc_real = psi_r + (h_r * V[batch['labels']]).sum(1)
        h_f, psi_f = Discriminator(ctx, fake, cfg.dim)
        disc_fake = psi_f[:, None] + h_f @ V.t()
        w = eye[batch['labels_random']] @ confusion_matrix(P, cfg, dt)
        cost = (torch.relu(1. + disc_fake) * w).sum(1).mean() + torch.relu(1. - disc_real).mean()
    else:
        h, psi = Discriminator(ctx, torch.cat([batch['real'], fake], 0), cfg.dim)
        if cfg.algorithm == 'unbiased':
            all_r = torch.relu(1. - (psi[:n, None] + h[:n] @ V.t()))
            disc_fake = psi[n:] + (h[n:] * V[batch['labels_random']]).sum(1)
            cost = (all_r * batch['inv_weights']).sum(1).mean() + torch.relu(1. + disc_fake).mean()
        else:
            lab_f = batch['labels_random'] if cfg.algorithm == 'biased' else batch['labels_biased']
            lab = torch.cat([batch['labels'], lab_f], 0)
            disc_all = psi + (h * V[lab]).sum(1)
            cost = torch.relu(1. - disc_all[:n]).mean() + torch.relu(1. + disc_all[n:]).mean()
    out = {'disc_wgan': cost}
    if cfg.perm_classifier:
        pl = O.sigmoid_ce(perm_classifier(ctx, batch['real']), eye[batch['labels']]).mean()
        out['perm_real'] = pl
        cost = cost + 1. * pl
    out['disc_cost'] = cost
    return out, ctx.u_new


def gen_cost(P, batch, cfg):
    """One tower of gan_resnet.py:715-786.  batch: noise_G [2n,128], labels_random_G, labels_biased_G [2n]."""
    dt = batch['noise_G'].dtype
    dctx = Ctx(P, False)          # Discriminator(..., update_collection="NO_OPS")
    pctx = Ctx(P, True)           # Discriminator_projection(..., update_collection=None)
    fake = Generator(dctx, batch['noise_G'], batch['labels_random_G'], cfg.dim)
    h, psi = Discriminator(dctx, fake, cfg.dim)
    V = Discriminator_projection(pctx, torch.arange(VOCAB))
    eye = torch.eye(VOCAB, dtype=dt)
    if cfg.algorithm == 'rcgan-u':
        disc_fake = psi[:, None] + h @ V.t()
        w = eye[batch['labels_random_G']] @ confusion_matrix(P, cfg, dt)
        cost = ((-disc_fake) * w).sum(1).mean()
    else:
        lab = batch['labels_biased_G'] if cfg.algorithm == 'rcgan' else batch['labels_random_G']
        cost = -(psi + (h * V[lab]).sum(1)).mean()
    out = {'gen_wgan': cost}
    if cfg.perm_classifier:
        # perm_classifier never passes update_collection (gan_resnet.py:458-466) -> default None -> u IS updated here
        pl = O.sigmoid_ce(perm_classifier(pctx, O.qg(fake)), eye[batch['labels_random_G']]).mean()     # one rounding per gradient writer
        out['perm_fake'] = pl
        cost = cost + cfg.perm_multiplier * pl
    out['gen_cost'] = cost
    return out, pctx.u_new


def lr_decay(it, niters=50000):
    """gan_resnet.py:700-703."""
    return max(0., 1. - it / 100000.) if it < 50000 else 0.5


def d_names(P):
    return [n for n in P if 'Discriminator' in n and not n.endswith('/u')]


def g_names(P):
    return [n for n in P if 'Generator' in n]


class Trainer:
    """gan_resnet.py:802-817 + :919-947: Adam(lr*decay, 0, .9) for D and G, Adam(lr*confuse_multiplier[*decay]) for C;
    an iteration = [G(+C) step if it > 0] + n_critic D steps."""

    def __init__(self, P, cfg):
        self.P, self.cfg = P, cfg
        self.dn, self.gn = d_names(P), g_names(P)
        self.d_opt = O.TFAdam(self.dn, cfg.lr, 0.0, 0.9)
        self.g_opt = O.TFAdam(self.gn, cfg.lr, 0.0, 0.9)
        self.c_opt = O.TFAdam(['confusion_logits'], cfg.lr * cfg.confuse_multiplier, 0.0, 0.9) if cfg.algorithm == 'rcgan-u' else None
        self.last = {}

    def _req(self, names):
        for n in self.P:
            self.P[n] = self.P[n].detach()
        for n in names:
            self.P[n] = self.P[n].requires_grad_(True)

    def d_step(self, batch, it=0):
        self._req(self.dn)
        out, u_new = disc_cost(self.P, batch, self.cfg)
        gs = torch.autograd.grad(out['disc_cost'], [self.P[n] for n in self.dn], allow_unused=True)
        grads = {n: (g if g is not None else torch.zeros_like(self.P[n])) for n, g in zip(self.dn, gs)}
        self.d_opt.step(self.P, grads, lr=self.cfg.lr * lr_decay(it))
        for n, v in u_new.items():
            self.P[n] = v
        self.last['d'] = {k: v.detach() for k, v in out.items()}
        self.last['d_grads'] = grads

    def g_step(self, batch, it=0):
        names = self.gn + (['confusion_logits'] if self.c_opt else [])
        self._req(names)
        out, u_new = gen_cost(self.P, batch, self.cfg)
        gs = torch.autograd.grad(out['gen_cost'], [self.P[n] for n in names], allow_unused=True)
        grads = {n: (g if g is not None else torch.zeros_like(self.P[n])) for n, g in zip(names, gs)}
        self.g_opt.step(self.P, grads, lr=self.cfg.lr * lr_decay(it))
        if self.c_opt:
            clr = self.cfg.lr * self.cfg.confuse_multiplier * (lr_decay(it) if self.cfg.confuse_lr_decay else 1.0)
            self.c_opt.step(self.P, grads, lr=clr)
        for n, v in u_new.items():
            self.P[n] = v
        self.last['g'] = {k: v.detach() for k, v in out.items()}
        self.last['g_grads'] = grads


def synthetic_batch(n, seed=0, dtype=torch.float64, alpha=0.5):
    """Config-4 style inputs: integer image CHW -> preprocessed NHWC without dequantisation noise, labels via
    the seeded cifar sampler, N(0,1) noise."""
    from . import sampler as S
    rs = np.random.RandomState(seed)
    raw = rs.randint(0, 256, size=(n, 3072)).astype(np.int32)
    true = rs.randint(10, size=3 * n)
    C = S.one_coin_confusion(alpha)
    lab, inv_w, rnd, biased = S.cifar_labels_numpy(true, C, 547)
    real = (2 * (raw.astype(np.float64) / 256. - .5)).reshape(n, 3, 32, 32).transpose(0, 2, 3, 1)
    t = lambda a, d=dtype: torch.as_tensor(np.ascontiguousarray(a), dtype=d)
    L = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.long)
    return dict(raw=torch.as_tensor(raw), real=t(real), labels=L(lab[:n]), labels_random=L(rnd[:n]), labels_biased=L(biased[:n]),
                inv_weights=t(inv_w[:n]), noise=t(rs.randn(n, 128)), noise_G=t(rs.randn(2 * n, 128)),
                labels_random_G=L(rnd[n:3 * n]), labels_biased_G=L(biased[n:3 * n]))
