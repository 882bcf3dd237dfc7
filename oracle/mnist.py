"""Oracle restatement of the MNIST DCGAN RCGAN graph and training step
(test infrastructure only -- see oracle/__init__.py; parity unpinned for floats).

Follows mnist/model.py:96-262 (build_model + optimizers), :335-372 (one iteration),
:644-768 (discriminator / generator / gen_sampler / classifier) and mnist/ops.py.
Parameters live in a flat dict keyed by the reference's TF variable names.
"""
from types import SimpleNamespace

import numpy as np
import torch

from . import nn as O


def default_config(**kw):
    """Flag defaults of mnist/main.py:13-66 (only what the train step reads)."""
    cfg = SimpleNamespace(
        batch_size=100, learning_rate=2e-4, beta1=0.5, z_dim=100, y_dim=10,
        gf_dim=64, df_dim=64, gfc_dim=1024, dfc_dim=1024, c_dim=1,
        output_height=28, output_width=28,
        algorithm='biased', estimate_confuse=True, confuse_multiplier=10.0,
        perm_regularizer=True, perm_multiplier=10.0, alpha=1.0,
        disc_type='vanilla', loss_fn='hinge', real_match=False,
        concat_y=False, concat_y_layers=(1,), spectral_norm=True, max_norm=True)
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def trunc_normal(shape, std, gen, dtype):
    t = torch.empty(shape, dtype=torch.float64)
    torch.nn.init.trunc_normal_(t, 0.0, std, -2 * std, 2 * std, generator=gen)
    return t.to(dtype)


def init_params(cfg, seed=0, dtype=torch.float64, confusion_actual=None):
    """Random-init parameters with the reference's shapes/names/initialisers
    (SURVEY appendix C).  Also creates SN `u` vectors and BN moving stats."""
    g = torch.Generator().manual_seed(seed)
    P = {}
    normal = lambda shape: (torch.randn(shape, generator=g, dtype=torch.float64) * 0.02).to(dtype)
    zeros = lambda *s: torch.zeros(*s, dtype=dtype)
    ones = lambda *s: torch.ones(*s, dtype=dtype)

    def bn(name, c):
        P[name + '/beta'] = zeros(c); P[name + '/gamma'] = ones(c)
        P[name + '/moving_mean'] = zeros(c); P[name + '/moving_variance'] = ones(c)

    def conv(name, cin, cout, sn):
        P[name + '/w'] = trunc_normal((5, 5, cin, cout), 0.02, g, dtype)
        P[name + '/biases'] = zeros(cout)
        if sn:
            P[name + '/spectral_norm/u'] = trunc_normal((1, cout), 1.0, g, dtype)

    def lin(name, cin, cout):
        P[name + '/Matrix'] = normal((cin, cout)); P[name + '/bias'] = zeros(cout)

    y, df, gf = cfg.y_dim, cfg.df_dim, cfg.gf_dim
    s4 = cfg.output_height // 4
    # generator (mnist/model.py:705-731)
    lin('generator/g_h0_lin', cfg.z_dim + y, cfg.gfc_dim); bn('generator/g_bn0', cfg.gfc_dim)
    lin('generator/g_h1_lin', cfg.gfc_dim + y, gf * 2 * s4 * s4); bn('generator/g_bn1', gf * 2 * s4 * s4)
    P['generator/g_h2/w'] = normal((5, 5, gf * 2, gf * 2 + y)); P['generator/g_h2/biases'] = zeros(gf * 2)
    bn('generator/g_bn2', gf * 2)
    P['generator/g_h3/w'] = normal((5, 5, cfg.c_dim, gf * 2 + y)); P['generator/g_h3/biases'] = zeros(cfg.c_dim)
    # discriminator (mnist/model.py:644-703)
    if cfg.disc_type == 'projection':
        cy = lambda l: y if (cfg.concat_y and l in cfg.concat_y_layers) else 0
        conv('discriminator/d_h0_conv', cfg.c_dim + cy(1), df, cfg.spectral_norm)
        conv('discriminator/d_h1_conv', df + cy(2), df, cfg.spectral_norm); bn('discriminator/d_bn1', df)
        conv('discriminator/d_h2_conv', df + cy(3), df, cfg.spectral_norm); bn('discriminator/d_bn2', df)
        conv('discriminator/d_h3_conv', df + cy(4), df, cfg.spectral_norm); bn('discriminator/d_bn3', df)
        lin('discriminator/d_h4_lin', df, 1)
        lin('discriminator/d_h5_y_lin', y, df)
    else:
        conv('discriminator/d_h0_conv', cfg.c_dim + y, cfg.c_dim + y, False)
        conv('discriminator/d_h1_conv', cfg.c_dim + 2 * y, df + y, False); bn('discriminator/d_bn1', df + y)
        s4v = -(-(-(-cfg.output_height // 2)) // 2)
        lin('discriminator/d_h3_lin', s4v * s4v * (df + y) + y, cfg.dfc_dim); bn('discriminator/d_bn2', cfg.dfc_dim)
        lin('discriminator/d_h4_lin', cfg.dfc_dim + y, 1)
    if cfg.perm_regularizer:
        lin('classifier/d_classifier_h1', cfg.output_height * cfg.output_width * cfg.c_dim, y)
    if cfg.estimate_confuse:
        lim = np.sqrt(6.0 / (2 * y))   # TF default glorot_uniform on [y,y]
        P['confusion_logits'] = ((torch.rand((y, y), generator=g, dtype=torch.float64) * 2 - 1) * lim).to(dtype)
    return P


def trainable(name):
    return not (name.endswith('/u') or 'moving_' in name)


def d_var_names(P):
    """mnist/model.py:244: substring 'd_' (this catches classifier/d_classifier_h1)."""
    return [n for n in P if trainable(n) and 'd_' in n]


def g_var_names(P):
    """mnist/model.py:245."""
    return [n for n in P if trainable(n) and 'g_' in n]


def max_norm_names(P, cfg):
    """ops.linear(max_norm=True) call sites: d_h4_lin, d_h5_y_lin (model.py:680-683)."""
    if not (cfg.max_norm and cfg.disc_type == 'projection'):
        return []
    return [n for n in P if n.startswith('discriminator/d_h4_lin/') or n.startswith('discriminator/d_h5_y_lin/')]


class Updates:
    """Side effects of one graph evaluation: new SN u's and BN moving stats."""

    def __init__(self):
        self.vals = {}

    def set(self, name, val):
        self.vals[name] = val.detach()

    def apply(self, P):
        for n, v in self.vals.items():
            P[n] = v


def _bn(P, name, x, train, upd, track=True):
    g, b = P[name + '/gamma'], P[name + '/beta']
    if not train:
        return O.batch_norm_infer(x, g, b, P[name + '/moving_mean'], P[name + '/moving_variance'])
    y, mean, var = O.batch_norm_train(x, g, b)
    if upd is not None and track:
        cnt = x.numel() // x.shape[-1]
        mm, mv = O.batch_norm_moving_update(P[name + '/moving_mean'], P[name + '/moving_variance'],
                                            mean.detach(), var.detach(), cnt)
        upd.set(name + '/moving_mean', mm); upd.set(name + '/moving_variance', mv)
    return y


def _conv(P, name, x, cfg_sn, upd):
    w = P[name + '/w']
    if cfg_sn:
        w, u_new, _ = O.spectral_normed_weight(w, P[name + '/spectral_norm/u'])
        if upd is not None:
            upd.set(name + '/spectral_norm/u', u_new)
    return O.conv2d(x, O.qw(w), 2) + P[name + '/biases']       # (O.qw: identity unless O.bf16_storage() is active)


def _lin(P, name, x, tc=False):
    """tc: a linear the product runs on the tensor cores in bf16 mode (bf16 weight pack); the [B, 64] head and the classifier
    run in fp32 there"""
    return x @ (O.qw(P[name + '/Matrix']) if tc else P[name + '/Matrix']) + P[name + '/bias']


def generator(P, z, y, cfg, train=True, upd=None):
    """mnist/model.py:705-731 (train=True) and gen_sampler :733-757 (train=False)."""
    B = z.shape[0]
    s_h, s_w = cfg.output_height, cfg.output_width
    s_h2, s_h4, s_w2, s_w4 = s_h // 2, s_h // 4, s_w // 2, s_w // 4
    n = 'generator/'
    q, qw = O.q, O.qw              # bf16 storage points of the product (identity unless O.bf16_storage() is active)
    zc = torch.cat([q(z), y], 1)
    h0 = q(torch.relu(_bn(P, n + 'g_bn0', q(_lin(P, n + 'g_h0_lin', zc, tc=True)), train, upd)))
    h0 = torch.cat([h0, y], 1)
    h1 = q(torch.relu(_bn(P, n + 'g_bn1', q(_lin(P, n + 'g_h1_lin', h0, tc=True)), train, upd)))
    h1 = h1.reshape(B, s_h4, s_w4, cfg.gf_dim * 2)
    h1 = O.conv_cond_concat(h1, y)
    h2 = q(O.conv2d_transpose(h1, qw(P[n + 'g_h2/w']), (s_h2, s_w2)) + P[n + 'g_h2/biases'])
    h2 = q(torch.relu(_bn(P, n + 'g_bn2', h2, train, upd)))
    h2 = O.conv_cond_concat(h2, y)
    h3 = O.conv2d_transpose(h2, qw(P[n + 'g_h3/w']), (s_h, s_w)) + P[n + 'g_h3/biases']
    return O.stored_act(h3, 'sigmoid')


def discriminator(P, image, y, cfg, upd=None):
    """mnist/model.py:644-703.  Returns logits [B,1].  D's BN always runs in training
    mode; its moving stats are never read, so they are not tracked here."""
    B = image.shape[0]
    n = 'discriminator/'
    if cfg.disc_type == 'projection':
        cc = lambda l, t: O.conv_cond_concat(t, y) if (cfg.concat_y and l in cfg.concat_y_layers) else t
        sn = cfg.spectral_norm
        q = O.q                    # bf16 storage points of the product's trunk (identity unless O.bf16_storage() is active)
        h0 = q(O.lrelu(O.qg(_conv(P, n + 'd_h0_conv', cc(1, q(image)), sn, upd))))
        h1 = q(O.lrelu(_bn(P, n + 'd_bn1', q(_conv(P, n + 'd_h1_conv', cc(2, h0), sn, upd)), True, upd, track=False)))
        h2 = q(O.lrelu(_bn(P, n + 'd_bn2', q(_conv(P, n + 'd_h2_conv', cc(3, h1), sn, upd)), True, upd, track=False)))
        h3 = q(O.lrelu(_bn(P, n + 'd_bn3', q(_conv(P, n + 'd_h3_conv', cc(4, h2), sn, upd)), True, upd, track=False)))
        h3 = q(h3.mean(dim=(1, 2)))
        h4 = _lin(P, n + 'd_h4_lin', h3.reshape(B, -1))
        h5 = _lin(P, n + 'd_h5_y_lin', y.reshape(B, 10))
        return h4 + (h3 * h5).sum(1, keepdim=True)
    x = O.conv_cond_concat(image, y)
    h0 = O.lrelu(_conv(P, n + 'd_h0_conv', x, False, upd))
    h0 = O.conv_cond_concat(h0, y)
    h1 = O.lrelu(_bn(P, n + 'd_bn1', _conv(P, n + 'd_h1_conv', h0, False, upd), True, upd, track=False))
    h1 = torch.cat([h1.reshape(B, -1), y], 1)
    h3 = O.lrelu(_bn(P, n + 'd_bn2', _lin(P, n + 'd_h3_lin', h1), True, upd, track=False))
    h3 = torch.cat([h3, y], 1)
    return _lin(P, n + 'd_h4_lin', h3)


def classifier(P, x):
    """mnist/model.py:759-768."""
    return _lin(P, 'classifier/d_classifier_h1', x.reshape(x.shape[0], -1))


def confusion_matrix(P, cfg, C_actual, dtype):
    """mnist/model.py:102-108."""
    if cfg.estimate_confuse:
        return torch.softmax(P['confusion_logits'], dim=-1)
    return torch.as_tensor(np.asarray(C_actual), dtype=dtype)


def losses(P, batch, cfg, C_actual=None, upd=None):
    """The loss section of build_model, mnist/model.py:126-233, evaluated literally
    (10 discriminator calls for the unbiased real branch and for learned C).
    batch: dict(x, z, y_real, y_gen, y_fake, y_real_weights) -- one-hot float labels."""
    x, z = batch['x'], batch['z']
    dtype = x.dtype
    B = x.shape[0]
    real_fn, fake_fn, g_fn = O.gan_loss_fns(cfg.loss_fn)
    out = {}
    G = generator(P, z, batch['y_gen'], cfg, True, upd)
    out['G'] = G
    eye = torch.eye(10, dtype=dtype)
    if cfg.algorithm in ('biased', 'rcgan', 'ambient'):
        D_logits = discriminator(P, x, batch['y_real'], cfg, upd)
        out['d_loss_real'] = real_fn(D_logits).mean()
    elif cfg.algorithm == 'unbiased':
        allr = torch.cat([real_fn(discriminator(P, x, eye[i].expand(B, 10), cfg, upd)) for i in range(10)], 1)
        out['d_loss_real'] = (allr * batch['y_real_weights']).sum(1).mean()
    else:
        raise ValueError(cfg.algorithm)
    d_loss_fake = g_loss = None
    if cfg.algorithm in ('rcgan', 'ambient'):
        if not cfg.estimate_confuse:
            D_logits_ = discriminator(P, G, batch['y_fake'], cfg, upd)
        else:
            la = torch.cat([discriminator(P, G, eye[i].expand(B, 10), cfg, upd) for i in range(10)], 1)
            C = confusion_matrix(P, cfg, C_actual, dtype)
            w = batch['y_gen'] @ C
            d_loss_fake = (fake_fn(la) * w).sum(1).mean()
            g_loss = (g_fn(la) * w).sum(1).mean()
            out['D_logits_all_'] = la
    else:
        D_logits_ = discriminator(P, G, batch['y_gen'], cfg, upd)
    if d_loss_fake is None:
        d_loss_fake = fake_fn(D_logits_).mean()
        g_loss = g_fn(D_logits_).mean()
    out['d_loss_fake'], out['g_loss'] = d_loss_fake, g_loss
    if cfg.perm_regularizer:
        out['class_loss_real'] = O.sigmoid_ce(classifier(P, x), batch['y_real']).mean()
        # (O.qg: in the product the classifier's gradient is WRITTEN to the image's bf16 gradient buffer before the discriminator's
        # is accumulated into it -- one rounding per writer; identity outside the bf16 storage emulation)
        out['class_loss_fake'] = O.sigmoid_ce(classifier(P, O.qg(G)), batch['y_gen']).mean()
    else:
        out['class_loss_real'] = torch.zeros((), dtype=dtype)
        out['class_loss_fake'] = torch.zeros((), dtype=dtype)
    out['d_loss'] = out['d_loss_real'] + out['d_loss_fake']
    return out


def _grads(loss, P, names, retain=False):
    leaves = [P[n] for n in names]
    gs = torch.autograd.grad(loss, leaves, allow_unused=True, retain_graph=retain)
    return {n: (g if g is not None else torch.zeros_like(P[n])) for n, g in zip(names, gs)}


class Trainer:
    """mnist/model.py:250-262 + :335-372: three TF-Adam optimizers; one iteration =
    1 D step, then 2 x (G step + C step) on the same z / labels."""

    def __init__(self, P, cfg, C_actual=None):
        self.P, self.cfg, self.C_actual = P, cfg, C_actual
        self.d_names, self.g_names = d_var_names(P), g_var_names(P)
        self.d_opt = O.TFAdam(self.d_names, cfg.learning_rate, cfg.beta1, clip=max_norm_names(P, cfg))
        self.g_opt = O.TFAdam(self.g_names, cfg.learning_rate, cfg.beta1)
        self.c_opt = O.TFAdam(['confusion_logits'], cfg.learning_rate * cfg.confuse_multiplier, cfg.beta1) \
            if cfg.estimate_confuse else None
        self.last = {}

    def _req(self, names):
        for n in names:
            self.P[n] = self.P[n].detach().requires_grad_(True)

    def d_step(self, batch):
        self._req(self.d_names)
        upd = Updates()
        L = losses(self.P, batch, self.cfg, self.C_actual, upd)
        total = L['d_loss'] + 1.0 * L['class_loss_real']
        grads = _grads(total, self.P, self.d_names)
        self.d_opt.step(self.P, grads)
        upd.apply(self.P)
        self.last['d'] = {k: v.detach() for k, v in L.items()}
        self.last['d_grads'] = grads
        return L

    def g_step(self, batch):
        names = self.g_names + (['confusion_logits'] if self.c_opt else [])
        self._req(names)
        upd = Updates()
        L = losses(self.P, batch, self.cfg, self.C_actual, upd)
        total = L['g_loss'] + self.cfg.perm_multiplier * L['class_loss_fake']
        grads = _grads(total, self.P, self.g_names, retain=bool(self.c_opt))
        if self.c_opt:
            cg = _grads(L['g_loss'], self.P, ['confusion_logits'])
        self.g_opt.step(self.P, grads)
        if self.c_opt:
            self.c_opt.step(self.P, cg)
            grads = dict(grads, **cg)
        upd.apply(self.P)
        self.last['g'] = {k: v.detach() for k, v in L.items()}
        self.last['g_grads'] = grads
        return L

    def iteration(self, batch):
        self.d_step(batch)
        self.g_step(batch)
        self.g_step(batch)
        for n in self.P:
            self.P[n] = self.P[n].detach()


def synthetic_batch(B, seed=0, dtype=torch.float64, C=None, cfg=None):
    """Config-1 style inputs: X~U[0,1), labels through the seeded numpy
    sampler, z from the same numpy stream."""
    from . import sampler as S
    rs = np.random.RandomState(seed)
    x = rs.uniform(0, 1, size=(B, 28, 28, 1))
    y_true = rs.randint(10, size=B)
    if C is None:
        C = S.one_coin_confusion(0.5)
    lab = S.mnist_labels_numpy(y_true, C, real_match=bool(cfg and cfg.real_match), seed=547)
    z = np.random.uniform(-1, 1, [B, 100]).astype(np.float32)
    t = lambda a: torch.as_tensor(np.asarray(a), dtype=dtype)
    return dict(x=t(x), z=t(z), y_real=t(lab['y_real']), y_gen=t(lab['y_gen']), y_fake=t(lab['y_fake']),
                y_real_weights=t(lab['y_real_weights']))


# ----------------------------------------------------------------------------- label recovery
def recover_loss(P, z_recover, y_logit_recover, sample_actual, cfg):
    """mnist/model.py:516-541.  z_recover [R*k, z_dim], y_logit_recover [R, k], sample_actual [R, H, W, c].
    gen_sampler (:733-757: batch norm with the MOVING statistics) on (z_recover, tile(eye(k), R)); returns
    (mse_loss, y_recover [R,k], sq_sum [R,k])."""
    R, k = y_logit_recover.shape
    y_recover = torch.softmax(1 * y_logit_recover, dim=-1)                      # bignum = 1 (:516,520)
    hard_y = torch.eye(k, dtype=z_recover.dtype).repeat(R, 1)                   # tf.tile(eye, [R, 1]) (:524-525)
    samples = generator(P, z_recover, hard_y, cfg, train=False)                 # [R*k, H, W, c]
    samples = samples.reshape((R, k) + tuple(sample_actual.shape[1:]))          # (:532-535)
    sq_sum = ((sample_actual.unsqueeze(1) - samples) ** 2).mean(-1).mean(-1).mean(-1)   # (:537-540)
    mse_loss = (sq_sum * y_recover).sum(-1).mean()                              # (:541)
    return mse_loss, y_recover, sq_sum


def recover_step(P, z_recover, y_logit_recover, sample_actual, cfg, lr=500.0):
    """One GradientDescentOptimizer(lr).minimize(mse_loss, var_list=[z_recover, y_logit_recover]) step (:612-617, 626-629).
    Returns (new z_recover, new y_logit_recover, mse_loss before the update, (dz, dlogit))."""
    z = z_recover.detach().clone().requires_grad_(True)
    yl = y_logit_recover.detach().clone().requires_grad_(True)
    loss, _, _ = recover_loss({n: v.detach() for n, v in P.items()}, z, yl, sample_actual, cfg)
    dz, dyl = torch.autograd.grad(loss, [z, yl])
    return (z - lr * dz).detach(), (yl - lr * dyl).detach(), loss.detach(), (dz, dyl)


def zero_one_loss(y_actual, y_recover):
    """tf.losses.cosine_distance(y_actual, one_hot(argmax y_recover), dim=-1) (:545-546) = mean(1 - <a, b>)."""
    onehot = torch.eye(y_recover.shape[1], dtype=y_recover.dtype)[y_recover.argmax(-1)]
    return (1 - (y_actual * onehot).sum(-1)).mean()
