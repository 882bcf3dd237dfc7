"""Oracle primitives (test infrastructure only -- see oracle/__init__.py).

PyTorch-CPU restatement of the TensorFlow-1.5 ops the reference calls.  All
activations are NHWC, conv filters HWIO ([kh,kw,cin,cout]), deconv filters
[kh,kw,cout,cin] exactly as in the reference.  Works in any float dtype
(fp64 = truth, fp32 = what the reference itself computes in).
"""
import math

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- padding
def same_pad(size, k, stride):
    """TF 'SAME' padding (before, after) for one spatial dim; out = ceil(size/stride)."""
    out = -(-size // stride)
    total = max((out - 1) * stride + k - size, 0)
    return total // 2, total - total // 2


# --------------------------------------------------------------------------- conv
def conv2d(x, w, stride=1):
    """tf.nn.conv2d(x, w, strides=[1,s,s,1], padding='SAME')  (mnist/ops.py:62,
    cifar10/common/ops/conv2d.py:181-187).  x NHWC, w HWIO."""
    kh, kw = w.shape[0], w.shape[1]
    pt, pb = same_pad(x.shape[1], kh, stride)
    pl, pr = same_pad(x.shape[2], kw, stride)
    xn = F.pad(x.permute(0, 3, 1, 2), (pl, pr, pt, pb))
    y = F.conv2d(xn, w.permute(3, 2, 0, 1), stride=stride)
    return y.permute(0, 2, 3, 1)


def conv2d_transpose(x, w, out_hw, stride=2):
    """tf.nn.conv2d_transpose(x, w, output_shape, strides) with the default SAME
    padding (mnist/ops.py:78).  w is [kh,kw,cout,cin].  Equals the gradient of
    conv2d(SAME, stride) w.r.t. its input: full transposed conv, cropped at the
    forward conv's pad_before."""
    kh, kw = w.shape[0], w.shape[1]
    oh, ow = out_hw
    pt, _ = same_pad(oh, kh, stride)
    pl, _ = same_pad(ow, kw, stride)
    full = F.conv_transpose2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), stride=stride)
    # the full result covers rows [-pt, ...); rows past its end are zero
    need_h, need_w = pt + oh, pl + ow
    if full.shape[2] < need_h or full.shape[3] < need_w:
        full = F.pad(full, (0, max(0, need_w - full.shape[3]), 0, max(0, need_h - full.shape[2])))
    return full[:, :, pt:pt + oh, pl:pl + ow].permute(0, 2, 3, 1)


def conv_cond_concat(x, y):
    """mnist/ops.py:46-51: broadcast y[B,1,1,C] over H,W and concat on channels."""
    B, H, W, _ = x.shape
    return torch.cat([x, y.reshape(B, 1, 1, -1).expand(B, H, W, y.shape[-1])], dim=3)


def lrelu(x, leak=0.2):
    """mnist/ops.py:94-95: tf.maximum(x, leak*x)."""
    return torch.maximum(x, leak * x)


# --------------------------------------------------------------------------- batch norm
def batch_norm_train(x, gamma, beta, eps=1e-5):
    """tf.contrib.layers.batch_norm(is_training=True, scale=True, center=True)
    (mnist/ops.py:30-44): normalise over all axes but the last with the BIASED
    batch variance.  Returns (y, batch_mean, biased_var)."""
    C = x.shape[-1]
    xf = x.reshape(-1, C)
    mean = xf.mean(0)
    var = ((xf - mean) ** 2).mean(0)
    y = (x - mean) * torch.rsqrt(var + eps) * gamma + beta
    return y, mean, var


def batch_norm_moving_update(mm, mv, mean, var, count, decay=0.9):
    """TF fused BN moving-average update (updates_collections=None, decay=.9,
    zero_debias off): moving_variance receives the Bessel-corrected variance."""
    unbiased = var * (count / max(count - 1, 1))
    return decay * mm + (1 - decay) * mean, decay * mv + (1 - decay) * unbiased


def batch_norm_infer(x, gamma, beta, mm, mv, eps=1e-5):
    """is_training=False branch (gen_sampler, mnist/model.py:733-757)."""
    return (x - mm) * torch.rsqrt(mv + eps) * gamma + beta


def cond_batchnorm(x, labels, offset, scale, eps=1e-5):
    """cifar10/common/ops/normalization.py:27-59: moments over [0,1,2] (biased var),
    per-sample offset/scale gathered from [n_labels,C] tables; no moving stats."""
    C = x.shape[-1]
    xf = x.reshape(-1, C)
    mean = xf.mean(0)
    var = ((xf - mean) ** 2).mean(0)
    o = offset[labels][:, None, None, :]
    s = scale[labels][:, None, None, :]
    return (x - mean) * torch.rsqrt(var + eps) * s + o


# --------------------------------------------------------------------------- spectral norm
def l2normalize(v, eps=1e-12):
    """mnist/sn.py:13-14 -- eps is added to the NORM, not inside the sqrt."""
    return v / ((v ** 2).sum() ** 0.5 + eps)


def spectral_normed_weight(W, u, num_iters=1):
    """mnist/sn.py:17-75 (== cifar10/common/ops/sn.py).  Returns (W_bar, u_new, sigma).
    Autograd flows THROUGH the power iteration (the reference has no stop_gradient
    and tf.while_loop back-props by default); u (the stored vector) is a constant."""
    Wr = W.reshape(-1, W.shape[-1])
    u_i = u.detach()
    v_i = None
    for _ in range(num_iters):
        v_i = l2normalize(u_i @ Wr.t())
        u_i = l2normalize(v_i @ Wr)
    sigma = (v_i @ Wr @ u_i.t())[0, 0]
    return (Wr / sigma).reshape(W.shape), u_i.detach(), sigma


def sn_grad_closed_form(W, u0, G):
    """Closed-form dL/dW given G = dL/dW_bar for ONE power iteration.
    Used by tests to cross-check autograd and as the spec of the CUDA backward."""
    eps = 1e-12
    Wr = W.reshape(-1, W.shape[-1])
    Gr = G.reshape(Wr.shape)
    a = u0 @ Wr.t()                        # [1,m]
    na = (a ** 2).sum() ** 0.5
    v = a / (na + eps)
    b = v @ Wr                             # [1,c]
    n = (b ** 2).sum() ** 0.5
    sigma = n * n / (n + eps)
    g_sigma = -(Gr * Wr).sum() / sigma ** 2
    g_b = g_sigma * b * (n + 2 * eps) / (n + eps) ** 2
    g_v = g_b @ Wr.t()
    g_a = g_v / (na + eps) - (g_v * a).sum() * a / (na * (na + eps) ** 2)
    dW = Gr / sigma + v.t() @ g_b + g_a.t() @ u0
    return dW.reshape(W.shape)


# --------------------------------------------------------------------------- losses
def sigmoid_ce(logits, targets):
    """tf.nn.sigmoid_cross_entropy_with_logits: max(x,0) - x*z + log(1+exp(-|x|))."""
    return torch.clamp(logits, min=0) - logits * targets + torch.log1p(torch.exp(-logits.abs()))


def gan_loss_fns(loss_fn):
    """mnist/model.py:135-147 (hinge | ce); cifar10/gan_resnet.py:604-606 (hinge)."""
    if loss_fn == 'hinge':
        return (lambda x: F.relu(1 - x)), (lambda x: F.relu(1 + x)), (lambda x: -x)
    if loss_fn == 'ce':
        return (lambda x: sigmoid_ce(x, torch.ones_like(x)),
                lambda x: sigmoid_ce(x, torch.zeros_like(x)),
                lambda x: sigmoid_ce(x, torch.ones_like(x)))
    raise ValueError('Unknown loss_fn: {}!'.format(loss_fn))


# --------------------------------------------------------------------------- optimiser
class TFAdam:
    """tf.train.AdamOptimizer: theta -= lr*sqrt(1-b2^t)/(1-b1^t) * m/(sqrt(v)+eps)
    (eps OUTSIDE the bias correction -- differs from torch.optim.Adam).
    `clip` lists parameter names that carry the max-norm variable constraint
    (mnist/ops.py:101-111): clip_by_value(-1,1) applied after the update."""

    def __init__(self, names, lr, beta1, beta2=0.999, eps=1e-8, clip=()):
        self.names, self.lr, self.b1, self.b2, self.eps = list(names), lr, beta1, beta2, eps
        self.clip = set(clip)
        self.t = 0
        self.m, self.v = {}, {}

    def step(self, P, grads, lr=None):
        lr = self.lr if lr is None else lr
        self.t += 1
        lr_t = lr * math.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        for n in self.names:
            g = grads[n]
            if g is None:
                continue
            if n not in self.m:
                self.m[n] = torch.zeros_like(P[n])
                self.v[n] = torch.zeros_like(P[n])
            self.m[n] = self.b1 * self.m[n] + (1 - self.b1) * g
            self.v[n] = self.b2 * self.v[n] + (1 - self.b2) * g * g
            new = P[n].detach() - lr_t * self.m[n] / (self.v[n].sqrt() + self.eps)
            if n in self.clip:
                new = new.clamp(-1.0, 1.0)
            P[n] = new


# --------------------------------------------------------------------------- 2x resampling folded into the filter
_FOLD_SETS = {0: [[0], [0, 1], [1, 2], [2]],          # ConvMeanPool: taps k with a - k in {0, 1}
              1: [[2], [1, 2], [0, 1], [0]]}          # UpsampleConv, conv2d_transpose tap order


def fold4(w, mode):
    """Checker for rcgan_wfold4.  w [3,3,cin,cout] ->
    mode 0: w4 [4,4,cin,cout] with meanpool2(conv2d(x, w)) == conv2d(x, w4, stride 2)   (cifar10/gan_resnet.py:231-241)
    mode 1: w4 [4,4,cout,cin] with conv2d(upsample2(x), w) == conv2d_transpose(x, w4, 2h x 2w, stride 2)   (:259-272)."""
    S = _FOLD_SETS[mode]
    cin, cout = w.shape[2], w.shape[3]
    w4 = w.new_zeros((4, 4, cin, cout) if mode == 0 else (4, 4, cout, cin))
    for a in range(4):
        for b in range(4):
            acc = sum(w[k, l] for k in S[a] for l in S[b])
            w4[a, b] = 0.25 * acc if mode == 0 else acc.t()
    return w4


# --------------------------------------------------------------------------- bf16 storage emulation
# The product's benchmark precision stores activations, activation gradients and the tensor-core weight packs in bf16 and
# accumulates / normalises / reduces in fp32.  With `bf16_storage()` active the oracle rounds to bf16 at exactly those storage
# points (and nowhere else: arithmetic stays fp64), so "CUDA bf16 step vs this oracle" isolates IMPLEMENTATION error from the
# quantisation gap that "CUDA bf16 step vs the plain fp64 oracle" also contains (a 10 % kernel bug in a bf16-only
# path must not hide behind the 25 % conditioning bound).
_EMUL = {'on': False}


class _RoundBoth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


class _RoundFwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundBwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


def qg(x):
    """a pre-activation that exists only inside a conv epilogue: its VALUE is never stored, but its gradient is (the fused
    activation's backward runs in place on the bf16 output gradient before dgrad / wgrad / the bias sum read it)"""
    return _RoundBwd.apply(x) if _EMUL['on'] else x


class _StoredAct(torch.autograd.Function):
    """sigmoid / tanh fused into a conv epilogue: only the bf16-rounded OUTPUT is kept, and the backward derives the slope from it
    (y(1-y), 1-y^2) -- near saturation that differs visibly from the slope at the unrounded value -- then rounds the gradient in
    place"""
    @staticmethod
    def forward(ctx, x, kind):
        y = (torch.sigmoid(x) if kind == 'sigmoid' else torch.tanh(x)).to(torch.bfloat16).to(x.dtype)
        ctx.save_for_backward(y)
        ctx.kind = kind
        return y

    @staticmethod
    def backward(ctx, g):
        y, = ctx.saved_tensors
        slope = y * (1 - y) if ctx.kind == 'sigmoid' else 1 - y * y
        return (g.to(torch.bfloat16).to(g.dtype) * slope).to(torch.bfloat16).to(g.dtype), None


def stored_act(x, kind):
    """sigmoid / tanh output of a conv epilogue (see _StoredAct); plain act(x) outside the emulation"""
    if _EMUL['on']:
        y = _StoredAct.apply(x, kind)
        if _EMUL.get('trace') is not None:
            _EMUL['trace'].append(y.detach())
        return y
    return torch.sigmoid(x) if kind == 'sigmoid' else torch.tanh(x)


def emulating():
    return _EMUL['on']


def q(x):
    """an activation tensor the product keeps in HBM as bf16: the value AND its gradient are rounded"""
    if not _EMUL['on']:
        return x
    y = _RoundBoth.apply(x)
    if _EMUL.get('trace') is not None:
        _EMUL['trace'].append(y.detach())          # diagnostic: the sequence of stored tensors (tools/bf16_gap.py)
    return y


def qw(w):
    """a weight as the tensor cores see it (bf16 pack of the fp32 parameter); its gradient (wgrad output) stays fp32"""
    return _RoundFwd.apply(w) if _EMUL['on'] else w


class bf16_storage:
    """with bf16_storage(): the oracle graphs round at the product's bf16 storage points"""

    def __enter__(self):
        self.prev, _EMUL['on'] = _EMUL['on'], True
        return self

    def __exit__(self, *a):
        _EMUL['on'] = self.prev
