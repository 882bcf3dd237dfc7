"""Drop-in for mnist/ops.py: same names, arguments and defaults; each call records ops of the
B200 static program (graph.py) instead of TensorFlow graph nodes."""
import torch

from . import scope as S
from . import sn
from .graph import cur
from .nnops import ActOp, BatchNormOp, ConcatLabelOp, ConvOp, DeconvOp


def _normal(std):
    return lambda shape: torch.randn(tuple(shape), generator=S.init_generator(), dtype=torch.float32) * std


def _trunc_normal(std):
    def f(shape):
        t = torch.empty(tuple(shape), dtype=torch.float32)
        torch.nn.init.trunc_normal_(t, 0.0, std, -2 * std, 2 * std, generator=S.init_generator())
        return t
    return f


def _const(v):
    return lambda shape: torch.full(tuple(shape), float(v), dtype=torch.float32)


class batch_norm(object):
    """mnist/ops.py:30-44 (tf.contrib.layers.batch_norm, decay=momentum, scale=True,
    updates_collections=None).  `fuse_act` lets the caller fold the relu / lrelu that always follows
    in the reference models into the same kernel; `track_moving=False` skips the moving averages of
    the discriminator's norms, which the reference updates but never reads."""

    def __init__(self, epsilon=1e-5, momentum=0.9, name="batch_norm"):
        self.epsilon = epsilon
        self.momentum = momentum
        self.name = name

    def __call__(self, x, train=True, fuse_act=None, track_moving=True, groups=1, concat_y=None):
        c = x.shape[-1]
        with S.variable_scope(self.name):
            beta = S.get_variable('beta', [c], _const(0.0))
            gamma = S.get_variable('gamma', [c], _const(1.0))
            mm = S.get_variable('moving_mean', [c], _const(0.0), trainable=False)
            mv = S.get_variable('moving_variance', [c], _const(1.0), trainable=False)
        moving = (mm, mv) if (track_moving or not train) else None
        # groups: equal sample ranges normalised with their own batch statistics (one pass over [real; fake])
        # concat_y: the conv_cond_concat / concat([h, y]) that follows the norm in the generator, written in the same pass
        return BatchNormOp(x, gamma, beta, None, moving, train, self.epsilon, self.momentum, fuse_act, groups=groups,
                           concat_y=concat_y).y


def conv_cond_concat(x, y):
    """mnist/ops.py:46-51.  y: fp32 [B, y_dim] (the reference reshapes it to [B,1,1,y_dim] first)."""
    return ConcatLabelOp(x, y).out


def concat_label(x, y):
    """concat([x, y], 1) for 2-D x (mnist/model.py:714, 718)."""
    return ConcatLabelOp(x, y).out


def conv2d(input_, output_dim, k_h=5, k_w=5, d_h=2, d_w=2, stddev=0.02, spectral_norm=False, name="conv2d",
           fuse_act=None, pre_norm=False):
    """mnist/ops.py:53-67."""
    assert d_h == d_w
    with S.variable_scope(name):
        w = S.get_variable('w', [k_h, k_w, input_.shape[-1], output_dim], _trunc_normal(stddev))
        if spectral_norm:
            w = sn.spectral_normed_weight(w, update_collection=None)
        biases = S.get_variable('biases', [output_dim], _const(0.0))
    return ConvOp(input_, w, biases, d_h, fuse_act, pre_norm=pre_norm).y


def deconv2d(input_, output_shape, k_h=5, k_w=5, d_h=2, d_w=2, stddev=0.02, name="deconv2d", with_w=False, fuse_act=None,
             pre_norm=False):
    """mnist/ops.py:69-92.  output_shape = [batch, h, w, channels]."""
    assert d_h == d_w
    with S.variable_scope(name):
        w = S.get_variable('w', [k_h, k_w, output_shape[-1], input_.shape[-1]], _normal(stddev))
        biases = S.get_variable('biases', [output_shape[-1]], _const(0.0))
    out = DeconvOp(input_, w, biases, (output_shape[1], output_shape[2]), d_h, fuse_act, pre_norm=pre_norm).y
    if with_w:
        return out, w, biases
    return out


def lrelu(x, leak=0.2, name="lrelu"):
    """mnist/ops.py:94-95."""
    return ActOp(x, 'lrelu', leak).y


def relu(x):
    return ActOp(x, 'relu').y


def sigmoid(x):
    return ActOp(x, 'sigmoid').y


def linear(input_, output_size, scope=None, stddev=0.02, bias_start=0.0, with_w=False, max_norm=False, pre_norm=False):
    """mnist/ops.py:97-116.  max_norm attaches the clip_by_value(-1,1) variable constraint, applied by
    the optimizer after each update to Matrix and bias."""
    with S.variable_scope(scope or "Linear"):
        matrix = S.get_variable("Matrix", [input_.shape[-1], output_size], _normal(stddev))
        bias = S.get_variable("bias", [output_size], _const(bias_start))
        if max_norm:
            matrix.clip = True
            bias.clip = True
    out = ConvOp(input_, matrix, bias, 1, None, pre_norm=pre_norm).y
    if with_w:
        return out, matrix, bias
    return out
