"""Drop-ins for cifar10/common/ops/{conv2d,linear,normalization,embedding}.py (the branches gan_resnet.py reaches)."""
import numpy as np
import torch

from .. import scope as S
from .. import sn
from ..graph import cur
from ..nnops import AddOp, BatchNormOp, ConvOp, DeconvOp, FoldWeightOp, GatherRowsOp, Upsample2Op


def _uniform(stdev):
    def f(shape):
        lim = stdev * np.sqrt(3)
        return (torch.rand(tuple(shape), generator=S.init_generator(), dtype=torch.float32) * 2 - 1) * lim
    return f


def _const(v):
    return lambda shape: torch.full(tuple(shape), float(v), dtype=torch.float32)


def Conv2D(inputs, input_dim, output_dim, filter_size=3, stride=1, name=None, conv_type='conv2d', channel_multiplier=0,
           padding='SAME', spectral_normed=False, update_collection=None, inputs_norm=False, he_init=True, mask_type=None,
           weightnorm=None, biases=True, gain=1., fuse_act=None, pre_norm=False, residual=None, up_op=None, fold=None,
           residual_up=False):
    """cifar10/common/ops/conv2d.py:31-218, plain conv2d branch: uniform He/Glorot init (:83-127), spectral norm under
    scope `filters` (:169-171), stride-`stride` SAME conv (:181-187), + Biases (:212-216).
    fold='pool': the caller mean-pools this 3x3 conv's output (ConvMeanPool) -> ONE 4x4 stride-2 conv with the folded filter;
    fold='up': the caller feeds the 2x nearest-neighbour upsampling of `inputs` (UpsampleConv; pass the SMALL tensor) -> ONE
    4x4 stride-2 conv2d_transpose.  Variables, initialisers and the spectral norm are those of the 3x3 filter.
    residual (+ residual_up): `shortcut + output` of the calling ResidualBlock (gan_resnet.py:328) added in this conv's epilogue;
    residual_up: the shortcut is still at HALF resolution (1x1 UpsampleConv computed on the small grid) and is upsampled on the
    fly.  Where the conv has no fused-epilogue path the same result is produced with explicit upsample / add kernels."""
    if conv_type != 'conv2d' or channel_multiplier or mask_type is not None or weightnorm or inputs_norm or padding != 'SAME':
        raise NotImplementedError('only the plain conv2d branch is reachable from gan_resnet.py')
    assert inputs.shape[-1] == input_dim
    with S.variable_scope(name):
        fan_in = input_dim * filter_size ** 2
        fan_out = output_dim * filter_size ** 2 / (stride ** 2)
        stdev = np.sqrt((4. if he_init else 2.) / (fan_in + fan_out))
        filters = S.get_variable('Filters', [filter_size, filter_size, input_dim, output_dim], _uniform(stdev))
        w = filters
        if spectral_normed:
            with S.variable_scope('filters'):
                w = sn.spectral_normed_weight(filters, update_collection=update_collection)
        b = S.get_variable('Biases', [output_dim], _const(0.)) if biases else None
    if fold == 'up':
        assert filter_size == 3 and stride == 1 and residual is None and up_op is None
        n, h, wd = inputs.shape[0], inputs.shape[1], inputs.shape[2]
        return DeconvOp(inputs, FoldWeightOp(w, fold).w4, b, (2 * h, 2 * wd), 2, fuse_act, pre_norm=pre_norm).y
    if fold == 'pool':
        assert filter_size == 3 and stride == 1 and up_op is None
        op = ConvOp(inputs, FoldWeightOp(w, fold).w4, b, 2, fuse_act, pre_norm=pre_norm)
    else:
        op = ConvOp(inputs, w, b, stride, fuse_act, pre_norm=pre_norm, up_op=up_op)
    if residual is None:
        return op.y
    if op._plain_tc_fprop() or not residual_up:
        op.attach_residual(residual, residual_up)
        return op.y
    return AddOp(Upsample2Op(residual).y if residual_up else residual, op.y).y


def Linear(inputs, input_dim, output_dim, name=None, spectral_normed=False, update_collection=None, reuse=None,
           inputs_norm=False, biases=True, initialization=None, weightnorm=None, gain=1., pre_norm=False):
    """cifar10/common/ops/linear.py:38-182 (Glorot-uniform init :76-80, spectral norm :161-171, + b :176-180)."""
    if weightnorm or inputs_norm or initialization is not None:
        raise NotImplementedError('only the default initialisation / no weightnorm branch is reachable')
    assert inputs.shape[-1] == input_dim
    with S.variable_scope(name):
        W = S.get_variable('W', [input_dim, output_dim], _uniform(np.sqrt(2. / (input_dim + output_dim))))
        w = sn.spectral_normed_weight(W, update_collection=update_collection) if spectral_normed else W
        b = S.get_variable('b', [output_dim], _const(0.)) if biases else None
    return ConvOp(inputs, w, b, 1, None, pre_norm=pre_norm).y


def cond_batchnorm(name, axes, inputs, is_training=None, stats_iter=None, update_moving_stats=True, fused=True, labels=None,
                   n_labels=None, fuse_act=None, groups=1):
    """cifar10/common/ops/normalization.py:27-59: moments over [0,1,2], per-label offset/scale tables, eps 1e-5, no
    moving statistics.  labels: int32 graph tensor [n]."""
    if list(axes) != [0, 1, 2]:
        raise Exception('Axes is not supported in Conditional BatchNorm!')
    c = inputs.shape[-1]
    with S.variable_scope('CondBatchNorm'):
        offset_m = S.get_variable('offset', [n_labels, c], _const(0.))
        scale_m = S.get_variable('scale', [n_labels, c], _const(1.))
    # the statistics pass rides in the producing conv's epilogue when that conv runs the persistent tensor-core kernel
    prod = getattr(inputs.base, 'producer', None)
    stats_from = None
    if (groups == 1 and isinstance(prod, (ConvOp, DeconvOp)) and prod.outputs[0].base is inputs.base and prod.outputs[0] is inputs
            and prod.can_emit_colstats()):
        stats_from = prod
    return BatchNormOp(inputs, scale_m, offset_m, labels, None, True, 1e-5, 0.9, fuse_act, groups=groups, stats_from=stats_from).y


def embed_y(inputs, vocab_size, embedding_dim, word2vec_file=None, name='Embedding.Label'):
    """cifar10/common/ops/embedding.py:12-51: trainable [vocab, dim] table U(-.08,.08) + embedding_lookup.
    inputs None (or the whole vocabulary 0..vocab-1) returns the table itself."""
    assert word2vec_file is None
    with S.variable_scope(name):
        table = S.get_variable('embedding_map', [vocab_size, embedding_dim],
                               lambda shape: (torch.rand(tuple(shape), generator=S.init_generator()) * 2 - 1) * 0.08)
    if inputs is None:
        return table
    return GatherRowsOp(table, inputs).out
