"""Host-side planning logic of the static-program runtime (graph.py / nnops.py) on a CPU-only box: programs are BUILT
and PLANNED here (buffers on the CPU, library loaded for its host-only shape queries); nothing is launched."""
import pytest
import torch

from robust_conditional_gan_b200 import _C, nnops, scope as S
from robust_conditional_gan_b200.graph import Program, VariableStore
from robust_conditional_gan_b200.nnops import (ActOp, AddOp, BatchNormOp, CastOp, ChannelLossOp, ConvOp, DeconvOp, MeanHWOp,
                                               SpectralNormOp, Upsample2Op)

DEV = torch.device('cpu')


@pytest.fixture(autouse=True)
def fused_act_backward_on(monkeypatch):
    """the planning of the (opt-in, RCGAN_FUSE_ACT_BWD=1) activation backward inside the dgrad epilogue is covered here too"""
    monkeypatch.setattr(nnops, 'FUSE_ACT_BWD', True)


def store_with(**shapes):
    st = VariableStore(DEV, lambda n: 'g')
    S.set_store(st, 0)
    vs = {k: st.get(k, s, lambda sh: torch.randn(sh) * 0.1) for k, s in shapes.items()}
    return st, vs


def test_residual_add_aliases_sole_reader_gradient_and_orders_writers():
    st, v = store_with(w1=(3, 3, 64, 64), w2=(3, 3, 64, 64), w3=(3, 3, 64, 64))
    p = Program('t', DEV, _C.BF16)
    with p:
        x = p.input('x', [2, 8, 8, 64], _C.BF16)
        h = ConvOp(x, v['w1'], None).y
        r = ActOp(h, 'relu')
        cop = ConvOp(r.y, v['w2'], None)
        c = cop.y
        z = AddOp(h, c)                      # c has one reader (the add); h has two (relu, add)
        c2 = ConvOp(z.y, v['w3'], None)
        z2 = AddOp(c2.y, z.y)                # here the SECOND operand (z.y) has two readers -> must not be aliased
    st.finalize()
    p.finalize([v['w1'], v['w2'], v['w3']])
    assert z.alias_b and c.grad.data_ptr() == z.y.grad.data_ptr()
    assert not z2.alias_b and z2.y.grad.data_ptr() != z.y.grad.data_ptr()
    # reverse sweep: h's gradient STARTS as the sum's gradient (same buffer, no copy: h's other reader, the relu, precedes c's
    # producer in program order); the relu's backward is fused into the dgrad of the conv that reads it, which accumulates
    # act'(relu output) * gradient straight into h's gradient
    # (not here: the conv that would accumulate into h's gradient is also the one that still reads the buffer as dL/dc)
    assert not z.alias_a and h.grad.data_ptr() != z.y.grad.data_ptr() and z.acc_a == 0
    assert r.bwd_fused and cop.dx_target is h and cop.dx_mask == (_C.ACT_RELU, 0.2) and cop.acc_x == 1
    # parameter gradients always accumulate into the zeroed arena
    assert c2.acc_w == 1 and c2.acc_x == 1 and z2.acc_b == 0


def test_residual_block_keeps_one_gradient_buffer_and_fuses_the_relus():
    """the reference's pre-activation block x -> relu -> Conv1(+relu) -> Conv2 -> (+x) (gan_resnet.py:275-328, D blocks 3-6)"""
    st, v = store_with(w0=(3, 3, 64, 64), w1=(3, 3, 64, 64), w2=(3, 3, 64, 64))
    p = Program('t', DEV, _C.BF16)
    with p:
        x0 = p.input('x', [2, 8, 8, 64], _C.BF16)
        x = ConvOp(x0, v['w0'], None).y
        a = ActOp(x, 'relu')
        c1 = ConvOp(a.y, v['w1'], None, act='relu')
        c2 = ConvOp(c1.y, v['w2'], None)
        z = AddOp(x, c2.y)
        MeanHWOp(z.y)
    st.finalize()
    p.finalize([v['w0'], v['w1'], v['w2']])
    # both branches of the add share dL/dz's buffer: c2 reads it as dL/dy, c1's dgrad later accumulates relu'(a) * g into it as dL/dx
    assert z.alias_a and z.alias_b and x.grad.data_ptr() == z.y.grad.data_ptr() == c2.y.grad.data_ptr()
    assert a.bwd_fused and c1.dx_target is x and c1.acc_x == 1 and c1.dx_mask[0] == _C.ACT_RELU
    # Conv1's own relu is differentiated by Conv2's dgrad as it writes dL/d(c1.y)
    assert c1.act_bwd_fused and c2.dx_mask[0] == _C.ACT_RELU and c2.dx_target is c1.y and c2.acc_x == 0


def test_fused_residual_and_second_relu_output():
    """Conv2's epilogue adds the shortcut and also writes relu(y) for the next block; the next block's Conv1 sends its input
    gradient straight into dL/dy through relu'(.)"""
    st, v = store_with(w0=(3, 3, 64, 64), w1=(3, 3, 64, 64), w2=(3, 3, 64, 64))
    p = Program('t', DEV, _C.BF16)
    with p:
        x0 = p.input('x', [2, 8, 8, 64], _C.BF16)
        c0 = ConvOp(x0, v['w0'], None)
        assert c0.can_emit_relu()
        r = c0.emit_relu()
        c1 = ConvOp(r, v['w1'], None, act='relu')
        c2 = ConvOp(c1.y, v['w2'], None, residual=c0.y)
        MeanHWOp(c2.y)
    st.finalize()
    p.finalize([v['w0'], v['w1'], v['w2']])
    assert c0.y2 is r and c0.y2_bwd_fused and c0.acc_y2 is None
    assert c1.dx_target is c0.y and c1.acc_x == 1 and c1.dx_mask == (_C.ACT_RELU, 0.0)
    assert c2.alias_r and c0.y.grad.data_ptr() == c2.y.grad.data_ptr()


def test_gradients_are_only_planned_for_the_requested_variables():
    st, v = store_with(wg=(3, 3, 64, 64), wd=(3, 3, 64, 64))
    p = Program('t', DEV, _C.BF16)
    with p:
        x = p.input('x', [2, 8, 8, 64], _C.BF16)
        g = ConvOp(x, v['wg'], None)         # "generator"
        d = ConvOp(g.y, v['wd'], None)       # "discriminator"
        MeanHWOp(d.y)
    st.finalize()
    p.finalize([v['wd']])                    # D step: var_list = discriminator variables only
    assert d.need[:3] == [False, True, False] and g.need[:3] == [False, False, False]
    assert g.y.grad is None                  # no buffer is allocated for a gradient nobody needs
    p2 = Program('t2', DEV, _C.BF16)
    with p2:
        x = p2.input('x', [2, 8, 8, 64], _C.BF16)
        g = ConvOp(x, v['wg'], None)
        d = ConvOp(g.y, v['wd'], None)
        m = MeanHWOp(d.y)
        h32 = CastOp(m.y, _C.F32).y
        eye = p2.input('eye', [2, 64])
        wgt = p2.input('w', [2, 2])
        ChannelLossOp(h32, None, eye, wgt, _C.HINGE_G, 'g_loss')
    p2.finalize([v['wg']])                   # G step: D is differentiated w.r.t. its INPUT only
    assert d.need[:3] == [True, False, False] and g.need[:3] == [False, True, False]


def test_conv_path_selection_by_shape():
    st, v = store_with(w_img=(5, 5, 1, 64), w_cat=(5, 5, 11, 64), w_out=(3, 3, 256, 3), w_big=(3, 3, 256, 256), w_up=(3, 3, 256, 256))
    p = Program('t', DEV, _C.BF16)
    with p:
        img = p.input('img', [4, 28, 28, 1], _C.BF16)
        cat = p.new((4, 28, 28, 11), _C.BF16, ld=16)
        feat = p.input('feat', [4, 16, 16, 256], _C.BF16)
        a = ConvOp(img, v['w_img'], None, stride=2)      # d_h0_conv: patch-matrix GEMM + scatter dgrad
        b = ConvOp(cat, v['w_cat'], None, stride=2)      # --concat_y first layer: patch GEMM + implicit-GEMM dgrad pack
        c = ConvOp(feat, v['w_out'], None)               # G.Output: transposed patch path for the backward
        d = ConvOp(feat, v['w_big'], None)               # ordinary tensor-core layer
        up = Upsample2Op(feat)
        e = ConvOp(up.y, v['w_up'], None, up_op=up)      # UpsampleConv pair
    st.finalize()
    p.finalize([])                                       # nothing differentiated: the pair folds onto the small input
    assert a.patch is not None and a.scatter.ok and a.dpack is None
    assert b.patch is not None and not b.scatter.ok and b.dpack is not None
    assert c.patch is None and c.tpatch is not None
    assert d.patch is None and d.tpatch is None and d.pack is not None
    assert e.up_op is up and e.folds_upsample()
    p2 = Program('t2', DEV, _C.BF16)
    with p2:
        feat = p2.input('feat', [4, 16, 16, 256], _C.BF16)
        up = Upsample2Op(feat)
        e = ConvOp(up.y, v['w_up'], None, up_op=up)
    p2.finalize([v['w_up']])                             # differentiated: upsample + ordinary conv
    assert not e.folds_upsample()


def test_sliced_channel_loss_zeroes_once_and_grouped_bn_keeps_two_statistics():
    st, v = store_with(w=(3, 3, 64, 64), gamma=(64,), beta=(64,))
    p = Program('t', DEV, _C.BF16)
    with p:
        x = p.input('x', [8, 4, 4, 64], _C.BF16)                  # [real; fake], 4 + 4 samples
        c = ConvOp(x, v['w'], None)
        bn = BatchNormOp(c.y, v['gamma'], v['beta'], act='lrelu', groups=2)
        h = CastOp(MeanHWOp(bn.y).y, _C.F32).y
        V = p.input('V', [10, 64])
        wr, wf = p.input('wr', [4, 10]), p.input('wf', [4, 10])
        lr = ChannelLossOp(h, None, V, wr, _C.HINGE_D_REAL, 'real', rows=(0, 4))
        lf = ChannelLossOp(h, None, V, wf, _C.HINGE_D_FAKE, 'fake', rows=(4, 4))
    st.finalize()
    p.finalize([v['w'], v['gamma'], v['beta']])
    assert bn.groups == 2 and bn.samples == 4 and len(bn.save) == 2
    # reverse order: the fake term runs first and clears the whole gradient once, both then accumulate their rows
    assert lf.zero_h and not lr.zero_h and lf.acc_h == 1 and lr.acc_h == 1
    assert p.loss_names == ['real', 'fake']


def test_spectral_norm_ops_batch_only_parameters():
    st, v = store_with(W=(3, 3, 64, 64), u=(1, 64))
    p = Program('t', DEV, _C.BF16)
    with p:
        a = SpectralNormOp(v['W'], v['u'])
        w2 = CastOp(a.wbar, _C.F32).y                              # a W produced by another op is not available up front
        b = SpectralNormOp(w2, v['u'], update=False)
    assert a.batched and not b.batched


def test_derived_weight_gradients_are_zeroed_once_per_sweep_and_accumulated():
    """A spectral-normed filter's gradient buffer belongs to the program, not to the parameter arena: its first writer (the conv's
    wgrad) registers it for the one batched zeroing at the start of the backward sweep and accumulates, so no memset node sits in
    front of each wgrad; a plain parameter's gradient keeps accumulating into the arena."""
    st, v = store_with(W=(3, 3, 64, 64), u=(1, 64), w_plain=(3, 3, 64, 64))
    p = Program('t', DEV, _C.BF16)
    with p:
        x = p.input('x', [2, 8, 8, 64], _C.BF16)
        sn = SpectralNormOp(v['W'], v['u'])
        c1 = ConvOp(x, sn.wbar, None)
        c2 = ConvOp(c1.y, sn.wbar, None)           # the same normalised filter used twice: the second wgrad accumulates too
        c3 = ConvOp(c2.y, v['w_plain'], None)
        MeanHWOp(c3.y)
    st.finalize()
    p.finalize([v['W'], v['w_plain']])
    assert [t is sn.wbar.base for t in p.prezero] == [True]
    assert c1.acc_w == 1 and c2.acc_w == 1 and c3.acc_w == 1
    assert sn.wbar.base.grad_written
