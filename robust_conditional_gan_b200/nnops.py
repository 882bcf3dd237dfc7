"""Op classes of the static program: each maps one reference op (cited per class) onto
forward/backward launches of librcgan_b200.so.  See graph.py for the planning protocol."""
import ctypes
import os

import torch

from . import _C
from ._C import ConvDesc, call, stream_ptr
from .graph import Op, Tensor, cur, is_static_weight, needs, round_up, same_pad

# A/B switches of the fused-epilogue planning.  The activation backward inside the dgrad epilogue is bit-identical but measured
# SLOWER (25.6 vs 23.8 ms per CIFAR iteration): its extra operand is a per-lane global load in a 4-warp epilogue, which cannot
# keep enough bytes in flight for the short-K discriminator layers (the folded stride-2 conv has K = 512 per output tile and
# would need ~35 GB/s per SM) -- it needs a TMA-staged mask tile; off by default until then
FUSE_ACT_BWD = os.environ.get('RCGAN_FUSE_ACT_BWD', '0') == '1'
ALIAS_RESIDUAL = os.environ.get('RCGAN_ALIAS_RESIDUAL', '1') == '1'
FUSE_RELU_OUT = os.environ.get('RCGAN_FUSE_RELU_OUT', '1') == '1'
FUSE_BN_STATS = os.environ.get('RCGAN_FUSE_BN_STATS', '1') == '1'
BN_MASK_FROM_X = os.environ.get('RCGAN_BN_MASK_FROM_X', '0') == '1'
# a conv's filter gradient on the side stream, next to its input gradient (both only read dL/dy; joined at the end of the op)
FORK_WGRAD = os.environ.get('RCGAN_FORK_WGRAD', '0') == '1'   # measured neutral (19.181 vs 19.193 ms): opt-in

ACT = {None: _C.ACT_NONE, 'none': _C.ACT_NONE, 'relu': _C.ACT_RELU, 'lrelu': _C.ACT_LRELU, 'sigmoid': _C.ACT_SIGMOID,
       'tanh': _C.ACT_TANH}


def dp(t):
    """device pointer of a graph tensor's data (None -> NULL)"""
    return None if t is None else t.data.data_ptr()


def pp(buf):
    """device pointer of a raw torch buffer (None -> NULL)"""
    return None if buf is None else buf.data_ptr()


def gp(t):
    return None if t is None or t.grad is None else t.grad.data_ptr()


def make_patch(prog, desc, rows, ld_out):
    """(patch tensor, GEMM desc) for a bf16 conv with few input channels, else (None, None).
    The GEMM desc is the 1x1 conv over the [rows, 1, 1, K'] patch matrix with the same weights viewed as [K', cout]."""
    kp = desc.kh * desc.kw * desc.cin
    # cin <= 4 (1- / 3-channel images), or the 11-channel image + label concat of --concat_y (run_rcgany.sh: 5x5x11 = 275
    # columns) -- below the 32 channels the implicit-GEMM kernel needs per K block
    if desc.dtype != _C.BF16 or desc.cin > 16 or kp > 512 or (desc.cin > 4 and kp <= 64) or desc.cout < 32 or desc.ldy % 8 != 0:
        return None, None
    ldp = round_up(kp, 8)
    patch = prog.new((rows, kp), _C.BF16, ld=ldp)
    g = ConvDesc(rows, 1, 1, kp, 1, 1, desc.cout, 1, 1, 1, 0, 0, ldp, ld_out, desc.dtype)
    return patch, g


class ScatterDgrad:
    """Input gradient of a conv with cin <= 4 (equivalently the forward of the conv2d_transpose it defines) as ONE dense
    GEMM T = dy[M, cout] * W^T[cout, kh*kw*cin] on the tensor cores followed by rcgan_col2im: dy is read once instead of
    once per filter tap.  None-like (ok == False) when the shape does not qualify."""

    def __init__(self, prog, desc, ld_dy):
        kp = desc.kh * desc.kw * desc.cin
        rows = desc.n * desc.ho * desc.wo
        self.ok = desc.dtype == _C.BF16 and desc.cin <= 4 and kp <= 128 and desc.cout >= 32 and ld_dy % 8 == 0
        if not self.ok:
            return
        self.desc, self.kp = desc, kp
        self.ldt = round_up(kp, 8)
        self.g = ConvDesc(rows, 1, 1, desc.cout, 1, 1, kp, 1, 1, 1, 0, 0, ld_dy, self.ldt, desc.dtype)
        nbytes = _C.load().rcgan_conv_wpack_bytes(self.g)
        self.ok = nbytes > 0
        if not self.ok:
            return
        self.wt = torch.zeros(kp * desc.cout, dtype=torch.float32, device=prog.device)
        self.pack = torch.zeros(nbytes, dtype=torch.uint8, device=prog.device)
        self.T = torch.zeros(rows * self.ldt, dtype=torch.float32, device=prog.device)

    def run(self, w_ptr, dy_ptr, out_ptr, out_dtype, bias_ptr, act, leak, accumulate, st):
        d = self.desc
        # W viewed as [kh*kw*cin, cout] -> W^T [cout, kh*kw*cin] = the 1x1 filter of the GEMM
        call('rcgan_wflip', w_ptr, pp(self.wt), 1, 1, self.kp, d.cout, 0, st)
        call('rcgan_conv_wpack', self.g, pp(self.wt), None, pp(self.pack), st)
        call('rcgan_conv2d_fprop', self.g, dy_ptr, pp(self.wt), pp(self.pack), None, pp(self.T), _C.F32, _C.ACT_NONE, 0.0, st)
        call('rcgan_col2im', d, pp(self.T), self.ldt, bias_ptr, out_ptr, out_dtype, act, leak, accumulate, st)


def spatial(t):
    """(n, h, w) of a 2-D [n,c] or 4-D [n,h,w,c] tensor"""
    if len(t.shape) == 2:
        return t.shape[0], 1, 1
    assert len(t.shape) == 4, t.shape
    return t.shape[0], t.shape[1], t.shape[2]


def derive(raw, out):
    """`out` = act(raw): whoever reads `out` may end up writing raw's gradient directly (fused activation backward)"""
    if not hasattr(raw.base, 'derived'):
        raw.base.derived = []
    raw.base.derived.append(out.base)


def grad_writer_ops(t):
    """every op whose backward can write t's gradient: t's readers and the readers of the activations derived from t"""
    ops = list(getattr(t.base, 'readers', []))
    for u in getattr(t.base, 'derived', []):
        ops += list(getattr(u, 'readers', []))
    return ops


class ConvOp(Op):
    """y = act(conv2d_SAME(x, w, stride) + b).  tf.nn.conv2d + bias_add (mnist/ops.py:53-67;
    cifar10/common/ops/conv2d.py:181-216); with a 2-D x it is tf.matmul + bias
    (mnist/ops.py:97-116; cifar10/common/ops/linear.py:161-180).  w: [kh,kw,cin,cout] or [cin,cout]."""

    def __init__(self, x, w, b, stride=1, act=None, leak=0.2, pre_norm=False, residual=None, up_op=None, residual_up=False):
        """pre_norm: the output feeds a batch norm -> stored in fp32 even in bf16 mode (gradient stays bf16).
        residual: y = conv(x) + b + residual, the ResidualBlock's shortcut add fused into this conv's epilogue;
        residual_up: the residual is [n, ho/2, wo/2, cout] and is added through a nearest-neighbour 2x upsampling (the 1x1
        UpsampleConv shortcut computed on the small grid, never materialised at the output resolution).
        up_op: the Upsample2Op that produced x (UpsampleConv): in programs that need no gradient through this conv the
        pair runs as ONE folded launch on the small input (rcgan_upconv2d_fprop) and the upsample is skipped."""
        prog = cur()
        n, h, wd = spatial(x)
        if len(w.shape) == 2:
            kh = kw = 1
            cin, cout = w.shape
        else:
            kh, kw, cin, cout = w.shape
        assert cin == x.c, 'conv: weight cin %d != input channels %d' % (cin, x.c)
        ho, pt = same_pad(h, kh, stride)
        wo, pl = same_pad(wd, kw, stride)
        self.x, self.w, self.b, self.res = x, w, b, residual
        self.act, self.leak = ACT[act], leak
        self.act_bwd_fused = False       # set by the consumer whose dgrad applies act'(y) while writing y's gradient
        oshape = (n, cout) if len(x.shape) == 2 else (n, ho, wo, cout)
        assert not (pre_norm and self.act != _C.ACT_NONE)
        self.res_up = bool(residual_up)
        rshape = (n, ho // 2, wo // 2, cout) if residual_up else oshape
        assert residual is None or (self.act == _C.ACT_NONE and not pre_norm and tuple(residual.shape) == tuple(rshape)
                                    and residual.ld == residual.c and residual.dtype == x.dtype)
        assert not residual_up or (residual is not None and ho % 2 == 0 and wo % 2 == 0)
        self.y2 = None                   # second output relu(y), attached on demand (emit_relu)
        self.y2_bwd_fused = False
        self.y = prog.new(oshape, _C.F32 if pre_norm else x.dtype, grad_dtype=x.dtype)
        self.desc = ConvDesc(n, h, wd, cin, ho, wo, cout, kh, kw, stride, pt, pl, x.ld, self.y.ld, x.dtype)
        self.inputs, self.outputs = (x, w, b, residual), (self.y,)
        prog.ws.request(_C.load().rcgan_conv2d_wgrad_workspace(self.desc))
        self.up_op = None
        if up_op is not None and residual is None and not pre_norm and _C.load().rcgan_upconv2d_pack_bytes(self.desc) > 0:
            self.up_op = up_op
            up_op.consumer = self
            self.up_wfold = torch.zeros(16 * cin * cout, dtype=torch.float32, device=prog.device)
            self.up_pack = torch.zeros(_C.load().rcgan_upconv2d_pack_bytes(self.desc), dtype=torch.uint8, device=prog.device)
        # few-channel inputs (cin <= 4): materialise the patch matrix once and run fprop / wgrad as dense GEMMs on it
        self.patch, self.gdesc = make_patch(prog, self.desc, n * ho * wo, self.y.ld)
        self.pack, self.pack_owner = prog.weight_pack(w, self.gdesc if self.patch is not None else self.desc)
        if self.pack_owner and prog.hoist_pack(w, self.gdesc if self.patch is not None else self.desc, self.pack):
            self.pack_owner = False          # refreshed with every other pack of the program at its start
        # ... and their input gradient (a 1..3-channel result from a wide dL/dy: d_h0_conv / D.Block.1 in the G step) as one
        # dense GEMM + col2im
        self.scatter = ScatterDgrad(prog, self.desc, self.y.ld) if self.patch is not None else None
        # 5..16 input channels: too wide for the scatter GEMM's tap matrix; the implicit-GEMM dgrad needs its own pack of the
        # TRUE conv geometry (self.pack belongs to the patch GEMM)
        self.dpack = None
        if self.patch is not None and not self.scatter.ok and _C.load().rcgan_conv_uses_tensor_cores(self.desc, 1):
            self.dpack = torch.zeros(_C.load().rcgan_conv_wpack_bytes(self.desc), dtype=torch.uint8, device=prog.device)
        # few-channel OUTPUTS (G.Output: 256 -> 3): the backward is the transposed conv of dL/dy (cin' = cout <= 4) with the
        # flipped filter, so it runs as GEMMs on the patch matrix of dL/dy (rcgan_wflip in the header)
        self.tpatch = None
        if stride == 1 and cout <= 4 and cin >= 32 and self.patch is None:
            self.tdesc = ConvDesc(n, ho, wo, cout, h, wd, cin, kh, kw, 1, kh - 1 - pt, kw - 1 - pl, self.y.ld, x.ld, x.dtype)
            self.tpatch, self.tgdesc = make_patch(prog, self.tdesc, n * h * wd, x.ld)
        if self.tpatch is not None:
            # ... and the forward as ONE dense GEMM T = x[M, cin] * W2[cin, kh*kw*cout] plus rcgan_col2im with the flipped tap
            # order (x is read once instead of once per filter tap)
            kp = kh * kw * cout
            self.f_ldt = round_up(kp, 8)
            self.f_g = ConvDesc(n * h * wd, 1, 1, cin, 1, 1, kp, 1, 1, 1, 0, 0, x.ld, self.f_ldt, x.dtype)
            self.f_cdesc = ConvDesc(n, ho, wo, cout, h, wd, cin, kh, kw, 1, kh - 1 - pt, kw - 1 - pl, self.y.ld, x.ld, x.dtype)
            self.f_w2 = torch.zeros(kp * cin, dtype=torch.float32, device=prog.device)
            self.f_pack = torch.zeros(_C.load().rcgan_conv_wpack_bytes(self.f_g), dtype=torch.uint8, device=prog.device)
            self.f_T = torch.zeros(n * h * wd * self.f_ldt, dtype=torch.float32, device=prog.device)
            nflip = kh * kw * cout * cin
            self.wflip = torch.zeros(nflip, dtype=torch.float32, device=prog.device)
            self.dwflip = torch.zeros(nflip, dtype=torch.float32, device=prog.device)
            self.tpack = torch.zeros(_C.load().rcgan_conv_wpack_bytes(self.tgdesc), dtype=torch.uint8, device=prog.device)
            prog.ws.request(_C.load().rcgan_conv2d_wgrad_workspace(self.tgdesc))
        prog.add(self)

    def _plain_tc_fprop(self):
        """the forward is ONE tensor-core fprop launch with a bf16 result: the fused epilogues (rcgan_conv2d_fprop_ex) apply"""
        return (self.x.dtype == _C.BF16 and self.y.dtype == _C.BF16 and self.pack is not None and self.patch is None
                and self.tpatch is None and self.up_op is None and bool(_C.load().rcgan_conv_uses_tensor_cores(self.desc, 0)))

    def attach_residual(self, residual, up=False):
        """y = conv(x) + b + [upsample2](residual), decided right after construction (the caller first asks _plain_tc_fprop())"""
        prog = cur()
        assert self.res is None and prog.ops[-1] is self and self.act == _C.ACT_NONE and self.y.dtype == self.x.dtype
        n, ho, wo = spatial(self.y)
        rshape = (n, ho // 2, wo // 2, self.y.c) if up else tuple(self.y.shape)
        assert tuple(residual.shape) == rshape and residual.ld == residual.c and residual.dtype == self.x.dtype
        assert not up or (ho % 2 == 0 and wo % 2 == 0 and self._plain_tc_fprop())
        self.res, self.res_up = residual, bool(up)
        self.inputs = (self.x, self.w, self.b, residual)

    def can_emit_colstats(self):
        """the following batch norm's statistics pass can ride in this conv's (persistent-kernel) epilogue"""
        # (only where the dispatch picks the persistent kernel anyway, >= 3 big tiles per SM: forcing it on a small layer costs more
        # than the statistics pass it saves)
        big_tiles = -(-self.y.rows // 128) if self.y.c > 128 else -(-self.y.rows // 256)
        return (FUSE_BN_STATS and getattr(self, 'colstats', None) is None and self._plain_tc_fprop() and self.y.c in (64, 128, 256)
                and self.y.ld == self.y.c and big_tiles >= 3 * 148)

    def emit_colstats(self):
        assert self.can_emit_colstats()
        self.colstats = torch.zeros(_C.load().rcgan_colstats_floats(self.y.c), dtype=torch.float32, device=cur().device)
        self.ep_fwd = None
        return self.colstats

    def can_emit_relu(self):
        return FUSE_RELU_OUT and self.y2 is None and self.act == _C.ACT_NONE and self._plain_tc_fprop()

    def emit_relu(self):
        """second output relu(y) written by this conv's epilogue (the next block's `nonlinearity(inputs)`, gan_resnet.py:318)"""
        prog = cur()
        assert self.can_emit_relu() and prog.ops[self.index] is self
        self.y2 = prog.new(self.y.shape, self.y.dtype)
        self.y2.base.producer = self
        self.outputs = (self.y, self.y2)
        derive(self.y, self.y2)
        return self.y2

    def _plain_tc_dgrad(self):
        """the input gradient is ONE tensor-core dgrad launch with a bf16 result (no patch / scatter / transposed special path)"""
        return (self.x.dtype == _C.BF16 and self.x.grad_dtype == _C.BF16 and self.pack is not None and self.patch is None
                and self.tpatch is None and bool(_C.load().rcgan_conv_uses_tensor_cores(self.desc, 1)))

    def plan_bwd(self, prog):
        nx, nw, nb, nr = self.need
        # (reverse program order: the residual's gradient is claimed first, as the separate AddOp used to do)
        self.alias_r = False
        if nr and self.res is not None and not self.res_up:
            # same-size residual: its gradient STARTS as dL/dy -> share the buffer when everything else that writes it runs
            # after this op's backward (i.e. precedes this op in program order); see AddOp.plan_bwd
            r, y = self.res.base, self.y.base
            self.alias_r = bool(ALIAS_RESIDUAL and needs(self.y) and r is self.res and y is self.y and not r.is_variable
                                and not r.grad_written and r._grad is not None and y._grad is not None
                                and r._grad.numel() == y._grad.numel() and r._grad.dtype == y._grad.dtype
                                and all(o.index < self.index for o in grad_writer_ops(r) if o is not self))
        if self.alias_r:
            self.res.base._grad = self.y.base._grad
            self.res.base.grad_written = True
            self.acc_r = 0
        else:
            self.acc_r = self.claim(self.res) if (nr and self.res is not None) else 0
        # relu(y) as second output whose consumer did not fuse the relu's backward: fold it into dL/dy first thing in backward
        self.acc_y2 = self.claim(self.y) if (self.y2 is not None and needs(self.y2) and not self.y2_bwd_fused) else None
        # The activation IN FRONT of this conv differentiated inside the dgrad that produces its gradient (rcgan_conv_epilogue.mask):
        # x = act(raw) was written by an ActOp, or by the fused epilogue activation of the producing conv, and this conv is its only
        # reader -> the dgrad multiplies by act'(x) as it stores; for an ActOp it stores straight into raw's gradient and the ActOp
        # has no backward pass left.  Results are bit-identical to the separate act_bwd kernel (same roundings).
        self.dx_target, self.dx_mask = self.x, None
        if nx and FUSE_ACT_BWD and self._plain_tc_dgrad() and getattr(self.x.base, 'n_readers', 0) == 1:
            prod = getattr(self.x.base, 'producer', None)
            if (isinstance(prod, ActOp) and prod.act in (_C.ACT_RELU, _C.ACT_LRELU) and prod.need[0] and prod.x.ld == self.x.ld
                    and prod.x.grad_dtype == _C.BF16 and not prod.x.is_variable):
                self.dx_target, self.dx_mask = prod.x, (prod.act, prod.leak)
                prod.bwd_fused = True
            elif isinstance(prod, ConvOp) and prod.y2 is not None and prod.y2.base is self.x.base:
                self.dx_target, self.dx_mask = prod.y, (_C.ACT_RELU, 0.0)
                prod.y2_bwd_fused = True
            elif (isinstance(prod, (ConvOp, DeconvOp)) and prod.act in (_C.ACT_RELU, _C.ACT_LRELU) and prod.y.base is self.x.base
                  and prod.y.dtype == _C.BF16):
                self.dx_mask = (prod.act, prod.leak)
                prod.act_bwd_fused = True
        self.acc_x = self.claim(self.dx_target) if nx else 0
        if self.dx_mask is not None:
            self.dx_ep = _C.ConvEpilogue(mask=dp(self.x), mask_act=self.dx_mask[0], mask_leak=self.dx_mask[1])
        self.acc_w = self.claim_wgrad(prog, self.w) if nw else 0
        self.acc_b = self.claim(self.b) if (nb and self.b is not None) else 0

    def _backward_transposed(self, prog, nx, nw, dy, st):
        d = self.desc
        call('rcgan_im2col', self.tdesc, dy, dp(self.tpatch), self.tpatch.ld, st)
        if nx:
            call('rcgan_wflip', dp(self.w), pp(self.wflip), d.kh, d.kw, d.cin, d.cout, 0, st)
            call('rcgan_conv_wpack', self.tgdesc, pp(self.wflip), None, pp(self.tpack), st)
            call('rcgan_conv2d_fprop', self.tgdesc, dp(self.tpatch), pp(self.wflip), pp(self.tpack), None, gp(self.x),
                 self.x.grad_dtype, _C.ACT_NONE, 0.0, st)
        if nw:
            # d(flipped filter)[k', ci] = sum_m patch[m, k'] * x[m, ci]: the GEMM desc's wgrad with x in the dy role
            call('rcgan_conv2d_wgrad', self.tgdesc, dp(self.tpatch), dp(self.x), pp(self.dwflip), 0, prog.ws.ptr(), prog.ws.bytes, st)
            call('rcgan_wflip', pp(self.dwflip), gp(self.w), d.kh, d.kw, d.cout, d.cin, self.acc_w, st)

    def _forward_scatter(self, prog):
        d, st = self.desc, stream_ptr()
        kp = d.kh * d.kw * d.cout
        # wflip: [tap'][co][ci] = w[flipped tap'][ci][co]; transposed once more it is the GEMM filter [cin][tap'*cout + co]
        call('rcgan_wflip', dp(self.w), pp(self.wflip), d.kh, d.kw, d.cin, d.cout, 0, st)
        call('rcgan_wflip', pp(self.wflip), pp(self.f_w2), 1, 1, kp, d.cin, 0, st)
        call('rcgan_conv_wpack', self.f_g, pp(self.f_w2), None, pp(self.f_pack), st)
        call('rcgan_conv2d_fprop', self.f_g, dp(self.x), pp(self.f_w2), pp(self.f_pack), None, pp(self.f_T), _C.F32, _C.ACT_NONE, 0.0, st)
        call('rcgan_col2im', self.f_cdesc, pp(self.f_T), self.f_ldt, dp(self.b), dp(self.y), self.y.dtype, self.act, self.leak, 0, st)

    def folds_upsample(self):
        """UpsampleConv without the upsampled tensor: legal when nothing of this conv is differentiated in this program."""
        return self.up_op is not None and not any(self.need) and not needs(self.y)

    def forward(self, prog):
        if self.folds_upsample():
            st = stream_ptr()
            call('rcgan_upconv2d_fold', self.desc, dp(self.w), pp(self.up_wfold), pp(self.up_pack), st)
            call('rcgan_upconv2d_fprop', self.desc, dp(self.up_op.x), pp(self.up_pack), dp(self.b), dp(self.y), self.y.dtype,
                 self.act, self.leak, st)
            return
        if self.tpatch is not None and self.desc.kh == self.desc.kw and self.desc.pad_t == self.desc.pad_l:
            return self._forward_scatter(prog)
        d, xin = self.desc, dp(self.x)
        if self.patch is not None:
            call('rcgan_im2col', self.desc, dp(self.x), dp(self.patch), self.patch.ld, stream_ptr())
            d, xin = self.gdesc, dp(self.patch)
        if self.pack_owner:
            call('rcgan_conv_wpack', d, dp(self.w), None, pp(self.pack), stream_ptr())
        colstats = getattr(self, 'colstats', None)
        if self.res_up or self.y2 is not None or colstats is not None or (self.res is not None and self._plain_tc_fprop()):
            if getattr(self, 'ep_fwd', None) is None:
                self.ep_fwd = _C.ConvEpilogue(res=dp(self.res), res_up=int(self.res_up), ld_res=self.res.ld if self.res is not None else 0,
                                              out2=dp(self.y2), out2_act=_C.ACT_RELU if self.y2 is not None else 0,
                                              colstats=pp(colstats))
            call('rcgan_conv2d_fprop_ex', d, xin, pp(self.pack), dp(self.b), dp(self.y), self.y.dtype, self.act, self.leak,
                 ctypes.byref(self.ep_fwd), stream_ptr())
            return
        if self.res is not None:
            call('rcgan_conv2d_fprop_res', d, xin, dp(self.w), pp(self.pack), dp(self.b), dp(self.res), dp(self.y), self.y.dtype,
                 self.act, self.leak, stream_ptr())
            return
        call('rcgan_conv2d_fprop', d, xin, dp(self.w), pp(self.pack), dp(self.b), dp(self.y), self.y.dtype, self.act,
             self.leak, stream_ptr())

    def backward(self, prog):
        if not needs(self.y):
            return
        nx, nw, nb, nr = self.need
        st = stream_ptr()
        y, dy = self.y, gp(self.y)
        if self.acc_y2 is not None:
            y2 = self.y2
            call('rcgan_act_bwd', gp(y2), dp(y2), dy, y2.rows, y2.c, y2.ld, y2.ld, y.ld, y2.dtype, _C.ACT_RELU, 0.0, self.acc_y2, st)
        if nr and self.res is not None and self.res_up:
            rn, rh, rw = spatial(self.res)
            call('rcgan_upsample2_bwd', dy, gp(self.res), rn, rh, rw, self.res.c, self.res.grad_dtype, self.acc_r, st)
        elif nr and self.res is not None and not self.alias_r:
            call('rcgan_copy_acc', dy, gp(self.res), self.res.numel(), self.res.grad_dtype, self.acc_r, st)
        bias_done = False
        if self.act != _C.ACT_NONE and not self.act_bwd_fused:
            if nb and self.b is not None and y.dtype == y.grad_dtype:
                # activation backward and the bias gradient's column sums in one pass over dL/dy
                call('rcgan_act_bwd_colsum', dy, dp(y), y.rows, y.c, y.ld, y.ld, y.dtype, self.act, self.leak, gp(self.b), self.acc_b, st)
                bias_done = True
            else:
                call('rcgan_act_bwd', dy, dp(y), dy, y.rows, y.c, y.ld, y.ld, y.ld, y.dtype, self.act, self.leak, 0, st)
        if nb and self.b is not None and not bias_done and torch.cuda.is_available():
            # the bias gradient's column sums only read dL/dy: on the side stream, next to this layer's dgrad / wgrad
            prog.fork(lambda: call('rcgan_colsum', dy, y.rows, y.c, y.ld, y.grad_dtype, gp(self.b), self.acc_b, stream_ptr()))
            bias_done = True
        if self.tpatch is not None and (nx or nw) and not (nx and self.acc_x):
            self._backward_transposed(prog, nx, nw, dy, st)
            nx = nw = False
        def wgrad():
            s = stream_ptr()
            if self.patch is not None:
                call('rcgan_conv2d_wgrad', self.gdesc, dp(self.patch), dy, gp(self.w), self.acc_w, prog.ws.ptr(), prog.ws.bytes, s)
            else:
                call('rcgan_conv2d_wgrad', self.desc, dp(self.x), dy, gp(self.w), self.acc_w, prog.ws.ptr(), prog.ws.bytes, s)
        if nw and nx and FORK_WGRAD and torch.cuda.is_available():
            prog.fork(wgrad)            # wgrad || dgrad: independent outputs, the small layers do not fill the machine alone
            nw = False
        if nx:
            if self.dx_mask is not None:
                call('rcgan_conv2d_dgrad_ex', self.desc, dy, pp(self.pack), None, gp(self.dx_target), self.x.grad_dtype, _C.ACT_NONE,
                     0.0, self.acc_x, ctypes.byref(self.dx_ep), st)
            elif self.scatter is not None and self.scatter.ok:
                self.scatter.run(dp(self.w), dy, gp(self.x), self.x.grad_dtype, None, _C.ACT_NONE, 0.0, self.acc_x, st)
            elif self.dpack is not None:
                call('rcgan_conv_wpack', self.desc, dp(self.w), None, pp(self.dpack), st)
                call('rcgan_conv2d_dgrad', self.desc, dy, dp(self.w), pp(self.dpack), None, gp(self.x), self.x.grad_dtype,
                     _C.ACT_NONE, 0.0, self.acc_x, st)
            else:
                call('rcgan_conv2d_dgrad', self.desc, dy, dp(self.w), None if self.patch is not None else pp(self.pack), None,
                     gp(self.x), self.x.grad_dtype, _C.ACT_NONE, 0.0, self.acc_x, st)
        if nw:
            wgrad()
        if nb and self.b is not None and not bias_done:
            call('rcgan_colsum', dy, y.rows, y.c, y.ld, y.grad_dtype, gp(self.b), self.acc_b, st)


class DeconvOp(Op):
    """y = act(conv2d_transpose_SAME(x, w, stride) + b), w: [kh,kw,cout,cin] (mnist/ops.py:69-92).
    Forward = the dgrad of the conv (cin_conv = cout, cout_conv = cin) whose input is y."""

    def __init__(self, x, w, b, out_hw, stride=2, act=None, leak=0.2, pre_norm=False):
        prog = cur()
        n, h, wd = spatial(x)
        kh, kw, cout, cin = w.shape
        assert cin == x.c
        oh, ow = out_hw
        ho, pt = same_pad(oh, kh, stride)
        wo, pl = same_pad(ow, kw, stride)
        assert (ho, wo) == (h, wd), 'deconv2d: output_shape inconsistent with SAME/stride'
        self.x, self.w, self.b = x, w, b
        self.act, self.leak = ACT[act], leak
        self.act_bwd_fused = False
        assert not (pre_norm and self.act != _C.ACT_NONE)
        self.y = prog.new((n, oh, ow, cout), _C.F32 if pre_norm else x.dtype, grad_dtype=x.dtype)
        # the conv being transposed: input = y [n,oh,ow,cout], output = x [n,h,w,cin]
        self.desc = ConvDesc(n, oh, ow, cout, h, wd, cin, kh, kw, stride, pt, pl, self.y.ld, x.ld, x.dtype)
        self.inputs, self.outputs = (x, w, b), (self.y,)
        prog.ws.request(_C.load().rcgan_conv2d_wgrad_workspace(self.desc))
        # the conv being transposed has cin = deconv cout: for a 1-channel image (g_h3) its fprop / wgrad (= this op's backward)
        # run as GEMMs on the patch matrix of dL/dy
        self.patch, self.gdesc = make_patch(prog, self.desc, n * h * wd, x.ld)
        self.pack, self.pack_owner = prog.weight_pack(w, self.gdesc if self.patch is not None else self.desc)
        if self.pack_owner and prog.hoist_pack(w, self.gdesc if self.patch is not None else self.desc, self.pack):
            self.pack_owner = False
        self.scatter = ScatterDgrad(prog, self.desc, x.ld) if self.patch is not None else None
        prog.add(self)

    def can_emit_colstats(self):
        return (FUSE_BN_STATS and getattr(self, 'colstats', None) is None and self.x.dtype == _C.BF16 and self.y.dtype == _C.BF16
                and self.pack is not None and self.patch is None and self.y.c in (64, 128, 256) and self.y.ld == self.y.c
                and bool(_C.load().rcgan_conv_uses_tensor_cores(self.desc, 1)))

    def emit_colstats(self):
        assert self.can_emit_colstats()
        self.colstats = torch.zeros(_C.load().rcgan_colstats_floats(self.y.c), dtype=torch.float32, device=cur().device)
        self.ep_fwd = _C.ConvEpilogue(colstats=pp(self.colstats))
        return self.colstats

    def plan_bwd(self, prog):
        nx, nw, nb = self.need
        self.acc_x = self.claim(self.x) if nx else 0
        assert self.acc_x == 0, 'deconv input gradient must have a single writer'
        self.acc_w = self.claim_wgrad(prog, self.w) if nw else 0
        self.acc_b = self.claim(self.b) if (nb and self.b is not None) else 0

    def forward(self, prog):
        if self.pack_owner:
            call('rcgan_conv_wpack', self.gdesc if self.patch is not None else self.desc, dp(self.w), None, pp(self.pack),
                 stream_ptr())
        if self.scatter is not None and self.scatter.ok:
            self.scatter.run(dp(self.w), dp(self.x), dp(self.y), self.y.dtype, dp(self.b), self.act, self.leak, 0, stream_ptr())
            return
        if getattr(self, 'colstats', None) is not None:
            call('rcgan_conv2d_dgrad_ex', self.desc, dp(self.x), pp(self.pack), dp(self.b), dp(self.y), self.y.dtype, self.act, self.leak,
                 0, ctypes.byref(self.ep_fwd), stream_ptr())
            return
        call('rcgan_conv2d_dgrad', self.desc, dp(self.x), dp(self.w), None if self.patch is not None else pp(self.pack),
             dp(self.b), dp(self.y), self.y.dtype, self.act,
             self.leak, 0, stream_ptr())

    def backward(self, prog):
        if not needs(self.y):
            return
        nx, nw, nb = self.need
        st = stream_ptr()
        y, dy = self.y, gp(self.y)
        bias_done = False
        if self.act != _C.ACT_NONE and not self.act_bwd_fused:
            if nb and self.b is not None and y.dtype == y.grad_dtype:
                call('rcgan_act_bwd_colsum', dy, dp(y), y.rows, y.c, y.ld, y.ld, y.dtype, self.act, self.leak, gp(self.b), self.acc_b, st)
                bias_done = True
            else:
                call('rcgan_act_bwd', dy, dp(y), dy, y.rows, y.c, y.ld, y.ld, y.ld, y.dtype, self.act, self.leak, 0, st)
        if nb and self.b is not None and not bias_done:
            prog.fork(lambda: call('rcgan_colsum', dy, y.rows, y.c, y.ld, y.grad_dtype, gp(self.b), self.acc_b, stream_ptr()))
            bias_done = True
        d, dyin = self.desc, dy
        if self.patch is not None and (nx or nw):
            call('rcgan_im2col', self.desc, dy, dp(self.patch), self.patch.ld, st)
            d, dyin = self.gdesc, dp(self.patch)
        wgrad = lambda: call('rcgan_conv2d_wgrad', d, dyin, dp(self.x), gp(self.w), self.acc_w, prog.ws.ptr(), prog.ws.bytes, stream_ptr())
        if nw and nx and FORK_WGRAD and torch.cuda.is_available():
            prog.fork(wgrad)
            nw = False
        if nx:
            call('rcgan_conv2d_fprop', d, dyin, dp(self.w), pp(self.pack), None, gp(self.x), self.x.grad_dtype, _C.ACT_NONE,
                 0.0, st)
        if nw:
            wgrad()
        if nb and self.b is not None and not bias_done:
            call('rcgan_colsum', dy, y.rows, y.c, y.ld, y.grad_dtype, gp(self.b), self.acc_b, st)


class BatchNormOp(Op):
    """y = act(BN(x)).  labels None: tf.contrib.layers.batch_norm (mnist/ops.py:30-44), scale/offset [c];
    labels int32 [n]: cond_batchnorm (cifar10/common/ops/normalization.py:27-59), tables [n_labels,c]."""

    def __init__(self, x, scale, offset, labels=None, moving=None, train=True, eps=1e-5, decay=0.9, act=None, leak=0.2, groups=1,
                 concat_y=None, stats_from=None):
        """concat_y (fp32 [n, c2]): the output is concat([act(BN(x)), y broadcast over H, W]) -- the generator's
        conv_cond_concat / concat right after each norm (mnist/model.py:714-728) -- written in the same pass.
        groups > 1: the batch is `groups` equal sample ranges with INDEPENDENT batch statistics (the reference's separate
        discriminator calls on the real and on the generated batch, mnist/model.py:150-207, run here as one pass)."""
        prog = cur()
        n, h, w = spatial(x)
        self.x, self.scale, self.offset, self.labels = x, scale, offset, labels
        assert n % groups == 0
        self.groups, self.samples, self.hw, self.c = groups, n // groups, h * w, x.c
        assert x.ld == x.c
        # stats_from: the conv that produces x accumulates the batch statistics in its epilogue (emit_colstats): no statistics pass
        self.stats_buf = None
        if stats_from is not None and train and groups == 1 and concat_y is None:
            self.stats_buf = stats_from.emit_colstats()
        self.n_labels = scale.numel() // x.c
        self.moving, self.train, self.eps, self.decay = moving, train, eps, decay
        self.act, self.leak = ACT[act], leak
        assert x.grad_dtype == prog.act_dtype or x.dtype == x.grad_dtype
        self.cat = concat_y
        if concat_y is None:
            self.y = prog.new(x.shape, x.grad_dtype)
            self.c2 = 0
        else:
            assert groups == 1 and concat_y.dtype == _C.F32 and concat_y.shape[0] == n and concat_y.ld == concat_y.c
            self.c2 = concat_y.c
            self.y = prog.new(tuple(x.shape[:-1]) + (x.c + self.c2,), x.grad_dtype, ld=round_up(x.c + self.c2, 8))
        self.save = [torch.zeros(2 * x.c, dtype=torch.float32, device=prog.device) for _ in range(groups)]
        self.inputs, self.outputs = (x, scale, offset), (self.y,)
        self.ws_bytes = _C.load().rcgan_bn_workspace(self.samples, self.hw, self.c)
        prog.ws.request(self.ws_bytes)
        self.dummy = None
        prog.add(self)

    def plan_bwd(self, prog):
        nx, ns, no = self.need
        self.acc_x = self.claim(self.x) if nx else 0
        if ns or no:
            self.claim(self.scale), self.claim(self.offset)
        elif needs(self.y):
            self.dummy = torch.zeros(2 * self.n_labels * self.c, dtype=torch.float32, device=prog.device)
        self.dx_dummy = None
        if self.cat is not None and needs(self.y) and not nx:
            # parameter gradients only: with a concat output dx cannot alias y.grad (different row stride)
            self.dx_dummy = torch.zeros(self.x.data.numel(), dtype=self.y.data.dtype, device=prog.device)

    def _mask_offset(self):
        """rcgan_bn_bwd can re-derive the relu mask from x (offset given) instead of reading y back: 5 instead of 7 tensor passes,
        but measured SLOWER on B200 (tools/bn_bench.py, 512x32x32x256: 444 vs 412 us -- these kernels are issue-bound, not
        bandwidth-bound, and the extra FMA costs more than the extra load): opt-in RCGAN_BN_MASK_FROM_X=1"""
        return dp(self.offset) if BN_MASK_FROM_X else None

    def _lab(self, g):
        """labels of sample range g (int32 [n])"""
        return None if self.labels is None else dp(self.labels) + 4 * g * self.samples

    def _off(self, ptr, g, esz):
        """pointer to sample range g of a [groups*samples*hw, c] buffer"""
        return None if ptr is None else ptr + g * self.samples * self.hw * self.c * esz

    def forward(self, prog):
        mm = mv = None
        if self.moving is not None:
            mm, mv = dp(self.moving[0]), dp(self.moving[1])
        if self.cat is not None:
            call('rcgan_bn_fwd_cat', dp(self.x), dp(self.y), self.y.ld, dp(self.cat), self.c2, self.samples, self.hw, self.c,
                 self.x.dtype, self.y.dtype, dp(self.scale), dp(self.offset), dp(self.labels), self.eps, self.act, self.leak,
                 1 if self.train else 0, self.decay, mm, mv, self.save[0].data_ptr(), prog.ws.ptr(), prog.ws.bytes, stream_ptr())
            return
        if self.stats_buf is not None:
            call('rcgan_bn_fwd_prestats', dp(self.x), dp(self.y), self.samples, self.hw, self.c, self.x.dtype, self.y.dtype, dp(self.scale),
                 dp(self.offset), dp(self.labels), self.eps, self.act, self.leak, self.decay, mm, mv, self.save[0].data_ptr(),
                 pp(self.stats_buf), stream_ptr())
            return
        xs, ys = self.x.data.element_size(), self.y.data.element_size()
        for g in range(self.groups):
            call('rcgan_bn_fwd', self._off(dp(self.x), g, xs), self._off(dp(self.y), g, ys), self.samples, self.hw, self.c,
                 self.x.dtype, self.y.dtype, dp(self.scale), dp(self.offset), self._lab(g), self.eps, self.act, self.leak,
                 1 if self.train else 0, self.decay, mm, mv, self.save[g].data_ptr(), prog.ws.ptr(), prog.ws.bytes, stream_ptr())

    def backward(self, prog):
        if not needs(self.y):
            return
        nx, ns, no = self.need
        if not self.train:
            # gen_sampler's norms (moving statistics are constants): only dL/dx exists -- the label-recovery path
            assert not (ns or no), 'inference-mode batch norm has no parameter gradients on any reference path'
            if nx:
                ys, gs = self.y.data.element_size(), self.y.grad.element_size()
                for g in range(self.groups):
                    call('rcgan_bn_infer_bwd', self._off(gp(self.y), g, gs), self._off(dp(self.y), g, ys), self.y.ld,
                         self._off(gp(self.x), g, self.x.grad.element_size()), self.samples, self.hw, self.c, self.y.dtype,
                         dp(self.scale), self._lab(g), self.save[g].data_ptr(), self.act, self.leak, self.acc_x, prog.ws.ptr(),
                         prog.ws.bytes, stream_ptr())
            return
        if self.dummy is not None:
            dsc, dof = self.dummy.data_ptr(), self.dummy.data_ptr() + 4 * self.n_labels * self.c
            accp = 0
        else:
            dsc, dof, accp = gp(self.scale), gp(self.offset), 1
        if self.cat is not None:
            dxp, accx = (gp(self.x), self.acc_x) if nx else (self.dx_dummy.data_ptr(), 0)
            call('rcgan_bn_bwd_cat', gp(self.y), dp(self.x), dp(self.y), self.y.ld, dxp, self.samples, self.hw, self.c, self.x.dtype,
                 self.y.dtype, dp(self.scale), dp(self.labels), self.n_labels, self.save[0].data_ptr(), self.act, self.leak, dsc, dof,
                 accx, accp, prog.ws.ptr(), prog.ws.bytes, self._mask_offset(), stream_ptr())
            return
        xs, ys, gs = self.x.data.element_size(), self.y.data.element_size(), self.y.grad.element_size()
        for g in range(self.groups):
            if not nx:
                # only parameter gradients wanted: dx still has to go somewhere -> reuse y.grad in place
                dxp, accx = self._off(gp(self.y), g, gs), 0
            else:
                dxp, accx = self._off(gp(self.x), g, self.x.grad.element_size()), self.acc_x
            call('rcgan_bn_bwd', self._off(gp(self.y), g, gs), self._off(dp(self.x), g, xs), self._off(dp(self.y), g, ys), dxp,
                 self.samples, self.hw, self.c, self.x.dtype, self.y.dtype, dp(self.scale), self._lab(g), self.n_labels,
                 self.save[g].data_ptr(), self.act, self.leak, dsc, dof, accx, accp if g == 0 else 1, prog.ws.ptr(),
                 prog.ws.bytes, self._mask_offset(), stream_ptr())


class ActOp(Op):
    """Standalone activation (lrelu mnist/ops.py:94-95; relu/tanh/sigmoid)."""

    def __init__(self, x, act, leak=0.2):
        prog = cur()
        self.x, self.act, self.leak = x, ACT[act], leak
        self.y = prog.new(x.shape, x.dtype)
        self.inputs, self.outputs = (x,), (self.y,)
        self.bwd_fused = False           # the consuming conv's dgrad writes act'(y) * gradient straight into x's gradient
        derive(x, self.y)
        prog.add(self)

    def plan_bwd(self, prog):
        self.acc_x = self.claim(self.x) if (self.need[0] and not self.bwd_fused) else 0

    def forward(self, prog):
        x, y = self.x, self.y
        call('rcgan_bias_act_fwd', dp(x), None, dp(y), x.rows, x.c, x.ld, y.ld, x.dtype, self.act, self.leak, stream_ptr())

    def backward(self, prog):
        if not needs(self.y) or not self.need[0] or self.bwd_fused:
            return
        x, y = self.x, self.y
        call('rcgan_act_bwd', gp(y), dp(y), gp(x), y.rows, y.c, y.ld, y.ld, x.ld, y.dtype, self.act, self.leak, self.acc_x,
             stream_ptr())


class ConcatLabelOp(Op):
    """concat([a, y broadcast over H,W], channel axis): conv_cond_concat (mnist/ops.py:46-51) and the
    z|y / h|y concats of the generator (mnist/model.py:714-728).  Output channel stride is padded to a
    multiple of 8 (zero filled) so 16-byte vector / TMA loads stay aligned."""

    def __init__(self, a, y):
        prog = cur()
        n, h, w = spatial(a)
        self.a, self.yb = a, y
        assert y.dtype == _C.F32 and y.shape[0] == n
        c = a.c + y.c
        oshape = (n, c) if len(a.shape) == 2 else (n, h, w, c)
        self.out = prog.new(oshape, a.dtype, ld=round_up(c, 8))
        self.rps = h * w
        self.inputs, self.outputs = (a, y), (self.out,)
        prog.add(self)

    def plan(self, prog):
        self.out.base.needs_grad = needs(self.a)   # labels are data, never differentiated

    def plan_bwd(self, prog):
        self.acc_a = self.claim(self.a) if self.need[0] else 0

    def forward(self, prog):
        a, o = self.a, self.out
        call('rcgan_concat_label_fwd', dp(a), a.ld, dp(self.yb), dp(o), o.ld, o.rows, self.rps, a.c, self.yb.c, a.dtype,
             stream_ptr())

    def backward(self, prog):
        if not needs(self.out) or not self.need[0]:
            return
        a, o = self.a, self.out
        call('rcgan_slice_bwd', gp(o), o.ld, gp(a), a.ld, o.rows, a.c, a.dtype, self.acc_a, stream_ptr())


class MeanHWOp(Op):
    """tf.reduce_mean(x, axis=(1,2)), optionally of relu(x) (mnist/model.py:678; gan_resnet.py:405-407)."""

    def __init__(self, x, relu=False):
        prog = cur()
        n, h, w = spatial(x)
        assert x.ld == x.c
        self.x, self.relu, self.hw = x, int(relu), h * w
        self.y = prog.new((n, x.c), x.dtype)
        self.inputs, self.outputs = (x,), (self.y,)
        prog.add(self)

    def plan_bwd(self, prog):
        self.acc_x = self.claim(self.x) if self.need[0] else 0

    def forward(self, prog):
        call('rcgan_meanhw_fwd', dp(self.x), dp(self.y), self.x.shape[0], self.hw, self.x.c, self.x.dtype, self.relu,
             stream_ptr())

    def backward(self, prog):
        if not needs(self.y) or not self.need[0]:
            return
        call('rcgan_meanhw_bwd', gp(self.y), dp(self.x), gp(self.x), self.x.shape[0], self.hw, self.x.c, self.x.dtype,
             self.relu, self.acc_x, stream_ptr())


class CastOp(Op):
    """dtype boundary between the bf16 trunk and the fp32 head / inputs."""

    def __init__(self, x, dtype):
        prog = cur()
        assert x.ld == x.c
        self.x = x
        self.y = prog.new(x.shape, dtype)
        self.inputs, self.outputs = (x,), (self.y,)
        prog.add(self)

    def plan_bwd(self, prog):
        self.acc_x = self.claim(self.x) if self.need[0] else 0
        if self.acc_x:
            from .graph import TORCH_DTYPE
            self.tmp = torch.zeros(self.x.data.numel(), dtype=TORCH_DTYPE[self.x.grad_dtype], device=prog.device)

    def forward(self, prog):
        call('rcgan_cast', dp(self.x), self.x.dtype, dp(self.y), self.y.dtype, self.x.numel(), stream_ptr())

    def backward(self, prog):
        if not needs(self.y) or not self.need[0]:
            return
        st = stream_ptr()
        if self.acc_x:
            call('rcgan_cast', gp(self.y), self.y.grad_dtype, self.tmp.data_ptr(), self.x.grad_dtype, self.x.numel(), st)
            call('rcgan_copy_acc', self.tmp.data_ptr(), gp(self.x), self.x.numel(), self.x.grad_dtype, 1, st)
        else:
            call('rcgan_cast', gp(self.y), self.y.grad_dtype, gp(self.x), self.x.grad_dtype, self.x.numel(), st)


class SpectralNormOp(Op):
    """W_bar = W / sigma(W, u) with one power iteration and the gradient THROUGH it
    (spectral_normed_weight, mnist/sn.py:17-75).  u_new is assigned to u after the step
    unless update_collection == NO_OPS."""

    def __init__(self, W, u, update=True):
        prog = cur()
        self.W, self.u = W, u
        self.c = W.shape[-1]
        self.m = W.numel() // self.c
        lib = _C.load()
        self.wbar = prog.new(W.shape, _C.F32)
        self.u_new = prog.new((1, self.c), _C.F32)
        self.save = torch.zeros(lib.rcgan_sn_save_floats(self.m, self.c), dtype=torch.float32, device=prog.device)
        prog.ws.request(lib.rcgan_sn_workspace(self.m, self.c))
        self.inputs, self.outputs = (W,), (self.wbar,)
        prog.add(self)
        self.updates_u = bool(update)
        if update:
            prog.add_update(u, self.u_new)
        # parameters only: a W computed by another op of the program is not available at program start
        self.batched = bool(W.is_variable and u.is_variable)
        self.weight_only = self.batched
        self.wbar.static_weight = self.batched
        self.updates_state = bool(update)          # u moves every run: never skippable as "unchanged since the last run"

    def plan_bwd(self, prog):
        self.acc_w = self.claim(self.W) if self.need[0] else 0

    # The ops of one program form a group: the FIRST one (program order) launches the batched forward for all of them
    # -- W and u are parameters, available at program start -- and, running last in the reverse sweep, the batched
    # backward once every dL/dW_bar is complete.  The others are bookkeeping only (rcgan_sn_*_batched in the header).
    def _group(self, prog):
        grp = prog.__dict__.get('sn_group')
        if grp is None:
            import ctypes
            ops = [o for o in prog.ops if isinstance(o, SpectralNormOp) and o.batched]
            n = len(ops)
            PA, IA = ctypes.c_void_p * n, ctypes.c_int * n
            grp = prog.sn_group = {
                'ops': ops, 'leader': ops[0], 'n': n,
                'W': PA(*[dp(o.W) for o in ops]), 'u': PA(*[dp(o.u) for o in ops]),
                'wbar': PA(*[dp(o.wbar) for o in ops]), 'unew': PA(*[dp(o.u_new) for o in ops]),
                'save': PA(*[o.save.data_ptr() for o in ops]),
                'm': IA(*[o.m for o in ops]), 'c': IA(*[o.c for o in ops]),
            }
            grp['ws'] = torch.zeros(_C.load().rcgan_sn_workspace_batched(n, grp['m'], grp['c']), dtype=torch.uint8,
                                    device=prog.device)
            bw = [o for o in ops if needs(o.wbar) and o.need[0]]
            if bw:
                nb = len(bw)
                PB, IB = ctypes.c_void_p * nb, ctypes.c_int * nb
                grp['bwd'] = {
                    'n': nb, 'W': PB(*[dp(o.W) for o in bw]), 'u': PB(*[dp(o.u) for o in bw]),
                    'G': PB(*[gp(o.wbar) for o in bw]), 'save': PB(*[o.save.data_ptr() for o in bw]),
                    'dW': PB(*[gp(o.W) for o in bw]), 'm': IB(*[o.m for o in bw]), 'c': IB(*[o.c for o in bw]),
                    'acc': IB(*[int(o.acc_w) for o in bw]),
                }
        return grp

    def forward(self, prog):
        if not self.batched:
            call('rcgan_sn_fwd', dp(self.W), dp(self.u), self.m, self.c, dp(self.wbar), dp(self.u_new), self.save.data_ptr(),
                 prog.ws.ptr(), prog.ws.bytes, stream_ptr())
            return
        g = self._group(prog)
        if g['leader'] is not self:
            return
        call('rcgan_sn_fwd_batched', g['n'], g['W'], g['u'], g['m'], g['c'], g['wbar'], g['unew'], g['save'],
             g['ws'].data_ptr(), g['ws'].numel(), stream_ptr())

    def backward(self, prog):
        if not self.batched:
            if needs(self.wbar) and self.need[0]:
                call('rcgan_sn_bwd', dp(self.W), dp(self.u), gp(self.wbar), self.m, self.c, self.save.data_ptr(), gp(self.W),
                     self.acc_w, prog.ws.ptr(), prog.ws.bytes, stream_ptr())
            return
        g = self._group(prog)
        if g['leader'] is not self or 'bwd' not in g:
            return
        b = g['bwd']
        call('rcgan_sn_bwd_batched', b['n'], b['W'], b['u'], b['G'], b['m'], b['c'], b['save'], b['dW'], b['acc'],
             g['ws'].data_ptr(), g['ws'].numel(), stream_ptr())


class ChannelLossOp(Op):
    """L = coef * mean_b sum_j wgt[b,j] * phi(psi[b] + <h[b], V[j]>)   (SURVEY appendix B).
    The value stored in the program's loss slot is the un-weighted mean (coef applies to gradients)."""

    def __init__(self, h, psi, V, wgt, mode, name, coef=1.0, rows=None):
        """rows=(row0, B): the term covers rows [row0, row0+B) of h / psi only -- D run once on [real; fake]
        (cifar10/gan_resnet.py:563-590 slices disc_all the same way) with one loss term per half."""
        prog = cur()
        self.h, self.psi, self.V, self.wgt = h, psi, V, wgt
        self.row0, self.B = (0, h.shape[0]) if rows is None else (int(rows[0]), int(rows[1]))
        self.sliced = rows is not None
        assert 0 <= self.row0 and self.row0 + self.B <= h.shape[0]
        self.d, self.k = h.c, V.shape[0]
        assert h.ld == h.c and V.c == h.c and wgt.shape == (self.B, self.k)
        assert psi is None or psi.numel() == h.shape[0]
        assert not self.sliced or (h.dtype == _C.F32 and (psi is None or psi.dtype == _C.F32))
        self.mode, self.coef = mode, coef
        self.slot = prog.loss_slot(name)
        self.logits = prog.new((self.B, self.k), _C.F32)
        self.inputs, self.outputs = (h, psi, V, wgt), ()
        prog.add(self)

    def plan(self, prog):
        pass

    def plan_bwd(self, prog):
        nh, np_, nv, nw = self.need
        self.acc_h = self.claim(self.h) if nh else 0
        self.zero_h = self.zero_psi = False
        if self.sliced and nh and self.acc_h == 0:
            self.zero_h, self.acc_h = True, 1        # first writer of a row-sliced gradient: clear it all, then everyone +=
        if np_:
            acc_psi = self.claim(self.psi)
            if self.sliced:
                self.zero_psi = acc_psi == 0
            else:
                assert acc_psi == 0, 'psi gradient must have a single writer'
        self.zero_v = False
        if nv:
            self.zero_v = (self.claim(self.V) == 0) and not self.V.is_variable
        if nw:
            assert self.claim(self.wgt) == 0, 'weight-matrix gradient must have a single writer'

    def _off(self, ptr, per_row):
        return None if ptr is None else ptr + 4 * self.row0 * per_row

    def forward(self, prog):
        call('rcgan_channel_loss', self._off(dp(self.h), self.d) if self.sliced else dp(self.h),
             self._off(dp(self.psi), 1) if self.sliced else dp(self.psi), dp(self.V), dp(self.wgt), self.B, self.d, self.k,
             self.h.dtype, self.mode, 1.0 / self.B, prog.losses.data_ptr() + 4 * self.slot, dp(self.logits), None, 0, None, None,
             None, stream_ptr())

    def backward(self, prog):
        nh, np_, nv, nw = self.need
        if not (nh or np_ or nv or nw):
            return
        st = stream_ptr()
        if nv and self.zero_v:
            call('rcgan_zero', gp(self.V), self.V.grad.numel() * 4, st)
        if self.zero_h:
            call('rcgan_zero', gp(self.h), self.h.grad.numel() * self.h.grad.element_size(), st)
        if self.zero_psi:
            call('rcgan_zero', gp(self.psi), self.psi.grad.numel() * self.psi.grad.element_size(), st)
        gh, gpsi = (gp(self.h) if nh else None), (gp(self.psi) if np_ else None)
        if self.sliced:
            gh, gpsi = self._off(gh, self.d), self._off(gpsi, 1)
        call('rcgan_channel_loss', self._off(dp(self.h), self.d) if self.sliced else dp(self.h),
             self._off(dp(self.psi), 1) if self.sliced else dp(self.psi), dp(self.V), dp(self.wgt), self.B, self.d, self.k,
             self.h.dtype, self.mode, self.coef / self.B, None, None, gh, self.acc_h, gpsi, gp(self.V) if nv else None,
             gp(self.wgt) if nw else None, st)


class RecoverMSEOp(Op):
    """mse_loss of DCGAN.recover_labels (mnist/model.py:538-541): mean_r sum_j y_rec[r,j] * mean_p (actual[r,p] - sample[r*k+j,p])^2."""

    def __init__(self, sample, actual, y_rec, name):
        prog = cur()
        assert sample.dtype == _C.F32 and actual.dtype == _C.F32 and y_rec.dtype == _C.F32 and sample.ld == sample.c
        self.sample, self.actual, self.y_rec = sample, actual, y_rec
        self.R, self.k = y_rec.shape
        self.npix = sample.numel() // (self.R * self.k)
        assert actual.numel() == self.R * self.npix
        self.slot = prog.loss_slot(name)
        self.sq = prog.new((self.R, self.k), _C.F32)
        self.inputs, self.outputs = (sample, y_rec), ()
        prog.add(self)

    def plan(self, prog):
        pass

    def plan_bwd(self, prog):
        for i, t in enumerate((self.sample, self.y_rec)):
            if self.need[i]:
                assert self.claim(t) == 0, 'recover-loss gradients must have a single writer'

    def forward(self, prog):
        call('rcgan_recover_mse', dp(self.sample), dp(self.actual), dp(self.y_rec), self.R, self.k, self.npix,
             prog.losses.data_ptr() + 4 * self.slot, dp(self.sq), None, None, stream_ptr())

    def backward(self, prog):
        ns, ny = self.need
        if not (ns or ny):
            return
        call('rcgan_recover_mse', dp(self.sample), dp(self.actual), dp(self.y_rec), self.R, self.k, self.npix, None, None,
             gp(self.sample) if ns else None, gp(self.y_rec) if ny else None, stream_ptr())


class SigmoidCEOp(Op):
    """coef * mean(sigmoid_cross_entropy_with_logits(logits, targets)) over all B*k elements
    (perm regulariser: mnist/model.py:214-224; cifar10/gan_resnet.py:687-695, 780-784)."""

    def __init__(self, logits, targets, name, coef=1.0):
        prog = cur()
        assert logits.dtype == _C.F32 and targets.dtype == _C.F32 and logits.ld == logits.c
        self.logits, self.targets, self.coef = logits, targets, coef
        self.slot = prog.loss_slot(name)
        self.inputs, self.outputs = (logits,), ()
        prog.add(self)

    def plan(self, prog):
        pass

    def plan_bwd(self, prog):
        if self.need[0]:
            assert self.claim(self.logits) == 0, 'classifier logits gradient must have a single writer'

    def forward(self, prog):
        n = self.logits.numel()
        call('rcgan_sigmoid_ce', dp(self.logits), dp(self.targets), n, 1.0 / n, prog.losses.data_ptr() + 4 * self.slot, None,
             stream_ptr())

    def backward(self, prog):
        if not self.need[0]:
            return
        n = self.logits.numel()
        call('rcgan_sigmoid_ce', dp(self.logits), dp(self.targets), n, self.coef / n, None, gp(self.logits), stream_ptr())


class LogitLossOp(Op):
    """coef * mean_b phi(l[b]) on a scalar logit per sample (vanilla discriminator, mnist/model.py:687-703)."""

    def __init__(self, logits, mode, name, coef=1.0):
        prog = cur()
        assert logits.dtype == _C.F32 and logits.ld == logits.c
        self.logits, self.mode, self.coef = logits, mode, coef
        self.slot = prog.loss_slot(name)
        self.inputs, self.outputs = (logits,), ()
        prog.add(self)

    def plan(self, prog):
        pass

    def plan_bwd(self, prog):
        if self.need[0]:
            assert self.claim(self.logits) == 0, 'logit gradient must have a single writer'

    def forward(self, prog):
        n = self.logits.numel()
        call('rcgan_logit_loss', dp(self.logits), n, self.mode, 1.0 / n, prog.losses.data_ptr() + 4 * self.slot, None,
             stream_ptr())

    def backward(self, prog):
        if not self.need[0]:
            return
        n = self.logits.numel()
        call('rcgan_logit_loss', dp(self.logits), n, self.mode, self.coef / n, None, gp(self.logits), stream_ptr())


class SoftmaxRowsOp(Op):
    """confusion_matrix = softmax(confusion_logits, -1) (mnist/model.py:106; gan_resnet.py:522)."""

    def __init__(self, logits):
        prog = cur()
        self.x = logits
        self.rows, self.k = logits.shape
        self.y = prog.new(logits.shape, _C.F32)
        self.inputs, self.outputs = (logits,), (self.y,)
        prog.add(self)

    def plan_bwd(self, prog):
        self.acc = self.claim(self.x) if self.need[0] else 0

    def forward(self, prog):
        call('rcgan_softmax_rows_fwd', dp(self.x), dp(self.y), self.rows, self.k, stream_ptr())

    def backward(self, prog):
        if not needs(self.y) or not self.need[0]:
            return
        call('rcgan_softmax_rows_bwd', dp(self.y), gp(self.y), gp(self.x), self.rows, self.k, self.acc, stream_ptr())


class GatherRowsOp(Op):
    """wgt[b,:] = C[y[b],:]  == tensordot(one_hot(y), C) (cifar10/gan_resnet.py:682-683, 757)."""

    def __init__(self, C, y):
        prog = cur()
        self.C, self.yidx = C, y
        self.B, self.k = y.numel(), C.shape[1]
        self.out = prog.new((self.B, self.k), _C.F32)
        self.inputs, self.outputs = (C,), (self.out,)
        prog.add(self)

    def plan_bwd(self, prog):
        self.acc = self.claim(self.C) if self.need[0] else 0

    def forward(self, prog):
        call('rcgan_gather_rows_fwd', dp(self.C), dp(self.yidx), dp(self.out), self.B, self.k, stream_ptr())

    def backward(self, prog):
        if not needs(self.out) or not self.need[0]:
            return
        call('rcgan_gather_rows_bwd', gp(self.out), dp(self.yidx), gp(self.C), self.B, self.k, self.C.shape[0], self.acc,
             stream_ptr())


class Pool2Op(Op):
    """2x2 mean pool as add_n of four strided slices / 4 (cifar10/gan_resnet.py:239-240, 248-249)."""

    def __init__(self, x):
        prog = cur()
        n, h, w = spatial(x)
        assert x.ld == x.c
        self.x = x
        self.y = prog.new((n, h // 2, w // 2, x.c), x.dtype)
        self.inputs, self.outputs = (x,), (self.y,)
        prog.add(self)

    def plan_bwd(self, prog):
        self.acc = self.claim(self.x) if self.need[0] else 0

    def forward(self, prog):
        n, h, w = spatial(self.x)
        call('rcgan_avgpool2_fwd', dp(self.x), dp(self.y), n, h, w, self.x.c, self.x.dtype, stream_ptr())

    def backward(self, prog):
        if not needs(self.y) or not self.need[0]:
            return
        n, h, w = spatial(self.x)
        call('rcgan_avgpool2_bwd', gp(self.y), gp(self.x), n, h, w, self.x.c, self.x.dtype, self.acc, stream_ptr())


class Upsample2Op(Op):
    """concat x4 + depth_to_space == nearest-neighbour 2x upsample (cifar10/gan_resnet.py:263-264)."""

    def __init__(self, x):
        prog = cur()
        n, h, w = spatial(x)
        assert x.ld == x.c
        self.x = x
        self.y = prog.new((n, 2 * h, 2 * w, x.c), x.dtype)
        self.inputs, self.outputs = (x,), (self.y,)
        self.consumer = None      # set by the ConvOp of an UpsampleConv pair that can fold this op away
        prog.add(self)

    def plan_bwd(self, prog):
        self.acc = self.claim(self.x) if self.need[0] else 0

    def forward(self, prog):
        if self.consumer is not None and self.consumer.folds_upsample():
            return                # the only reader runs on self.x directly (rcgan_upconv2d_fprop)
        n, h, w = spatial(self.x)
        call('rcgan_upsample2_fwd', dp(self.x), dp(self.y), n, h, w, self.x.c, self.x.dtype, stream_ptr())

    def backward(self, prog):
        if not needs(self.y) or not self.need[0]:
            return
        n, h, w = spatial(self.x)
        call('rcgan_upsample2_bwd', gp(self.y), gp(self.x), n, h, w, self.x.c, self.x.dtype, self.acc, stream_ptr())


class AddOp(Op):
    """shortcut + output (cifar10/gan_resnet.py:328, 353)."""

    def __init__(self, a, b):
        prog = cur()
        assert a.shape == b.shape and a.ld == a.c and b.ld == b.c
        self.a, self.b = a, b
        self.y = prog.new(a.shape, a.dtype)
        self.inputs, self.outputs = (a, b), (self.y,)
        prog.add(self)

    def plan_bwd(self, prog):
        # d(a + b)/db = identity: when this op is b's ONLY reader, b's gradient IS y's gradient -- alias the buffer instead of
        # copying it (b's producer runs its backward after this op and only reads it; y's gradient is complete by then)
        a, b, y = self.a.base, self.b.base, self.y.base
        ok = lambda t, me: bool(needs(self.y) and t is me and y is self.y and not t.is_variable and not t.grad_written
                                and t._grad is not None and y._grad is not None and t._grad.numel() == y._grad.numel()
                                and t._grad.dtype == y._grad.dtype)
        self.alias_b = bool(self.need[1] and ok(b, self.b) and getattr(b, 'n_readers', 0) == 1)
        # ... and a's gradient STARTS as y's gradient: the same buffer can be a's too, accumulated into in place by a's other
        # readers (the residual trunk keeps ONE gradient buffer through the blocks) -- provided those all run their backward after
        # everything that still reads the buffer as dL/dy, i.e. they precede b's producer in program order, and that producer does
        # not modify its output gradient in place
        pb = getattr(b, 'producer', None)
        others = [o for o in grad_writer_ops(a) if o is not self]
        self.alias_a = bool(ALIAS_RESIDUAL and self.need[0] and ok(a, self.a) and a is not b
                            and (not self.alias_b or (pb is not None and getattr(pb, 'act', None) == _C.ACT_NONE
                                                      and all(o.index < pb.index for o in others))))
        if self.alias_a:
            a._grad = y._grad
            a.grad_written = True
            self.acc_a = 0
        else:
            self.acc_a = self.claim(self.a) if self.need[0] else 0
        if self.alias_b:
            b._grad = y._grad
            b.grad_written = True
            self.acc_b = 0
        else:
            self.acc_b = self.claim(self.b) if self.need[1] else 0

    def forward(self, prog):
        call('rcgan_add', dp(self.a), dp(self.b), dp(self.y), self.a.numel(), self.a.dtype, stream_ptr())

    def backward(self, prog):
        if not needs(self.y):
            return
        st = stream_ptr()
        if self.need[0] and not self.alias_a:
            call('rcgan_copy_acc', gp(self.y), gp(self.a), self.a.numel(), self.a.dtype, self.acc_a, st)
        if self.need[1] and not self.alias_b:
            call('rcgan_copy_acc', gp(self.y), gp(self.b), self.b.numel(), self.b.dtype, self.acc_b, st)


class ConcatRowsOp(Op):
    """tf.concat([a, b], axis=0) of two dense tensors (cifar10/gan_resnet.py:563-578: D on [real; fake])."""

    def __init__(self, a, b):
        prog = cur()
        assert a.shape[1:] == b.shape[1:] and a.ld == a.c and b.ld == b.c and a.dtype == b.dtype
        self.a, self.b = a, b
        self.y = prog.new((a.shape[0] + b.shape[0],) + tuple(a.shape[1:]), a.dtype)
        self.inputs, self.outputs = (a, b), (self.y,)
        prog.add(self)

    def plan_bwd(self, prog):
        self.acc_a = self.claim(self.a) if self.need[0] else 0
        self.acc_b = self.claim(self.b) if self.need[1] else 0

    def _esz(self):
        return 4 if self.a.dtype == _C.F32 else 2

    def forward(self, prog):
        st = stream_ptr()
        call('rcgan_copy_acc', dp(self.a), dp(self.y), self.a.numel(), self.a.dtype, 0, st)
        call('rcgan_copy_acc', dp(self.b), dp(self.y) + self.a.numel() * self._esz(), self.b.numel(), self.a.dtype, 0, st)

    def backward(self, prog):
        if not needs(self.y):
            return
        st = stream_ptr()
        if self.need[0]:
            call('rcgan_copy_acc', gp(self.y), gp(self.a), self.a.numel(), self.a.dtype, self.acc_a, st)
        if self.need[1]:
            call('rcgan_copy_acc', gp(self.y) + self.a.numel() * self._esz(), gp(self.b), self.b.numel(), self.a.dtype,
                 self.acc_b, st)


class FoldWeightOp(Op):
    """3x3 filter -> the 4x4 stride-2 filter that absorbs a 2x resampling next to the conv (rcgan_wfold4):
    mode 'pool': ConvMeanPool (cifar10/gan_resnet.py:231-241) = one stride-2 conv, 4/9 of the flops, no full-resolution output;
    mode 'up':   UpsampleConv (:259-272) = one stride-2 conv2d_transpose on the small input, no upsampled tensor.
    Both folds are linear; backward is the adjoint on the filter gradient."""

    def __init__(self, w, mode):
        prog = cur()
        kh, kw, cin, cout = w.shape
        assert kh == 3 and kw == 3
        self.w, self.mode, self.cin, self.cout = w, {'pool': 0, 'up': 1}[mode], cin, cout
        self.w4 = prog.new((4, 4, cin, cout) if self.mode == 0 else (4, 4, cout, cin), _C.F32)
        self.inputs, self.outputs = (w,), (self.w4,)
        self.weight_only = is_static_weight(w)
        self.w4.static_weight = self.weight_only
        prog.add(self)

    def plan_bwd(self, prog):
        self.acc_w = self.claim(self.w) if self.need[0] else 0

    def forward(self, prog):
        call('rcgan_wfold4', dp(self.w), dp(self.w4), self.cin, self.cout, self.mode, stream_ptr())

    def backward(self, prog):
        if not needs(self.w4) or not self.need[0]:
            return
        call('rcgan_wfold4_bwd', gp(self.w4), gp(self.w), self.cin, self.cout, self.mode, self.acc_w, stream_ptr())


class PreprocessCifarOp(Op):
    """int32 CHW [n,3072] -> 2*(v/256 - .5) (+ dequantisation noise) -> NHWC (cifar10/gan_resnet.py:548-552)."""

    def __init__(self, raw, noise, dtype):
        prog = cur()
        self.raw, self.noise = raw, noise
        self.n = raw.shape[0]
        self.y = prog.new((self.n, 32, 32, 3), dtype)
        self.inputs, self.outputs = (), (self.y,)
        prog.add(self)

    def plan(self, prog):
        self.y.base.needs_grad = False

    def forward(self, prog):
        fn = 'rcgan_preprocess_cifar_u8' if self.raw.data.dtype == torch.uint8 else 'rcgan_preprocess_cifar'
        call(fn, self.raw.data.data_ptr(), None if self.noise is None else self.noise.data.data_ptr(), dp(self.y), self.n,
             self.y.dtype, stream_ptr())


class RandomFillOp(Op):
    """An in-graph random input: tf.random_normal([n, 128]) (gan_resnet.py:363-364) / tf.random_uniform([n, 3072], 0, 1/128)
    (:550).  `step` is a device int64 the host bumps before every step, so captured graphs draw fresh numbers on replay."""
    _streams = 0

    def __init__(self, shape, normal, a, b, seed, step_dev):
        prog = cur()
        self.y = prog.new(shape, _C.F32)
        self.normal, self.a, self.b, self.seed, self.step_dev = int(normal), float(a), float(b), int(seed), step_dev
        RandomFillOp._streams += 1
        self.stream_id = RandomFillOp._streams
        self.inputs, self.outputs = (), (self.y,)
        prog.add(self)

    def plan(self, prog):
        self.y.base.needs_grad = False

    def forward(self, prog):
        call('rcgan_random_fill', dp(self.y), self.y.numel(), self.normal, self.a, self.b, self.seed, self.step_dev.data_ptr(),
             self.stream_id, stream_ptr())


def adam_step(group, lr_t_dev, b1, b2, eps, grad_scale=1.0):
    """One tf.train.AdamOptimizer update of a whole variable group (one launch)."""
    import ctypes
    clip = [(v.offset, v.offset + v.numel()) for v in group.vars if v.clip]
    assert len(clip) <= 8
    lo = (ctypes.c_long * 8)(*([c[0] for c in clip] + [0] * (8 - len(clip))))
    hi = (ctypes.c_long * 8)(*([c[1] for c in clip] + [0] * (8 - len(clip))))
    call('rcgan_adam_tf', group.params.data_ptr(), group.grads.data_ptr(), group.m.data_ptr(), group.v.data_ptr(),
         group.numel, 0.0, lr_t_dev.data_ptr(), b1, b2, eps, grad_scale, lo, hi, len(clip), stream_ptr())
