"""Static-program runtime under the reference's graph-building API.

The reference builds a TensorFlow graph once (`DCGAN.build_model`, the module-level graph of
gan_resnet.py) and then runs `sess.run(train_op)` per step.  This module is the B200-native
equivalent of that split: the op wrappers (ops.py, sn.py, cifar/...) record *ops* into a
`Program`; `Program.finalize()` plans the backward sweep (which tensors need gradients, which
gradient writes overwrite and which accumulate) and allocates every buffer once; a step is then
a fixed sequence of launches into `librcgan_b200.so` that is captured into a CUDA graph.

PyTorch's role here is device memory (`torch.empty`), streams and CUDA-graph capture only: no
autograd, no aten compute kernels on the step path.
"""
import math

import numpy as np
import torch

from . import _C
from ._C import ConvDesc, call, ptr, stream_ptr

SIDE_STREAM = __import__('os').environ.get('RCGAN_SIDE_STREAM', '1') == '1'   # A/B switch of Program.fork
# derived (spectral-normed / folded) conv weights: their gradient buffers are zeroed by ONE kernel at the start of the backward
# sweep and every wgrad accumulates, instead of one cudaMemset node in front of each wgrad (a memset node is a full dependency:
# the wgrad behind it loses its programmatic-dependent-launch overlap)
PREZERO_WGRAD = __import__('os').environ.get('RCGAN_PREZERO_WGRAD', '1') == '1'
I32 = 100   # host-side dtype tag for integer label tensors (never passed to the library as an activation dtype)
U8 = 101    # raw image bytes (CIFAR pixels)
TORCH_DTYPE = {_C.F32: torch.float32, _C.BF16: torch.bfloat16, I32: torch.int32, U8: torch.uint8}


def same_pad(size, k, stride):
    """TF SAME: (out, pad_before)."""
    out = -(-size // stride)
    total = max((out - 1) * stride + k - size, 0)
    return out, total // 2


def round_up(x, m):
    return (x + m - 1) // m * m


class Tensor:
    """A node of the static graph: logical channels-last shape, channel stride `ld`,
    device buffer `data`, gradient buffer `grad` (allocated by Program.finalize when needed)."""

    def __init__(self, shape, dtype=_C.F32, ld=None, name=None, device=None, data=None, grad_dtype=None):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = dtype
        self.grad_dtype = dtype if grad_dtype is None else grad_dtype   # fp32 pre-norm tensors carry bf16 gradients
        self.c = self.shape[-1]
        self.ld = int(ld) if ld is not None else self.c
        self.rows = int(np.prod(self.shape[:-1])) if len(self.shape) > 1 else 1
        self.name = name
        self.base = self           # views share the base's buffers and planning state
        self._data = data if data is not None else torch.zeros(self.rows * self.ld, dtype=TORCH_DTYPE[dtype], device=device)
        self._grad = None
        self.needs_grad = False    # planning state (per program for activations; variables: see Variable)
        self.grad_written = False
        self.is_variable = False

    # buffers live on the base so that views alias them
    @property
    def data(self):
        return self.base._data

    @property
    def grad(self):
        return self.base._grad

    def view(self, shape):
        """Reinterpret a dense tensor (ld == c) with another channels-last shape."""
        assert self.ld == self.c, 'cannot view a padded tensor'
        assert int(np.prod(shape)) == self.rows * self.c, (shape, self.shape)
        v = Tensor.__new__(Tensor)
        v.shape = tuple(int(s) for s in shape)
        v.dtype, v.c, v.ld = self.dtype, int(shape[-1]), int(shape[-1])
        v.grad_dtype = self.grad_dtype
        v.rows = int(np.prod(shape[:-1]))
        v.name, v.base = self.name, self.base
        v._data = v._grad = None
        v.is_variable = self.is_variable
        return v

    def numel(self):
        return self.rows * self.c

    def torch(self):
        """Dense logical view (strips channel padding) as a torch tensor -- for tests/feeding only."""
        t = self.data.view(self.rows, self.ld)[:, :self.c]
        return t.reshape(self.shape)

    def grad_torch(self):
        t = self.grad.view(self.rows, self.ld)[:, :self.c]
        return t.reshape(self.shape)


def needs(t):
    return t is not None and t.base.needs_grad


def is_static_weight(t):
    """known at program start: a parameter, or computed from parameters only by weight_only ops"""
    return t.is_variable or getattr(t.base, 'static_weight', False)


class Variable(Tensor):
    """A named parameter / state tensor (fp32).  After VariableStore.finalize() `data`, `grad`,
    `m`, `v` are views into the flat per-group arenas."""

    def __init__(self, name, shape, init, trainable, group, device):
        super().__init__(shape, _C.F32, name=name, device=device, data=init.to(device=device, dtype=torch.float32).reshape(-1).contiguous())
        self.trainable = trainable
        self.group = group
        self.is_variable = True
        self.clip = False          # max-norm variable constraint (mnist/ops.py:101-111)
        self.offset = None


class Group:
    """One optimizer's variables packed into flat fp32 arenas: params, grads, Adam m/v."""

    def __init__(self, name):
        self.name = name
        self.vars = []
        self.params = self.grads = self.m = self.v = None
        self.numel = 0
        self.t = 0


class VariableStore:
    """tf.get_variable with scope reuse + the reference's variable naming."""

    def __init__(self, device, group_fn):
        self.device = device
        self.vars = {}
        self.group_fn = group_fn       # name -> group name or None
        self.groups = {}
        self.finalized = False
        self.scope = []
        self.on_load = []              # callbacks after load_state_dict (cached derived state, e.g. weight packs, is stale)

    def full_name(self, name):
        return '/'.join(self.scope + [name])

    def get(self, name, shape, init_fn, trainable=True):
        full = self.full_name(name)
        if full in self.vars:
            v = self.vars[full]
            assert tuple(v.shape) == tuple(shape), 'variable %s reused with another shape' % full
            return v
        assert not self.finalized, 'cannot create %s after finalize()' % full
        init = init_fn(shape)
        v = Variable(full, shape, torch.as_tensor(init), trainable, self.group_fn(full) if trainable else None, self.device)
        self.vars[full] = v
        return v

    def finalize(self):
        """Pack trainable variables into per-group arenas (256-byte aligned offsets)."""
        for v in self.vars.values():
            if v.group is not None:
                self.groups.setdefault(v.group, Group(v.group)).vars.append(v)
        for g in self.groups.values():
            off = 0
            for v in g.vars:
                v.offset = off
                off += round_up(v.numel(), 64)
            g.numel = off
            g.params = torch.zeros(off, dtype=torch.float32, device=self.device)
            g.m = torch.zeros(off, dtype=torch.float32, device=self.device)
            g.v = torch.zeros(off, dtype=torch.float32, device=self.device)
        # the gradient arenas of all groups are slices of ONE buffer: groups trained by the same step (generator + confusion
        # matrix) are adjacent, so one collective covers them
        self.all_grads = torch.zeros(sum(g.numel for g in self.groups.values()), dtype=torch.float32, device=self.device)
        pos = 0
        for g in sorted(self.groups.values(), key=lambda g_: {'g': 0, 'c': 1}.get(g_.name, 2)):
            g.grad_offset = pos
            g.grads = self.all_grads[pos:pos + g.numel]
            pos += g.numel
        for g in self.groups.values():
            for v in g.vars:
                seg = g.params[v.offset:v.offset + v.numel()]
                seg.copy_(v._data)
                v._data = seg
                v._grad = g.grads[v.offset:v.offset + v.numel()]
        self.finalized = True

    def state_dict(self):
        return {n: v.data.detach().reshape(v.shape).clone() for n, v in self.vars.items()}

    def load_state_dict(self, sd):
        for n, t in sd.items():
            self.vars[n].data.copy_(torch.as_tensor(t).detach().to(self.vars[n].data.device, torch.float32).reshape(-1))
        for cb in self.on_load:
            cb()


class Workspace:
    """One scratch buffer shared by every op of a program (ops run serially on one stream)."""

    def __init__(self):
        self.bytes = 0
        self.buf = None

    def request(self, nbytes):
        self.bytes = max(self.bytes, int(nbytes))

    def allocate(self, device):
        self.buf = torch.empty(max(self.bytes, 256), dtype=torch.uint8, device=device)

    def ptr(self):
        return self.buf.data_ptr()


class Op:
    """An op records its inputs/outputs at build time; plan() runs in forward order and decides
    which input gradients it must produce; plan_bwd() runs in reverse order and decides, per
    gradient buffer, whether this op is the first writer (overwrite) or a later one (accumulate)."""
    inputs = ()
    outputs = ()
    foreign = False
    # an op whose inputs are parameters only (spectral norm, filter folds): Program.run_forward runs these first, followed by ONE
    # batched refresh of every tensor-core weight pack, and run_backward runs their backward last
    weight_only = False

    def plan(self, prog):
        ng = any(needs(t) for t in self.inputs)
        for o in self.outputs:
            o.base.needs_grad = ng

    def claim(self, t):
        """Returns the accumulate flag for writing t's gradient in the backward sweep."""
        if t.is_variable:
            return 1            # parameter gradients: arena zeroed at step start, always +=
        acc = 1 if t.base.grad_written else 0
        t.base.grad_written = True
        return acc

    def claim_wgrad(self, prog, t):
        """claim() for a conv filter gradient: the first writer of a derived weight's gradient registers the buffer for the
        program's batched zeroing and accumulates (see PREZERO_WGRAD)"""
        if PREZERO_WGRAD and not t.is_variable and not t.base.grad_written:
            prog.prezero.append(t.base)
            t.base.grad_written = True
            return 1
        return self.claim(t)

    def plan_bwd(self, prog):
        pass

    def forward(self, prog):
        raise NotImplementedError

    def backward(self, prog):
        pass


class Program:
    """A fixed launch sequence: forward ops in order, then backward ops in reverse."""
    current = None

    def __init__(self, name, device, act_dtype):
        self.name = name
        self.device = device
        self.act_dtype = act_dtype
        self.ops = []
        self.ws = Workspace()
        self.tensors = []
        self.inputs = {}
        self.loss_names = []
        self.losses = None          # fp32 [n_losses] device
        self.updates = []           # (dst tensor, src tensor) state assignments applied after backward
        self.packs = {}             # id(weight base) -> bf16 tensor-core weight pack (torch uint8 buffer)
        self.pack_jobs = []         # (desc, weight tensor, pack buffer) refreshed in one launch at the start of every run
        self._pack_args = None
        self._pack_args_own = None
        self._update_args = None
        self.prezero = []           # tensors whose gradient buffer is zeroed at the start of the backward sweep (claim_wgrad)
        self._prezero_args = None
        self.after_backward = {}    # op index -> [callable]; index len(ops) = before the sweep starts
        self._side, self._forked = None, False
        self.finalized = False

    def __enter__(self):
        self.prev = Program.current
        Program.current = self
        return self

    def __exit__(self, *a):
        Program.current = self.prev

    # ---- build-time helpers
    def new(self, shape, dtype=None, ld=None, name=None, grad_dtype=None):
        t = Tensor(shape, self.act_dtype if dtype is None else dtype, ld=ld, name=name, device=self.device,
                   grad_dtype=grad_dtype)
        self.tensors.append(t)
        return t

    def input(self, name, shape, dtype=_C.F32):
        t = self.new(shape, dtype, name=name)
        self.inputs[name] = t
        return t

    def add(self, op):
        op.index = len(self.ops)          # program order: backward planning reasons about who runs before whom
        for o in op.outputs:
            o.base.producer = op
        self.ops.append(op)
        return op

    def weight_pack(self, w, desc):
        """bf16 tensor-core pack of a conv weight, shared by every op of this program that uses the weight.
        Returns (buffer or None, owner): the owner (first user in program order) refreshes the pack each run."""
        nbytes = _C.load().rcgan_conv_wpack_bytes(desc)
        if nbytes == 0:
            return None, False
        key = id(w.base)
        if key in self.packs:
            assert self.packs[key].numel() == nbytes
            return self.packs[key], False
        self.packs[key] = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        return self.packs[key], True

    def hoist_pack(self, w, desc, buf):
        """Register the pack of a weight that is known at program start (a parameter or the output of a weight_only op) for the
        batched refresh (rcgan_conv_wpack_batched); False if the weight is computed mid-program and its user must pack it."""
        if not is_static_weight(w):
            return False
        self.pack_jobs.append((desc, w, buf))
        return True

    def _refresh_packs(self, refresh_foreign=True):
        if not self.pack_jobs:
            return
        if self._pack_args is None:
            import ctypes

            def args(jobs):
                n = len(jobs)
                PA = ctypes.c_void_p * max(n, 1)
                return (n, PA(*[ctypes.addressof(d) for d, _, _ in jobs]), PA(*[w.data.data_ptr() for _, w, _ in jobs]),
                        PA(*[b.data_ptr() for _, _, b in jobs]))
            self._pack_args = args(self.pack_jobs)
            # (variables are shared between programs, so their foreign-ness is decided against THIS program's wrt set)
            foreign = lambda w: (id(w.base) not in self.wrt) if w.is_variable else getattr(w.base, 'foreign', False)
            self._pack_args_own = args([j for j in self.pack_jobs if not foreign(j[1])])
        n, descs, ws, packs = self._pack_args if refresh_foreign else self._pack_args_own
        if n:
            call('rcgan_conv_wpack_batched', n, descs, ws, packs, stream_ptr())

    def fork(self, fn):
        """Run fn() -- launches that only READ what the current stream has produced so far and whose results nobody needs before
        the optimizer -- on a side stream, concurrently with what the caller enqueues next (a bias gradient's column sums next to
        the same layer's dgrad / wgrad: bandwidth-bound work in the shadow of tensor-bound work).  join() before anything may
        overwrite what fn reads.  Captured into the step's CUDA graph as a fork / join."""
        if not SIDE_STREAM:
            fn()
            return
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        cur_s = torch.cuda.current_stream(self.device)
        self._side.wait_stream(cur_s)
        with torch.cuda.stream(self._side):
            fn()
        self._forked = True

    def join(self):
        if self._forked:
            torch.cuda.current_stream(self.device).wait_stream(self._side)
            self._forked = False

    def loss_slot(self, name):
        self.loss_names.append(name)
        return len(self.loss_names) - 1

    def add_update(self, dst, src):
        self.updates.append((dst, src))

    # ---- planning
    def finalize(self, wrt):
        """wrt: iterable of Variables this program differentiates with respect to."""
        wrt = set(id(v) for v in wrt)
        self.wrt = wrt
        self.prezero, self._prezero_args = [], None
        for t in self.tensors:
            t.needs_grad = False
            t.grad_written = False
        seen_vars = {}
        for op in self.ops:
            for t in op.inputs:
                if t is not None and t.is_variable:
                    t.base.needs_grad = id(t.base) in wrt
                    seen_vars[id(t.base)] = t.base
            op.plan(self)
            op.need = [needs(t) for t in op.inputs]     # frozen: variables are shared across programs
        # how many ops read each activation (lets a sole reader alias gradient buffers instead of copying, see AddOp)
        for t in self.tensors:
            t.n_readers = 0
            t.readers = []
        for op in self.ops:
            for t in op.inputs:
                if t is not None and not t.is_variable:
                    t.base.n_readers = getattr(t.base, 'n_readers', 0) + 1
                    if not hasattr(t.base, 'readers'):
                        t.base.readers = []
                    t.base.readers.append(op)
        for t in self.tensors:
            if t.needs_grad and t._grad is None:
                t._grad = torch.zeros(t._data.numel(), dtype=TORCH_DTYPE[t.grad_dtype], device=self.device)
        for op in reversed(self.ops):
            op.plan_bwd(self)
        # weight-only ops whose parameters are all outside `wrt` ("foreign": another optimizer's variables)
        for op in self.ops:
            op.foreign = False
            if op.weight_only:
                src = set()
                for t in op.inputs:
                    if t is None:
                        continue
                    src |= {id(t.base)} if t.is_variable else getattr(t.base, 'src_vars', set())
                for u in getattr(op, 'state_inputs', ()):
                    src.add(id(u.base))
                op.foreign = bool(src) and not (src & wrt) and not getattr(op, 'updates_state', False)
                for o in op.outputs:
                    o.base.src_vars = src
                    o.base.foreign = op.foreign
        self.ws.allocate(self.device)
        self.losses = torch.zeros(max(len(self.loss_names), 1), dtype=torch.float32, device=self.device)
        self.finalized = True

    # ---- execution (enqueue only; never synchronises)
    def run_forward(self, refresh_foreign=True):
        """refresh_foreign=False: the folds / weight packs that depend only on variables this program does NOT train are taken
        as still valid from the previous run (the caller knows those variables have not changed since: the generator's weights
        over the discriminator steps of one iteration)"""
        call('rcgan_zero', self.losses.data_ptr(), self.losses.numel() * 4, stream_ptr())
        for op in self.ops:
            if op.weight_only and (refresh_foreign or not op.foreign):
                op.forward(self)
        self._refresh_packs(refresh_foreign)
        for op in self.ops:
            if not op.weight_only:
                op.forward(self)

    def run_backward(self):
        """reverse program order (the weight-only ops too: a fold's backward directly follows its conv's, the batched spectral-norm
        backward sits at the first normalised weight, i.e. after everything that feeds it).  after_backward[i]: callables run once
        op i's backward is enqueued -- the data-parallel reducer launches a gradient bucket there (parallel.GradReducer)."""
        hooks = self.after_backward
        if self.prezero:
            if self._prezero_args is None:
                import ctypes
                n = len(self.prezero)
                numels = [t._grad.numel() for t in self.prezero]
                assert all(t._grad.dtype == torch.float32 for t in self.prezero)
                self._zero_src = torch.zeros(max(numels), dtype=torch.float32, device=self.device)
                PA, LA = ctypes.c_void_p * n, ctypes.c_long * n
                self._prezero_args = (n, PA(*[self._zero_src.data_ptr()] * n), PA(*[t._grad.data_ptr() for t in self.prezero]), LA(*numels))
            call('rcgan_copy_batched', *self._prezero_args, stream_ptr())
        for h in hooks.get(len(self.ops), ()):
            h()
        for op in reversed(self.ops):
            op.backward(self)
            self.join()             # (ops that fork join themselves; this is the safety net before the hooks / the next op)
            for h in hooks.get(op.index, ()):
                h()

    def run_updates(self):
        if not self.updates:
            return
        if self._update_args is None:
            import ctypes
            n = len(self.updates)
            PA, LA = ctypes.c_void_p * n, ctypes.c_long * n
            self._update_args = (n, PA(*[s.data.data_ptr() for _, s in self.updates]), PA(*[d.data.data_ptr() for d, _ in self.updates]),
                                 LA(*[d.numel() for d, _ in self.updates]))
        n, srcs, dsts, numels = self._update_args
        call('rcgan_copy_batched', n, srcs, dsts, numels, stream_ptr())

    def loss_dict(self, host):
        return {n: float(host[i]) for i, n in enumerate(self.loss_names)}


def cur():
    assert Program.current is not None, 'no active Program (use `with program:`)'
    return Program.current


def tf_adam_lr(lr, b1, b2, t):
    """tf.train.AdamOptimizer's effective step size at step t (1-based)."""
    return lr * math.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
