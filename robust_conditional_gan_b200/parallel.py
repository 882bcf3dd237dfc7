"""Data parallelism of the hot path: towers of the reference (gan_resnet.py:183-192, 529-546, 697, 786)
are ranks here.  Parameters, spectral-norm u vectors, Adam state and confusion_logits are replicated; every feed is
split contiguously across ranks exactly like tf.split(x, len(DEVICES)); each rank's loss is the mean over ITS shard
and the tower costs are averaged, so the gradient is the mean over ranks: NCCL all-reduce(sum) of the flat gradient
arenas per optimizer step, with the 1/world_size folded into the Adam kernel (grad_scale).

The exchange is overlapped with the backward sweep (north_star; SURVEY 8e "bucketed from G.Output backwards"): the arena
of the variables a step trains is cut into contiguous buckets, and a bucket's all-reduce is issued the moment the last
kernel that writes one of its gradients has been enqueued -- torch.distributed runs it on NCCL's own (high-priority) stream
behind an event, so it proceeds while the remaining dgrad / wgrad kernels run; the step waits for all buckets right before
Adam.  Everything (collectives included) is part of the step's single CUDA graph: no host round trip between backward,
exchange and update."""
import os

import torch

BUCKET_BYTES = int(os.environ.get('RCGAN_DP_BUCKET_MB', '8')) << 20


def shard(t, rank, world):
    """tf.split(t, world, axis=0)[rank]"""
    n = t.shape[0]
    assert n % world == 0, 'batch %d is not divisible by %d towers' % (n, world)
    k = n // world
    return t[rank * k:(rank + 1) * k]


def allreduce_sum_(flat_grads, world):
    """in-place sum over ranks of one flat gradient arena (blocking form; GradReducer is the overlapped one)"""
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM)
    return flat_grads


def last_writer_index(prog, var):
    """index of the op whose backward enqueues the LAST kernel that writes var's gradient (the reverse sweep runs ops in
    decreasing index, so that is the smallest index among the ops that differentiate with respect to var); len(ops) when no op
    of the program does (the zero-filled gradient is final before the sweep starts)."""
    from .nnops import SpectralNormOp
    idx = len(prog.ops)
    leader = None
    for op in prog.ops:
        if isinstance(op, SpectralNormOp) and op.batched:
            leader = op if leader is None else leader
        for i, t in enumerate(op.inputs):
            if t is not None and t.is_variable and t.base is var.base and i < len(op.need) and op.need[i]:
                j = op.index
                if isinstance(op, SpectralNormOp) and op.batched:
                    j = leader.index          # every batched spectral-norm backward is one launch at the group's first op
                idx = min(idx, j)
    return idx


def plan_buckets(prog, store, group_names, bucket_bytes=None):
    """Cut the gradient span of `group_names` (adjacent slices of store.all_grads) into contiguous buckets of about bucket_bytes,
    walking from the END of the span (gradients of the last-created variables are final first).  Returns a list of
    (lo, hi, trigger) in elements of store.all_grads, trigger = op index after whose backward the bucket is complete."""
    bucket_bytes = BUCKET_BYTES if bucket_bytes is None else bucket_bytes
    groups = sorted((store.groups[k] for k in group_names if k in store.groups), key=lambda g: g.grad_offset)
    if not groups:
        return []
    spans = []                                   # (lo, hi, trigger) per variable, in arena order
    for g in groups:
        for v in g.vars:
            lo = g.grad_offset + v.offset
            spans.append([lo, lo + ((v.numel() + 63) // 64) * 64, last_writer_index(prog, v)])
    for a, b in zip(groups, groups[1:]):
        assert a.grad_offset + a.numel == b.grad_offset, 'groups of one step must be adjacent in the gradient buffer'
    buckets, cur = [], None
    for lo, hi, trig in reversed(spans):
        if cur is None:
            cur = [lo, hi, trig]
        else:
            cur[0], cur[2] = lo, min(cur[2], trig)
        if (cur[1] - cur[0]) * 4 >= bucket_bytes:
            buckets.append(tuple(cur))
            cur = None
    if cur is not None:
        buckets.append(tuple(cur))
    # a bucket whose trigger is not later than its predecessor's can only be launched after it anyway: keep them in order
    return buckets


class GradReducer:
    """Per (program, trained groups): the bucket plan and the hooks that launch each bucket's all-reduce from inside
    Program.run_backward; wait() joins them in front of the optimizer."""

    def __init__(self, prog, store, group_names, world, process_group=None, bucket_bytes=None):
        self.world, self.pg, self.store = world, process_group, store
        self.buckets = plan_buckets(prog, store, group_names, bucket_bytes)
        self.pending = []
        for lo, hi, trig in self.buckets:
            prog.after_backward.setdefault(trig, []).append(lambda lo=lo, hi=hi: self._launch(lo, hi))

    def _launch(self, lo, hi):
        if self.world <= 1:
            return
        import torch.distributed as dist
        self.pending.append(dist.all_reduce(self.store.all_grads[lo:hi], op=dist.ReduceOp.SUM, group=self.pg, async_op=True))

    def wait(self):
        for w in self.pending:
            w.wait()                    # stream dependency only: the current stream waits for the collective's completion event
        self.pending = []


def shutdown(models=(), grace_s=20.0):
    """End of a multi-rank run: the step graphs hold captured NCCL collectives, and tearing the communicator down while they are
    alive can block forever (seen on 2 x B200: the job printed its result and then sat in destroy_process_group until killed).
    Order: flush output, barrier, release the captured graphs, then destroy the process group with a watchdog that ends the
    process if the teardown does not return within grace_s."""
    import gc
    import sys
    import threading
    import torch.distributed as dist
    sys.stdout.flush(); sys.stderr.flush()
    if not (dist.is_available() and dist.is_initialized()):
        return
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    for m in models:
        getattr(m, '_graphs', {}).clear()
        for r in getattr(m, 'reducers', {}).values():
            r.pending = []
    gc.collect()
    torch.cuda.synchronize()
    t = threading.Timer(grace_s, lambda: os._exit(0))
    t.daemon = True
    t.start()
    dist.destroy_process_group()
    t.cancel()
