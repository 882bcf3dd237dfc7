"""Data parallelism of the hot path: towers of the reference (gan_resnet.py:183-192, 529-546, 697, 786)
are ranks here.  Parameters, spectral-norm u vectors, Adam state and confusion_logits are replicated; every feed is
split contiguously across ranks exactly like tf.split(x, len(DEVICES)); each rank's loss is the mean over ITS shard
and the tower costs are averaged, so the gradient is the mean over ranks: one NCCL all-reduce(sum) of each flat
gradient arena per optimizer step, with the 1/world_size folded into the Adam kernel (grad_scale)."""
import torch


def shard(t, rank, world):
    """tf.split(t, world, axis=0)[rank]"""
    n = t.shape[0]
    assert n % world == 0, 'batch %d is not divisible by %d towers' % (n, world)
    k = n // world
    return t[rank * k:(rank + 1) * k]


def allreduce_sum_(flat_grads, world):
    """in-place sum over ranks of one flat gradient arena (6.7 MB D / 31.5 MB G for the CIFAR nets: latency-bound on
    NVLink 5, so one message per arena rather than per-variable buckets)"""
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM)
    return flat_grads
