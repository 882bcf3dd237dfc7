"""N>1 host logic on CPU: world_size-2 gloo.  Towers == ranks: each rank differentiates the mean loss over ITS shard,
the flat gradient arenas are all-reduced (sum) and scaled by 1/world -- this must equal the gradient of the reference's
two-tower cost add_n(costs)/len(DEVICES) (cifar10/gan_resnet.py:697, 786), including the PER-TOWER conditional-BN
statistics of the generator."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import cifar as OC
from robust_conditional_gan_b200.parallel import allreduce_sum_, shard


def _tower_batch(b, r, w):
    out = {}
    for k, v in b.items():
        out[k] = shard(v, r, w) if torch.is_tensor(v) and v.ndim >= 1 else v
    return out


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)
    cfg = OC.default_config(algorithm='rcgan-u', alpha=0.5, perm_classifier=True, dim=16, confuse_init=True)
    P = OC.init_params(cfg, 1, torch.float64)
    full = OC.synthetic_batch(8, 3, torch.float64)
    mine = _tower_batch(full, rank, world)
    tr = OC.Trainer(P, cfg)
    tr._req(tr.gn + ['confusion_logits'])
    out, _ = OC.gen_cost(tr.P, mine, cfg)
    names = tr.gn + ['confusion_logits']
    gs = torch.autograd.grad(out['gen_cost'], [tr.P[n] for n in names])
    flat = torch.cat([g.reshape(-1) for g in gs])          # the flat gradient arena of this rank
    allreduce_sum_(flat, world)
    flat *= 1.0 / world                                     # what rcgan_adam_tf's grad_scale applies
    if rank == 0:
        q.put(flat.clone())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_equals_two_tower_reference_cost():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    flat = q.get(timeout=300)
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    # single process, two towers, exactly as the reference builds it
    cfg = OC.default_config(algorithm='rcgan-u', alpha=0.5, perm_classifier=True, dim=16, confuse_init=True)
    P = OC.init_params(cfg, 1, torch.float64)
    full = OC.synthetic_batch(8, 3, torch.float64)
    tr = OC.Trainer(P, cfg)
    names = tr.gn + ['confusion_logits']
    tr._req(names)
    cost = sum(OC.gen_cost(tr.P, _tower_batch(full, r, world), cfg)[0]['gen_cost'] for r in range(world)) / world
    gs = torch.autograd.grad(cost, [tr.P[n] for n in names])
    ref = torch.cat([g.reshape(-1) for g in gs])
    assert float((flat - ref).abs().max()) < 1e-12
    # and it is NOT the gradient of one big tower (the per-tower BN statistics matter)
    tr._req(names)
    big = OC.gen_cost(tr.P, full, cfg)[0]['gen_cost']
    gb = torch.cat([g.reshape(-1) for g in torch.autograd.grad(big, [tr.P[n] for n in names])])
    assert float((gb - ref).norm() / ref.norm()) > 1e-3


def test_shard_is_tf_split():
    t = torch.arange(24).reshape(8, 3)
    assert torch.equal(torch.cat([shard(t, r, 4) for r in range(4)]), t)
    with pytest.raises(AssertionError):
        shard(t, 0, 3)
