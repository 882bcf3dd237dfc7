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


# ---------------------------------------------------------------------------------------------- overlapped bucket reducer
def _small_program():
    """a 3-conv 'generator' + a 'confusion' softmax on the CPU (planned, never launched): variables in two optimizer groups"""
    from robust_conditional_gan_b200 import _C, scope as S
    from robust_conditional_gan_b200.graph import Program, VariableStore
    from robust_conditional_gan_b200.nnops import ConvOp, MeanHWOp, SoftmaxRowsOp
    dev = torch.device('cpu')
    st = VariableStore(dev, lambda n: 'c' if n == 'confusion_logits' else ('g' if n.startswith('g') else 'd'))
    S.set_store(st, 0)
    mk = lambda n, s: st.get(n, s, lambda sh: torch.randn(sh) * 0.1)
    v = {n: mk(n, s) for n, s in [('g1', (3, 3, 64, 64)), ('g2', (3, 3, 64, 64)), ('g3', (3, 3, 64, 64)), ('d1', (3, 3, 64, 64)),
                                  ('confusion_logits', (10, 10))]}
    p = Program('g_step', dev, _C.BF16)
    with p:
        x = p.input('x', [2, 8, 8, 64], _C.BF16)
        a = ConvOp(x, v['g1'], None)
        b = ConvOp(a.y, v['g2'], None)
        c = ConvOp(b.y, v['g3'], None)
        d = ConvOp(c.y, v['d1'], None)          # the other optimizer's variable: differentiated through, not trained
        MeanHWOp(d.y)
        sm = SoftmaxRowsOp(v['confusion_logits'])
    st.finalize()
    p.finalize([v['g1'], v['g2'], v['g3'], v['confusion_logits']])
    return st, p, v, (a, b, c, d, sm)


def test_bucket_plan_covers_the_trained_span_in_backward_order():
    from robust_conditional_gan_b200.parallel import last_writer_index, plan_buckets
    st, p, v, (a, b, c, d, sm) = _small_program()
    g, cg = st.groups['g'], st.groups['c']
    assert g.grad_offset + g.numel == cg.grad_offset                      # generator and confusion arenas are adjacent
    assert last_writer_index(p, v['g3']) == c.index and last_writer_index(p, v['g1']) == a.index
    assert last_writer_index(p, v['d1']) == len(p.ops)                    # not trained here: nothing writes its gradient
    per_var = 9 * 64 * 64 * 4
    bk = plan_buckets(p, st, ('g', 'c'), bucket_bytes=per_var)            # ~one conv weight per bucket
    assert bk[0][1] == cg.grad_offset + cg.numel and bk[-1][0] == g.grad_offset
    for (lo, hi, t), (lo2, hi2, t2) in zip(bk, bk[1:]):
        assert lo == hi2 and lo2 < lo                                     # contiguous, walking towards the arena start
    # the first bucket (end of the span: confusion_logits + g3) completes when g3's wgrad is enqueued, the last one with g1's
    assert bk[0][2] == min(c.index, sm.index) and bk[-1][2] == a.index
    assert len(plan_buckets(p, st, ('g', 'c'), bucket_bytes=1 << 30)) == 1


def _reducer_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from robust_conditional_gan_b200.parallel import GradReducer
    st, p, v, ops_ = _small_program()
    red = GradReducer(p, st, ('g', 'c'), world, bucket_bytes=9 * 64 * 64 * 4)
    gen = torch.Generator().manual_seed(100 + rank)
    # emulate the reverse sweep: an op's gradient becomes final, then its hooks (bucket launches) run
    launched = []
    for h in p.after_backward.get(len(p.ops), ()):
        h()
    for op in reversed(p.ops):
        for t in op.inputs:
            if t is not None and t.is_variable and t.grad is not None and id(t.base) in p.wrt:
                t.grad.copy_(torch.randn(t.grad.shape, generator=gen))
        for h in p.after_backward.get(op.index, ()):
            h()
            launched.append(op.index)
    red.wait()
    lo, hi = st.groups['g'].grad_offset, st.groups['c'].grad_offset + st.groups['c'].numel
    if rank == 0:
        q.put((st.all_grads[lo:hi].clone(), st.groups['d'].grads.clone(), launched))
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_reducer_sums_over_ranks_from_inside_the_backward_sweep():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_reducer_worker, args=(r, world, port, q)) for r in range(world)]
    for p_ in procs:
        p_.start()
    got, dgrads, launched = q.get(timeout=300)
    for p_ in procs:
        p_.join(timeout=300)
        assert p_.exitcode == 0
    st, p, v, _ = _small_program()
    ref = torch.zeros_like(got)
    lo = st.groups['g'].grad_offset
    for rank in range(world):
        gen = torch.Generator().manual_seed(100 + rank)
        for op in reversed(p.ops):
            for t in op.inputs:
                if t is not None and t.is_variable and t.grad is not None and id(t.base) in p.wrt:
                    r = torch.randn(t.grad.shape, generator=gen)
                    g = st.groups[t.base.group]
                    o = g.grad_offset + t.base.offset - lo
                    ref[o:o + r.numel()] += r
    assert torch.equal(got, ref)
    assert float(dgrads.abs().max()) == 0.0                       # the other optimizer's arena is not touched
    assert launched == sorted(launched, reverse=True) and len(launched) >= 3
