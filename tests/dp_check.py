"""Multi-GPU check (run under torchrun on >= 2 GPUs; not a pytest file):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dp_check.py
R ranks x tower batch n (fp32) vs the oracle's R-tower cost: identical D-step and G-step gradients after the NCCL
all-reduce, identical parameters on every rank after Adam."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cifar as OC
from robust_conditional_gan_b200.cifar.gan_resnet import RCGANCifar, default_flags
from robust_conditional_gan_b200.parallel import shard


def mark(rank, what):
    """progress marker on stderr (a hang is located by the last marker of each rank)"""
    if os.environ.get('DP_CHECK_TRACE'):
        sys.stderr.write('[dp_check rank %d] %s\n' % (rank, what))
        sys.stderr.flush()


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    n, dim, alg = 4, 32, 'rcgan-u'
    flags = default_flags(algorithm=alg, alpha=0.5, perm_classifier=True, perm_multiplier=2.0, confuse_init=True)
    ocfg = OC.default_config(algorithm=alg, alpha=0.5, perm_classifier=True, perm_multiplier=2.0, confuse_init=True, dim=dim)
    model = RCGANCifar(flags, tower_batch=n, precision='fp32', dim=dim, use_cuda_graph=True, world_size=world, rank=rank)
    P = OC.init_params(ocfg, seed=1, dtype=torch.float64)
    model.store.load_state_dict(P)
    full = OC.synthetic_batch(n * world, seed=3, dtype=torch.float64)
    tower = lambda r: {k: (shard(v, r, world) if torch.is_tensor(v) else v) for k, v in full.items()}
    mine = tower(rank)
    model.feed(model.d_prog, all_real_data_int=mine['raw'], all_real_labels=mine['labels'], all_random_labels=mine['labels_random'],
               all_labels_biased=mine['labels_biased'], all_labels_inv_weights=mine['inv_weights'], noise=mine['noise'],
               dequant_noise=torch.zeros(n, 3072))
    model.feed(model.g_prog, noise=mine['noise_G'], all_random_labels_G=mine['labels_random_G'],
               all_labels_biased_G=mine['labels_biased_G'])
    mark(rank, 'fed')
    model.d_step(0)
    torch.cuda.synchronize()
    mark(rank, 'd_step done')
    # reference: R towers in one process
    tr = OC.Trainer(P, ocfg)
    tr._req(tr.dn)
    cost = sum(OC.disc_cost(tr.P, tower(r), ocfg)[0]['disc_cost'] for r in range(world)) / world
    gs = dict(zip(tr.dn, torch.autograd.grad(cost, [tr.P[k] for k in tr.dn], allow_unused=True)))
    worst = 0.0
    for v in model.disc_params:
        ref = gs[v.name]
        if ref is None or float(ref.norm()) < 1e-10:
            continue
        got = v.grad.double().cpu().reshape(ref.shape) / world          # arena holds the SUM; Adam applies 1/world
        worst = max(worst, float((got - ref).norm() / ref.norm()))
    # parameters must be bit-identical across ranks after the update
    flat = model.groups['d'].params.clone()
    other = flat.clone()
    dist.broadcast(other, src=0)
    same = bool(torch.equal(flat, other))
    mark(rank, 'broadcast done')
    # G step from the same initial state: the bucketed all-reduce (generator arena + confusion_logits, launched from inside the
    # backward sweep) against the R-tower generator cost
    model.store.load_state_dict(P)
    model.g_step(1)
    torch.cuda.synchronize()
    mark(rank, 'g_step done')
    tr = OC.Trainer(OC.init_params(ocfg, seed=1, dtype=torch.float64), ocfg)
    names = tr.gn + ['confusion_logits']
    tr._req(names)
    cost = sum(OC.gen_cost(tr.P, tower(r), ocfg)[0]['gen_cost'] for r in range(world)) / world
    gs = dict(zip(names, torch.autograd.grad(cost, [tr.P[k] for k in names], allow_unused=True)))
    worst_g = 0.0
    for v in model.gen_params + model.c_params:
        ref = gs[v.name]
        if ref is None or float(ref.norm()) < 1e-10:
            continue
        got = v.grad.double().cpu().reshape(ref.shape) / world
        worst_g = max(worst_g, float((got - ref).norm() / ref.norm()))
    flatg = model.groups['g'].params.clone(); og = flatg.clone(); dist.broadcast(og, src=0)
    nb = {k: len(r.buckets) for k, r in model.reducers.items()}
    res = torch.tensor([worst, 0.0 if same else 1.0, 0.0 if torch.equal(flatg, og) else 1.0, worst_g], device='cuda', dtype=torch.float64)
    mark(rank, 'oracle done')
    dist.all_reduce(res, op=dist.ReduceOp.MAX)
    mark(rank, 'final all_reduce done')
    if rank == 0:
        print('DP_CHECK world=%d worst_D_grad_relerr=%.2e worst_G_grad_relerr=%.2e params_diverged_D=%d params_diverged_G=%d buckets=%s' % (
            world, float(res[0]), float(res[3]), int(res[1]), int(res[2]), nb))
        sys.stdout.flush()
        assert float(res[0]) < 3e-2 and float(res[3]) < 3e-2 and int(res[1]) == 0 and int(res[2]) == 0
    from robust_conditional_gan_b200.parallel import shutdown
    shutdown([model])


if __name__ == '__main__':
    try:
        main()
    except BaseException:
        # a rank that dies with captured NCCL graphs alive can block in interpreter teardown: report and leave at once
        import traceback
        traceback.print_exc()
        sys.stderr.flush()
        os._exit(1)
