"""Debug helper: fp32 G-step gradient errors with the product's state re-synchronised to the oracle after the D step."""
import sys
import torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from test_gpu_mnist import build, feed
from util import relerr
for run in ('rcgan', 'rcganu'):
    for sync in (False, True):
        B = 16
        model, tr, batch = build(run, B, 'fp32', use_graph=False)
        feed(model, batch)
        tr.d_step(batch); model.d_step()
        torch.cuda.synchronize()
        if sync:
            model.store.load_state_dict({k: v.detach() for k, v in tr.P.items()})
        tr.g_step(batch); model.g_step()
        torch.cuda.synchronize()
        errs = {}
        for v in model.g_vars + model.c_vars:
            ref = tr.last['g_grads'][v.name]
            if float(ref.norm()) > 1e-9:
                errs[v.name] = relerr(v.grad.reshape(ref.shape), ref)
        print(run, 'sync' if sync else 'nosync', ' '.join('%s=%.1e' % (k.split('/')[-2] + '/' + k.split('/')[-1] if '/' in k else k, e) for k, e in errs.items()))
