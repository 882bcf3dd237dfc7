"""Debug helper (not a test): op-by-op comparison of the bf16 program against the fp32 program, and fp32 vs oracle
parameter deltas after Adam.  Run on the GPU box: python tests/debug_parity.py rcgan"""
import sys

import torch

sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
from test_gpu_mnist import build, feed
from util import relerr

run = sys.argv[1] if len(sys.argv) > 1 else 'rcgan'
B = 32
m16, tr, batch = build(run, B, 'bf16', use_graph=False)
m32, _, _ = build(run, B, 'fp32', use_graph=False)
for m in (m16, m32):
    feed(m, batch)
    m._d_body_a()
torch.cuda.synchronize()
print('== D step forward/backward, bf16 vs fp32, op by op')
for o16, o32 in zip(m16.d_prog.ops, m32.d_prog.ops):
    name = type(o16).__name__
    for t16, t32 in zip(o16.outputs, o32.outputs):
        e = relerr(t16.torch().float(), t32.torch().float())
        ge = relerr(t16.grad_torch().float(), t32.grad_torch().float()) if (t16.grad is not None and t32.grad is not None) else -1
        print('%-16s out%-22s fwd %.2e  grad %.2e' % (name, str(tuple(t16.shape)), e, ge))
print('== D var grads bf16 vs fp32 / fp32 vs oracle')
tr.d_step(batch)
for v16, v32 in zip(m16.d_vars, m32.d_vars):
    ref = tr.last['d_grads'][v32.name]
    print('%-40s %.2e  %.2e   |g|=%.2e' % (v16.name, relerr(v16.grad, v32.grad), relerr(v32.grad.reshape(ref.shape), ref), float(ref.norm())))
