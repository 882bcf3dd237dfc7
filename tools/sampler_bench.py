"""Label-noise sampler timings: full load_mnist_labels path and the individual launches (run from the repo root on a B200)."""
import sys, json, time
sys.path.insert(0, ".")
import numpy as np, torch
import bench
from robust_conditional_gan_b200.sampler import LabelNoiseSampler, one_coin_confusion
from robust_conditional_gan_b200._C import call, stream_ptr
print(json.dumps(bench.sampler_bench()))
smp = LabelNoiseSampler('cuda'); C = one_coin_confusion(0.5); n = 70000
y = torch.randint(0, 10, (n,), dtype=torch.int32, device='cuda')
real = torch.zeros_like(y); gen = torch.zeros_like(y); fake = torch.zeros_like(y)
tab = smp.table(C)
def t(fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); e1.synchronize(); return e0.elapsed_time(e1)
print('seed ms', t(lambda: smp.seed(547)))
print('shuffle ms', t(lambda: smp.shuffle_perm(n)))
print('sample ms', t(lambda: call('rcgan_sample_labels_mnist', smp.state.data_ptr(), tab.data_ptr(), 10, y.data_ptr(), n, 0, real.data_ptr(), gen.data_ptr(), fake.data_ptr(), stream_ptr())))
