"""Debug helper: count ReLU branch flips between the fp32 product and the fp64 oracle in the CIFAR G step."""
import sys
import torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from test_gpu_cifar import build, feed_d, feed_g
from oracle import cifar as OC
from robust_conditional_gan_b200 import nnops, _C
model, tr, b = build('biased', 6, 'fp32', 32)
feed_d(model, b); feed_g(model, b)
tr.d_step(b, 0); model.d_step(0)
torch.cuda.synchronize()
model.store.load_state_dict({k: v.detach() for k, v in tr.P.items()})
rec = []
orig = torch.relu
def relu(x):
    rec.append(x.detach())
    return orig(x)
OC.torch.relu = relu
tr.g_step(b, 1)
OC.torch.relu = orig
model._body_a(model.g_prog, ('g', 'c'))
torch.cuda.synchronize()
prod = []
for o in model.g_prog.ops:
    if isinstance(o, nnops.BatchNormOp) and o.act == _C.ACT_RELU: prod.append(('cbn', o.y.torch() > 0))
    elif isinstance(o, nnops.ConvOp) and o.act == _C.ACT_RELU: prod.append(('conv', o.y.torch() > 0))
    elif isinstance(o, nnops.ActOp) and o.act == _C.ACT_RELU: prod.append(('act', o.y.torch() > 0))
    elif isinstance(o, nnops.MeanHWOp) and o.relu: prod.append(('mean', o.x.torch() > 0))
print(len(rec), len(prod))
tot = 0
for (kind, pm), ox in zip(prod, rec):
    om = ox > 0
    if tuple(pm.shape) != tuple(om.shape):
        print('shape mismatch', kind, tuple(pm.shape), tuple(om.shape)); continue
    mism = (pm.cpu() != om)
    k = int(mism.sum())
    tot += k
    if k:
        print(kind, tuple(om.shape), 'flips', k, 'oracle |pre| at flips', ox[mism].abs().max().item(), 'numel', om.numel())
print('total flips', tot)
