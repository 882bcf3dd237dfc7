"""The drop-in entry points end to end on the GPU (synthetic data): mnist/main.py and cifar10/gan_resnet.py flag sets -> training
iterations through the double-buffered feeds -> checkpoints under TF names -> restore -> sample grids / label recovery."""
import glob
import os

import numpy as np
import pytest
import torch

from robust_conditional_gan_b200 import _C, checkpoint, flags as flags_lib, utils
from robust_conditional_gan_b200._C import call
from util import st

pytestmark = pytest.mark.gpu


def test_mnist_main_trains_saves_restores_and_recovers(lib, tmp_path):
    from robust_conditional_gan_b200 import main as M
    argv = ['main.py', '--model', 'rcganu', '--alpha', '0.5', '--batch_size', '64', '--epoch', '1', '--max_iters', '4', '--synthetic_size',
            '512', '--checkpoint_dir', str(tmp_path), '--checkpoint', 'run', '--train', '--norecover']
    m1 = flags_lib.run(M.main, M.flags, argv)
    ck = os.path.join(str(tmp_path), 'run', m1.model_dir)
    assert m1.model_dir == 'mnist_64_28_28' and checkpoint.latest_checkpoint(ck) is not None
    # second process-equivalent: same flags without --train -> load() finds the checkpoint (mnist/main.py:137-140), no training
    argv2 = [a if a != '--train' else '--notrain' for a in argv]      # (the module-level FLAGS object persists inside one process)
    m2 = flags_lib.run(M.main, M.flags, argv2)
    for n, v in m1.store.vars.items():
        assert torch.equal(v.data, m2.store.vars[n].data), n
    for k in m1.groups:
        assert torch.equal(m1.groups[k].m, m2.groups[k].m) and m1.groups[k].t == m2.groups[k].t
    # recover_labels(config) as mnist/main.py:142 calls it: restores the checkpoint, random rows of data_X, SGD on z / y logits
    M.FLAGS.recover_batch_size, M.FLAGS.recover_epoch = 8, 5
    y_rec, mse, zo = m2.recover_labels(M.FLAGS)
    assert y_rec.shape == (8, 10) and np.isfinite(mse) and 0.0 <= zo <= 1.0
    assert abs(float(y_rec.sum(-1).mean()) - 1.0) < 1e-5


def test_cifar_main_trains_with_prefetch_device_rng_checkpoints_and_samples(lib, tmp_path):
    from robust_conditional_gan_b200.cifar import main as M
    argv = ['gan_resnet.py', '--model', 'rcganu', '--alpha', '0.5', '--batch_size', '8', '--ngpus', '1', '--niters', '3', '--log_file',
            str(tmp_path / 'run.log'), '--parent_dir', str(tmp_path), '--expt_dir', 'e', '--sample_freq', '2', '--rng', 'device']
    M.flags.FLAGS.parse(argv[1:])
    cfg = M.configure(M.FLAGS)
    model = M.train(M.FLAGS, cfg)
    assert model.groups['d'].t == 15 and model.groups['g'].t == 2 and model.groups['c'].t == 2      # iteration 0 has no G step
    assert os.path.exists(os.path.join(cfg['DIR'], 'samples_1.png'))
    assert utils.read_png(os.path.join(cfg['DIR'], 'samples_1.png')).shape == (320, 320, 3)
    ck = checkpoint.latest_checkpoint(cfg['CHECKPOINT_DIR'])
    assert ck is not None and checkpoint.step_of(ck) == 2
    out = model.fetch_losses()
    assert all(np.isfinite(v) for v in out.values()), out
    # restore (gan_resnet.py:909-914) continues after the stored iteration and reproduces the parameters
    M.FLAGS.niters = 4
    cfg2 = M.configure(M.FLAGS)
    m2 = M.train(M.FLAGS, cfg2, max_iters=0)
    for n, v in model.store.vars.items():
        assert torch.equal(v.data, m2.store.vars[n].data), n
    log = open(str(tmp_path / 'run.log')).read()
    assert 'restore model from' in log


def test_prefetcher_equals_direct_feed(lib):
    from robust_conditional_gan_b200.cifar.gan_resnet import RCGANCifar, default_flags
    from robust_conditional_gan_b200.feeds import Prefetcher
    model = RCGANCifar(default_flags(algorithm='rcgan', alpha=0.5), tower_batch=4, precision='bf16', dim=128, use_cuda_graph=False)
    rs = np.random.RandomState(0)
    batches = [dict(all_real_data_int=rs.randint(0, 256, size=(4, 3072)), all_real_labels=rs.randint(10, size=4),
                    all_random_labels=rs.randint(10, size=4), all_labels_biased=rs.randint(10, size=4),
                    all_labels_inv_weights=rs.rand(4, 10), noise=rs.randn(4, 128), dequant_noise=rs.rand(4, 3072) / 128) for _ in range(5)]
    pf = Prefetcher(model.d_prog, 'cuda')
    pf.prefetch(**batches[0])
    for i in range(5):
        pf.commit()
        if i + 1 < 5:
            pf.prefetch(**batches[i + 1])          # staged while "step i" would run
        torch.cuda.synchronize()
        got = {k: t.data.clone() for k, t in model.d_prog.inputs.items()}
        model.feed(model.d_prog, **batches[i])
        torch.cuda.synchronize()
        for k, t in model.d_prog.inputs.items():
            assert torch.equal(got[k], t.data), (i, k)
    assert model.d_prog.inputs['all_real_data_int'].data.dtype == torch.uint8


def test_uint8_preprocess_equals_int32(lib):
    rs = np.random.RandomState(1)
    raw = rs.randint(0, 256, size=(6, 3072))
    noise = torch.as_tensor(rs.rand(6, 3072) / 128, dtype=torch.float32).cuda()
    a32, a8 = torch.as_tensor(raw, dtype=torch.int32).cuda(), torch.as_tensor(raw, dtype=torch.uint8).cuda()
    for dt, td in ((_C.F32, torch.float32), (_C.BF16, torch.bfloat16)):
        y32, y8 = torch.zeros(6, 32, 32, 3, device='cuda', dtype=td), torch.zeros(6, 32, 32, 3, device='cuda', dtype=td)
        call('rcgan_preprocess_cifar', a32.data_ptr(), noise.data_ptr(), y32.data_ptr(), 6, dt, st())
        call('rcgan_preprocess_cifar_u8', a8.data_ptr(), noise.data_ptr(), y8.data_ptr(), 6, dt, st())
        assert torch.equal(y32, y8)
    ref = (2 * (raw / 256. - .5) + noise.cpu().numpy()).reshape(6, 3, 32, 32).transpose(0, 2, 3, 1)       # gan_resnet.py:548-552
    assert float((y32.float().cpu() - torch.as_tensor(ref)).abs().max()) < 1e-2
    y32f = torch.zeros(6, 32, 32, 3, device='cuda')
    call('rcgan_preprocess_cifar_u8', a8.data_ptr(), noise.data_ptr(), y32f.data_ptr(), 6, _C.F32, st())
    assert float((y32f.cpu() - torch.as_tensor(ref)).abs().max()) < 1e-6


def test_in_graph_random_inputs(lib):
    """rcgan_random_fill: N(0,1) and U[0,1/128) moments, a new draw per step, the same draw for the same (seed, step, stream)"""
    n = 1 << 20
    step = torch.zeros(1, dtype=torch.int64, device='cuda')
    a, b, c = (torch.zeros(n, device='cuda') for _ in range(3))
    call('rcgan_random_fill', a.data_ptr(), n, 1, 0.0, 1.0, 1234, step.data_ptr(), 1, st())
    call('rcgan_random_fill', b.data_ptr(), n, 1, 0.0, 1.0, 1234, step.data_ptr(), 1, st())
    assert torch.equal(a, b)
    assert abs(float(a.mean())) < 5e-3 and abs(float(a.std()) - 1.0) < 5e-3 and float(a.abs().max()) < 7.0
    assert abs(float((a ** 4).mean()) - 3.0) < 0.05                                 # kurtosis of a normal
    step.fill_(1)
    call('rcgan_random_fill', c.data_ptr(), n, 1, 0.0, 1.0, 1234, step.data_ptr(), 1, st())
    assert not torch.equal(a, c) and abs(float((a * c).mean())) < 5e-3               # independent of the previous step
    call('rcgan_random_fill', b.data_ptr(), n, 1, 0.0, 1.0, 1234, step.data_ptr(), 2, st())
    assert abs(float((b * c).mean())) < 5e-3                                         # and of the other stream of the same step
    u = torch.zeros(n + 3, device='cuda')
    call('rcgan_random_fill', u.data_ptr(), n + 3, 0, 0.0, 1.0 / 128, 99, step.data_ptr(), 3, st())
    assert float(u.min()) >= 0.0 and float(u.max()) < 1.0 / 128
    assert abs(float(u.mean()) * 256 - 1.0) < 5e-3 and abs(float(u.var()) * 12 * 128 * 128 - 1.0) < 1e-2
