"""Benchmark of the RCGAN training hot path (BASELINE.json metric: RCGAN G+D train images/sec).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W     # the reference-equivalent CPU path (oracle)

A "step" is one full training iteration of the reference hot loop.  Default workload at N=1 (and per GPU for N>1,
weak scaling): BASELINE configs[3], the CIFAR-10 SN-ResNet RCGAN with projection discriminator, bf16, batch 256 per
GPU, synthetic 32x32x3 -- the largest single-GPU configuration of BASELINE.json; at N=8 the global batch is 2048
(configs[4]).  One iteration = gan_resnet.py:919-947 = 1 G step (batch 2B) + 5 D steps (B real + B fake each);
images/sec = real images consumed per second = 5B / t_iter.  The other configs (`--workload ...`: RCGAN-U
with learned confusion matrix + permutation classifier, the MNIST DCGAN configs, label recovery) are timed briefly as
`secondary` records of the same JSON line at N=1.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE configs[3] / [4]: CIFAR-10 SN-ResNet, batch = per-GPU BATCH_SIZE (one tower per GPU); a step = one
    # iteration of gan_resnet.py:919-947 = 1 G step (batch 2B) + 5 D steps (B real + B fake each) -> 5B real images
    'cifar_rcgan_b256': dict(kind='cifar', batch=256, flags=dict(algorithm='rcgan', perm_classifier=False)),
    'cifar_rcganu_b256': dict(kind='cifar', batch=256, flags=dict(algorithm='rcgan-u', perm_classifier=True, confuse_init=True)),
    'mnist_rcganu_b1024': dict(batch=1024, flags=dict(algorithm='rcgan', disc_type='projection', estimate_confuse=True)),
    'mnist_rcgan_b64': dict(batch=64, flags=dict(algorithm='rcgan', disc_type='projection', estimate_confuse=False)),
    # DCGAN.recover_labels (model.py:494-640), recover_batch_size 500 -> 5000 generated images per step;
    # a step = one gradient-descent step on (z_recover, y_logit_recover); value counts the R real images per step
    'mnist_recover_r500': dict(kind='recover', batch=500, flags=dict(algorithm='rcgan', disc_type='projection', estimate_confuse=True)),
    'mnist_rcgany_b1024': dict(batch=1024, flags=dict(algorithm='rcgan', disc_type='projection', estimate_confuse=False,
                                                      concat_y=True, concat_y_layers=[1])),
}
METRIC = 'RCGAN G+D train images/sec'


def metric_of(wl):
    return 'recover_labels real images x steps/sec' if wl.get('kind') == 'recover' else METRIC


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q, '--format=csv,noheader,nounits',
                                          '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons), 'samples': len(sm)}


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d['hbm_gbs'], tf_burst=d['bf16_tflops'], tf_sust=d['bf16_tflops_sustained'], src='measured')
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src='fallback')


def synthetic(B, seed=0):
    """Config-2 inputs: X~U[0,1) 28x28x1, one-hot labels from randint (the label sampler is not in the step)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    eye = torch.eye(10)
    lab = lambda: eye[torch.randint(0, 10, (B,), generator=g)]
    d = dict(batch_images=torch.rand(B, 28, 28, 1, generator=g), batch_z=torch.rand(B, 100, generator=g) * 2 - 1,
             batch_labels_real=lab(), batch_labels_gen=lab(), batch_labels_fake=lab(), batch_labels_real_weights=lab())
    return {k: v.contiguous().pin_memory() if torch.cuda.is_available() else v for k, v in d.items()}


def oracle_iteration_timer(B, flags, threads):
    """The reference-equivalent CPU path: the oracle restatement (PyTorch-CPU fp32, literal reference graph incl.
    its 10 label-wise discriminator calls) -- TensorFlow 1.5 itself cannot be installed here."""
    import torch
    from oracle import mnist as OM, sampler as OS
    torch.set_num_threads(threads)
    cfg = OM.default_config(batch_size=B, alpha=0.5, perm_regularizer=True, **flags)
    P = OM.init_params(cfg, 0, torch.float32)
    C = OS.one_coin_confusion(0.5)
    b = OM.synthetic_batch(B, 0, torch.float32, C, cfg)
    tr = OM.Trainer(P, cfg, C)

    def step():
        t = time.perf_counter()
        tr.iteration(b)
        return time.perf_counter() - t
    return step


def oracle_recover_timer(R, flags, threads):
    """CPU port of one recover_labels step (oracle/mnist.py recover_step)."""
    import torch
    from oracle import mnist as OM
    torch.set_num_threads(threads)
    cfg = OM.default_config(batch_size=64, alpha=0.5, perm_regularizer=True, **flags)
    P = OM.init_params(cfg, 0, torch.float32)
    g = torch.Generator().manual_seed(0)
    z = torch.rand(R * 10, 100, generator=g) - 0.5
    yl = torch.randn(R, 10, generator=g)
    actual = torch.rand(R, 28, 28, 1, generator=g)
    st = {'z': z, 'yl': yl}

    def step():
        t = time.perf_counter()
        st['z'], st['yl'], _, _ = OM.recover_step(P, st['z'], st['yl'], actual, cfg, lr=500.0)
        return time.perf_counter() - t
    return step


STEP_DESC = {'cifar': '1 G step (batch 2B) + 5 D steps (B real + B fake)',
             'recover': '1 recover_labels step: gen_sampler on 10 B images, mse, SGD on z_recover / y_logit_recover'}


def config_of(name, wl, world):
    """the workload description both arms print (same keys and values: the driver compares them)"""
    B = wl['batch']
    # bytes one iteration streams (activations + their gradients, bf16): ~12 MB per CIFAR image and iteration (the G step alone holds eight 268 MB tensors at 512x32x32x256), ~0.35 MB per MNIST
    # image (g_h2's 14x14x128 output dominates); recover_labels generates 10 images per real one
    imgs = B * (10 if wl.get('kind') == 'recover' else 1)
    mb = imgs * (12.0 if wl.get('kind') == 'cifar' else 0.35)
    if mb > 126:
        l2 = 'no explicit flush: one iteration streams ~%.1f GB of activations/gradients, > 126 MB L2' % (mb / 1e3)
    else:
        l2 = ('no explicit flush: one iteration streams ~%d MB, which fits the 126 MB L2 -- an L2-warm number, as steady-state '
              'training of this small config is' % mb)
    return {'workload': name, 'batch_per_gpu': B, 'global_batch': world * B,
            'step': STEP_DESC.get(wl.get('kind'), '1 D step + 2 G(+C) steps'), 'parallelism': 'dp%d' % world, 'l2': l2}


def cpu_sample(wl, cores, timed, warm, budget_s):
    """The reference-equivalent CPU path (oracle port; TensorFlow 1.5 cannot run here) on a BOUNDED sample of the workload
    at the workload's own batch.  Returns (images/s, seconds per iteration, description, units timed).
    CIFAR: one iteration = 1 G step (batch 2B) + 5 D steps; the sample is ONE G step plus `timed` D steps (after `warm`),
    iteration time = t_G + 5 * mean(t_D).  MNIST / recover: `timed` whole iterations / steps.  Stops early at budget_s."""
    B = wl['batch']
    t_start = time.perf_counter()
    kind = wl.get('kind')
    if kind == 'cifar':
        import torch
        from oracle import cifar as OC
        torch.set_num_threads(cores)
        cfg = OC.default_config(alpha=0.5, **wl['flags'])
        P = OC.init_params(cfg, seed=0, dtype=torch.float32)
        b = OC.synthetic_batch(B, seed=0, dtype=torch.float32)
        tr = OC.Trainer(P, cfg)
        for _ in range(warm):
            tr.d_step(b, it=1)
        t = time.perf_counter(); tr.g_step(b, it=1); t_g = time.perf_counter() - t
        ts = []
        for _ in range(timed):
            t = time.perf_counter(); tr.d_step(b, it=1); ts.append(time.perf_counter() - t)
            if time.perf_counter() - t_start > budget_s:
                break
        t_it = t_g + 5 * sum(ts) / len(ts)
        return 5 * B / t_it, t_it, ('oracle port, fp32, %d threads, tower batch %d (the workload\'s): 1 G step (batch %d) timed once '
                                    '(%.2f s) + %d timed D steps after %d warm-up (mean %.2f s); iteration = t_G + 5 t_D'
                                    % (cores, B, 2 * B, t_g, len(ts), warm, sum(ts) / len(ts))), len(ts)
    if kind == 'recover':
        R = min(B, 50)
        step = oracle_recover_timer(R, wl['flags'], cores)
        per, what = R, 'oracle recover_step at recover_batch_size %d (%d generated images)' % (R, 10 * R)
    else:
        step = oracle_iteration_timer(B, wl['flags'], cores)
        per, what = B, 'oracle iteration (1 D + 2 G steps, literal reference graph incl. its 10 label-wise D calls) at batch %d' % B
    for _ in range(warm):
        step()
    ts = []
    for _ in range(timed):
        ts.append(step())
        if time.perf_counter() - t_start > budget_s:
            break
    t_it = sum(ts) / len(ts)
    return per / t_it, t_it, '%s, fp32, %d threads: %d timed after %d warm-up' % (what, cores, len(ts), warm), len(ts)


def run_reference(args, wl):
    """--impl reference: the reference's own CPU implementation of the path = the oracle port (the reference is TF-1.5
    Python, not installable here), all host threads, same config as our arm; every step is a bounded sample (see cpu_sample)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count()
    v, t_it, sample, n = cpu_sample(wl, cores, args.steps, min(args.warmup, 2), budget_s=240.0)
    print(json.dumps({
        'impl': 'reference', 'metric': metric_of(wl), 'value': v, 'unit': 'images/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': t_it * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic', 'config': config_of(args.workload, wl, max(args.gpus, 1)), 'units_timed': n,
        'cpu_baseline': {'value': v, 'unit': 'images/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': v, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))


def op_profile(model, reps=5):
    """Per-op device time (CUDA events on the launching stream, L2 flushed between reps) of both programs; returns
    the dominant op with its algorithmic work for the roofline entry."""
    import torch
    from robust_conditional_gan_b200 import nnops
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    rows = []
    for prog in ((model.d_prog,) if model.d_prog is model.g_prog else (model.d_prog, model.g_prog)):
        prog.run_forward(); prog.run_backward()
        for op in prog.ops:
            for direction in ('forward', 'backward'):
                fn = getattr(op, direction)
                ts = []
                for _ in range(reps):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); fn(prog); prog.join(); e1.record(); e1.synchronize()
                    ts.append(e0.elapsed_time(e1))
                ms = sorted(ts)[len(ts) // 2]
                flops = 0
                if isinstance(op, (nnops.ConvOp, nnops.DeconvOp)):
                    d = op.desc
                    f1 = 2.0 * d.n * d.ho * d.wo * d.cout * d.kh * d.kw * d.cin
                    if direction == 'forward':
                        flops = f1
                    else:
                        flops = f1 * (int(op.need[0]) + int(op.need[1])) if any(op.need) else 0
                rows.append(dict(prog=prog.name, op=type(op).__name__, shape=str(tuple(op.outputs[0].shape)) if op.outputs else '',
                                 dir=direction, ms=ms, flops=flops))
    return rows


def kernel_roofline(model, rows, pk):
    """Roofline entry of the dominant KERNEL: the tcgen05 implicit-GEMM conv (conv_tc family, fprop / dgrad form).
    The conv layer whose forward costs most in the op profile is re-launched bare through the C ABI on the op's own
    resident buffers (no weight pack, no patch build), 20 launches, CUDA events on the launching stream around each launch,
    L2 flushed in between; achieved = algorithmic flops (2*M*N*K of the layer) / median launch time."""
    import json as _json
    import torch
    from robust_conditional_gan_b200 import _C, nnops
    from robust_conditional_gan_b200.nnops import dp, pp
    lib = _C.load()
    cands = []
    for prog in ((model.d_prog,) if model.d_prog is model.g_prog else (model.d_prog, model.g_prog)):
        for op in prog.ops:
            if isinstance(op, nnops.ConvOp) and op.patch is None and op.pack is not None and lib.rcgan_conv_uses_tensor_cores(op.desc, 0):
                cands.append((prog, op, 'fprop'))
            elif isinstance(op, nnops.DeconvOp) and op.patch is None and op.pack is not None and lib.rcgan_conv_uses_tensor_cores(op.desc, 1):
                cands.append((prog, op, 'dgrad'))
    if not cands:
        return None
    key = lambda prog, op: (prog.name, type(op).__name__, str(tuple(op.outputs[0].shape)), 'forward')
    ms_of = {}
    for r in rows:
        ms_of.setdefault((r['prog'], r['op'], r['shape'], r['dir']), 0.0)
        ms_of[(r['prog'], r['op'], r['shape'], r['dir'])] = max(ms_of[(r['prog'], r['op'], r['shape'], r['dir'])], r['ms'])
    prog, op, form = max(cands, key=lambda c: ms_of.get(key(c[0], c[1]), 0.0))
    d = op.desc
    st = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')

    def launch():
        if form == 'fprop':
            _C.call('rcgan_conv2d_fprop', d, dp(op.x), dp(op.w), pp(op.pack), dp(op.b), dp(op.y), op.y.dtype, op.act, op.leak, st)
        else:
            _C.call('rcgan_conv2d_dgrad', d, dp(op.x), dp(op.w), pp(op.pack), dp(op.b), dp(op.y), op.y.dtype, op.act, op.leak, 0, st)
    launch()
    ts = []
    for _ in range(20):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); launch(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    flops = 2.0 * d.n * d.ho * d.wo * d.cout * d.kh * d.kw * d.cin
    # a stride-2 transposed conv is ONE persistent launch whose tile list covers the 4 output parity classes
    nl = 1
    ach = flops / (ms * 1e-3) / 1e12
    fam = sum(r['ms'] for r in rows if r['op'] in ('ConvOp', 'DeconvOp'))
    tot = sum(r['ms'] for r in rows)
    name = 'conv_tc %s n%d %dx%dx%d -> %dx%dx%d k%d s%d (%s %s)' % (form, d.n, d.h, d.w, d.cin, d.ho, d.wo, d.cout, d.kh, d.stride,
                                                                   prog.name, type(op).__name__)
    # DRAM bytes per launch: not measurable inside the run (it needs the profiler) -- taken from the committed `ncu --set full`
    # capture of this very kernel and shape; traffic_source names the file, null when there is no capture of the shape
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, 'profiles', 'kernel_traffic.json')
    if os.path.exists(tp):
        kt = _json.load(open(tp))
        traffic = kt.get(name.split(' (')[0])
        traffic_src = kt.get('_source', {}).get(name.split(' (')[0]) if traffic is not None else None
    return {'bound': 'tensor', 'achieved': ach, 'peak': pk['tf_burst'], 'unit': 'TFLOP/s', 'frac': ach / pk['tf_burst'],
            'traffic': traffic, 'traffic_source': traffic_src, 'algorithmic_bytes': 2.0 * (d.n * d.h * d.w * d.cin + d.n * d.ho * d.wo * d.cout) + 2.0 * d.kh * d.kw * d.cin * d.cout,
            'peak_source': pk['src'] + ' bf16 burst (kernel timed alone)', 'kernel': name,
            'launch_ms': ms / nl, 'flops_per_launch': flops / nl, 'launches_per_call': nl, 'conv_family_share_of_step': fam / tot}


def sampler_bench(n=70000):
    """Label-noise sampler (mnist/model.py:795-834) on the full 70 000-sample MNIST label set: CUDA kernel (device time of
    the sampling launch, labels resident) against the reference's literal numpy loop (oracle port) on a 7 000-sample slice
    scaled to 70 000; outputs compared element by element (bit-exact)."""
    import numpy as np
    import torch
    from oracle import sampler as OS
    from robust_conditional_gan_b200.sampler import LabelNoiseSampler, one_coin_confusion
    C = one_coin_confusion(0.5)
    y = np.random.RandomState(1).randint(10, size=n)
    smp = LabelNoiseSampler('cuda')
    smp.load_mnist_labels(y, C)                     # warm-up (table upload, allocations)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = smp.load_mnist_labels(y, C); e1.record(); e1.synchronize()
    gpu_ms = e0.elapsed_time(e1)
    m = n // 10
    t = time.perf_counter(); ref = OS.mnist_labels_numpy(y[:m], C); cpu_s = (time.perf_counter() - t) * (n / m)
    small = smp.load_mnist_labels(y[:m], C)
    exact = bool((small['real'] == ref['y_real'].argmax(1)).all() and (small['fake'] == ref['y_fake'].argmax(1)).all()
                 and (small['gen'] == ref['y_gen'].argmax(1)).all() and (small['perm'] == ref['perm']).all())
    return {'samples': n, 'gpu_ms_incl_shuffles_and_readback': gpu_ms, 'cpu_numpy_loop_s_scaled_from_%d' % m: cpu_s,
            'bit_exact_vs_numpy': exact}


class CifarBench:
    """adapter giving RCGANCifar the same (feed / train_iteration / programs) surface bench.py drives for DCGAN"""

    def __init__(self, B, flags, precision, world, rank):
        import torch
        from robust_conditional_gan_b200.cifar.gan_resnet import RCGANCifar, default_flags
        self.m = RCGANCifar(default_flags(alpha=0.5, **flags), tower_batch=B, precision=precision, world_size=world, rank=rank)
        self.m.iteration = 1                       # steady state: every iteration has its G step
        self.d_prog, self.g_prog, self.B = self.m.d_prog, self.m.g_prog, B
        g = torch.Generator().manual_seed(rank)
        ri = lambda n: torch.randint(0, 10, (n,), generator=g).to(torch.int32)
        pin = lambda t: t.contiguous().pin_memory()
        # 5 distinct real batches per iteration (uint8-range ints as int32, like the reference's placeholders)
        self.d_feeds = [dict(all_real_data_int=pin(torch.randint(0, 256, (B, 3072), generator=g).to(torch.int32)),
                             all_real_labels=pin(ri(B)), all_random_labels=pin(ri(B)), all_labels_biased=pin(ri(B)),
                             all_labels_inv_weights=pin(torch.eye(10)[ri(B).long()]), noise=pin(torch.randn(B, 128, generator=g)),
                             dequant_noise=pin(torch.rand(B, 3072, generator=g) / 128)) for _ in range(5)]
        self.g_feeds = dict(noise=pin(torch.randn(2 * B, 128, generator=g)), all_random_labels_G=pin(ri(2 * B)),
                            all_labels_biased_G=pin(ri(2 * B)))
        self.images_per_step = 5 * B
        self.h2d = sum(v.numel() * v.element_size() for f in self.d_feeds for v in f.values()) + \
            sum(v.numel() * v.element_size() for v in self.g_feeds.values())
        self.d2h = 4 * (len(self.d_prog.loss_names) + len(self.g_prog.loss_names))

    def set_graph(self, on):
        self.m.use_cuda_graph = on

    def resident(self):
        self.m.feed(self.d_prog, **self.d_feeds[0]); self.m.feed(self.g_prog, **self.g_feeds)

    def step(self, e2e):
        if e2e:
            return self.m.train_iteration(self.d_feeds, self.g_feeds, fetch=True)
        return self.m.train_iteration(None, None, fetch=False)


class RecoverBench:
    """DCGAN.recover_labels steps: R real images against R*10 gen_sampler images, SGD on z_recover / y_logit_recover."""

    def __init__(self, R, flags_kw, precision, world, rank):
        import torch
        from robust_conditional_gan_b200.model import DCGAN, default_flags
        flags = default_flags(batch_size=64, alpha=0.5, **flags_kw)
        self.m = DCGAN(batch_size=64, algorithm=flags.algorithm, estimate_confuse=flags.estimate_confuse, perm_regularizer=True,
                       alpha=0.5, disc_type=flags.disc_type, config=flags, precision=precision, world_size=1, rank=0, seed=0)
        self.m.build_recover(R, seed=rank)
        self.d_prog = self.g_prog = self.m.r_prog
        g = torch.Generator().manual_seed(rank)
        self.actual = torch.rand(R, 28, 28, 1, generator=g).contiguous().pin_memory()
        self.images_per_step = R
        self.h2d = self.actual.numel() * 4
        self.d2h = 4

    def set_graph(self, on):
        self.m.use_cuda_graph = on

    def resident(self):
        self.m.recover_step(self.actual, 500.0, fetch=False)

    def step(self, e2e):
        if e2e:
            return {'mse_loss': self.m.recover_step(self.actual, 500.0, fetch=True)}
        self.m.recover_step(None, 500.0, fetch=False)
        return None


class MnistBench:
    def __init__(self, B, flags_kw, precision, world, rank):
        from robust_conditional_gan_b200.model import DCGAN, default_flags
        flags = default_flags(batch_size=B, alpha=0.5, **flags_kw)
        self.m = DCGAN(batch_size=B, algorithm=flags.algorithm, estimate_confuse=flags.estimate_confuse, perm_regularizer=True,
                       alpha=0.5, disc_type=flags.disc_type, config=flags, precision=precision, world_size=world, rank=rank,
                       seed=0)
        self.d_prog, self.g_prog, self.B = self.m.d_prog, self.m.g_prog, B
        self.feeds = synthetic(B, seed=rank)
        self.images_per_step = B
        self.h2d = sum(v.numel() * 4 for v in self.feeds.values()) + \
            sum(self.feeds[k].numel() * 4 for k in ('batch_z', 'batch_labels_gen', 'batch_labels_fake'))
        self.d2h = 4 * (len(self.d_prog.loss_names) + len(self.g_prog.loss_names))

    def set_graph(self, on):
        self.m.use_cuda_graph = on

    def resident(self):
        self.m.feed(**self.feeds)

    def step(self, e2e):
        if e2e:
            return self.m.train_iteration(fetch_losses=True, **self.feeds)
        return self.m.train_iteration(fetch_losses=False)


def make_bench(wl, precision, world, rank):
    return {'cifar': CifarBench, 'recover': RecoverBench}.get(wl.get('kind'), MnistBench)(wl['batch'], wl['flags'], precision, world, rank)


def time_workload(bench, steps, warmup, world, local, sample_clocks):
    """W warm-up iterations, then (1) `steps` iterations with inputs resident in HBM, device-timed with CUDA events, and
    (2) `steps` iterations end to end through the public API (pinned-host inputs copied every step, losses read back).
    Barrier + synchronize on both sides of each leg, max over ranks."""
    import torch
    import torch.distributed as dist
    from robust_conditional_gan_b200 import _C
    lib = _C.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    bench.set_graph(False)                   # launches per iteration: count once with eager (uncaptured) launches
    n0 = lib.rcgan_launch_count()
    bench.step(True)
    launches_per_iter = lib.rcgan_launch_count() - n0
    bench.set_graph(True)
    for _ in range(max(warmup, 3)):
        bench.step(True)
    bench.resident()
    clocks = ClockSampler(local) if sample_clocks else None
    barrier()
    if clocks:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        bench.step(False)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = bench.step(True)
    barrier()
    e2e_s = time.perf_counter() - t0
    clk = clocks.stop() if clocks else None
    if world > 1:
        t = torch.tensor([ms, e2e_s], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0]), float(t[1])
    ms_step = ms / steps
    return dict(ms_per_step=ms_step, value=world * bench.images_per_step / (ms_step / 1e3),
                e2e=world * bench.images_per_step * steps / e2e_s, launches_per_iter=launches_per_iter, clocks=clk, losses=out)


SECONDARY = ('cifar_rcganu_b256', 'mnist_rcganu_b1024', 'mnist_rcgany_b1024', 'mnist_rcgan_b64')


def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    B = wl['batch']
    bench = make_bench(wl, args.precision, world, rank)
    dp_model = bench.m
    r = time_workload(bench, args.steps, args.warmup, world, local, True)
    h2d, d2h = bench.h2d, bench.d2h
    if rank != 0:
        if world > 1:
            from robust_conditional_gan_b200.parallel import shutdown
            shutdown([bench.m])
        return
    pk = peaks()
    result = {
        'metric': metric_of(wl), 'value': r['value'], 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': r['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'bf16' if args.precision == 'bf16' else 'f32', 'data': 'synthetic',
        'config': config_of(args.workload, wl, world),
        'clocks': r['clocks'],
        'e2e': {'value': r['e2e'], 'unit': 'images/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
        'gpu_launches': r['launches_per_iter'] * args.steps,
        'losses': {k: round(v, 5) for k, v in r['losses'].items()},
    }
    if wl.get('kind') == 'cifar':
        # step-level utilisation on the REFERENCE formulation's flops (15.56 TFLOP per iteration at B=256, incl. the
        # generator forward of every D step; the folded pool / upsample convs issue fewer) against the sustained bf16 rate
        tf_iter = 15.56e12 * B / 256
        result['step_tflops'] = {'algorithmic_tflop_per_iteration': tf_iter / 1e12, 'achieved': tf_iter / (r['ms_per_step'] * 1e-3) / 1e12,
                                 'peak_sustained': pk['tf_sust'], 'frac': tf_iter / (r['ms_per_step'] * 1e-3) / 1e12 / pk['tf_sust']}
    if world == 1 and not args.no_op_profile:
        # roofline of the dominant kernel (timed live, CUDA events, L2 flushed) ...
        rows = op_profile(bench)
        rf = kernel_roofline(bench, rows, pk)
        if rf is not None:
            result['roofline'] = rf
        else:
            top = max(rows, key=lambda r: r['ms'])
            result['roofline'] = {'bound': 'hbm', 'achieved': None, 'peak': pk['hbm'], 'unit': 'GB/s', 'frac': None,
                                  'traffic': None, 'kernel': '%s %s %s' % (top['prog'], top['op'], top['dir'])}
        result['op_profile_top'] = [dict(r, ms=round(r['ms'], 4)) for r in sorted(rows, key=lambda r: -r['ms'])[:8]]
        os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
        with open(os.path.join(ROOT, 'gpurun_out', 'op_profile.json'), 'w') as f:
            json.dump(rows, f, indent=1)
    del bench
    torch.cuda.empty_cache()
    if world == 1 and not args.no_secondary:
        # the other BASELINE configs, timed briefly (same legs, fewer steps); parity of each is in tests/
        sec = {}
        for name in SECONDARY:
            if name == args.workload:
                continue
            w2 = WORKLOADS[name]
            b2 = make_bench(w2, args.precision, 1, 0)
            r2 = time_workload(b2, 10, 3, 1, local, False)
            sec[name] = {'value': r2['value'], 'unit': 'images/s', 'ms_per_step': r2['ms_per_step'], 'e2e': r2['e2e'],
                         'gpu_launches_per_step': r2['launches_per_iter'], 'batch_per_gpu': w2['batch']}
            del b2
            torch.cuda.empty_cache()
        result['secondary'] = sec
        result['sampler'] = sampler_bench()
    if world == 1 and not args.no_cpu_baseline:
        # the CPU baseline (oracle port) on a bounded sample of the same workload at the same batch
        cores = os.cpu_count()
        v, t_it, sample, n = cpu_sample(wl, cores, 3, 1, budget_s=45.0)
        result['cpu_baseline'] = {'value': v, 'unit': 'images/s', 'cores': cores, 'kind': 'port', 'sample': sample}
    print(json.dumps(result), flush=True)
    if world > 1:
        from robust_conditional_gan_b200.parallel import shutdown
        shutdown([] if world == 1 else [dp_model])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cifar_rcgan_b256', choices=sorted(WORKLOADS))
    ap.add_argument('--no-secondary', action='store_true', help='skip the secondary workload records')
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-op-profile', action='store_true', help='skip the per-op roofline pass (clean ncu launch lists)')
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == 'reference':
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == '__main__':
    main()
