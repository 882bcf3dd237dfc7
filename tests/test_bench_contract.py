"""bench.py's reference arm and the JSON contract of both arms, on the CPU: `--impl reference` times the oracle port (the only
place outside tests/ and smoke() that executes oracle/), so it runs here; the repo arm needs a GPU and is checked for the
config it would print."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--workload', 'mnist_rcgan_b64',
                          '--steps', '1', '--warmup', '0'], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == bench.METRIC and d['unit'] == 'images/s' and d['higher_is_better'] is True
    assert d['n_gpus'] == 1 and d['steps'] == 1 and d['warmup'] == 0 and d['value'] > 0 and d['ms_per_step'] > 0
    # both arms describe the workload with the same dict
    assert d['config'] == bench.config_of('mnist_rcgan_b64', bench.WORKLOADS['mnist_rcgan_b64'], 1)
    assert d['e2e'] == {'value': d['value'], 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['value'] == d['value'] and cb['cores'] >= 1 and cb['sample']
    assert d['vs_baseline'] is None and d['data'] == 'synthetic' and d['dtype'] == 'f32'


def test_default_workload_is_the_largest_single_gpu_baseline_config():
    import argparse
    base = json.load(open(os.path.join(ROOT, 'BASELINE.json')))
    assert 'cifar_rcgan_b256' in bench.WORKLOADS
    cfg = bench.config_of('cifar_rcgan_b256', bench.WORKLOADS['cifar_rcgan_b256'], 8)
    assert cfg['global_batch'] == 2048 and cfg['parallelism'] == 'dp8' and 'L2' in cfg['l2']
    assert isinstance(base.get('configs'), list) and len(base['configs']) >= 4
    # every workload states how its timed iterations relate to the 126 MB L2
    for name, wl in bench.WORKLOADS.items():
        assert '126 MB L2' in bench.config_of(name, wl, 1)['l2']
