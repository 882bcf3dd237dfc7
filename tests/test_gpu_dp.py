"""Data-parallel path on real GPUs (skipped on a single-GPU box): tests/dp_check.py under torchrun, 2 ranks over NCCL --
gradients after the in-graph bucketed all-reduce == the reference's 2-tower cost gradients, parameters bit-identical across
ranks after Adam; both exchange forms (overlapped buckets inside one CUDA graph / blocking all-reduce between two graphs)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('split', ['0', '1'])
def test_two_rank_step_matches_two_tower_oracle(lib, split):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs >= 2 GPUs')
    env = dict(os.environ, RCGAN_DP_SPLIT_GRAPH=split, RCGAN_DP_BUCKET_MB='1')
    port = 29600 + os.getpid() % 300 + int(split)
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
                        '--master-port', str(port), os.path.join(ROOT, 'tests', 'dp_check.py')], capture_output=True, text=True,
                       env=env, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert 'DP_CHECK world=2' in r.stdout, r.stdout[-2000:]
