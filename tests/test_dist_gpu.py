"""Multi-GPU correctness on the hardware (needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_dist_gpu.py -m gpu`): one
view sharded over the NCCL ranks in interleaved ray tiles + ONE all-gather equals the single-GPU render bit for bit,
with and without the importance pass.  The reference has no counterpart (evaluation is rank-0 only, training_loop.py:311-328)."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs >= 2 GPUs on one node')
def test_nccl_sharded_render_equals_single_gpu_render():
    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={n}', '--master-addr', '127.0.0.1', '--master-port', '29531',
           os.path.join(ROOT, 'tests', 'helpers', 'dist_nccl_check.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    print(r.stdout[-2000:])
    assert r.returncode == 0 and 'DIST_OK' in r.stdout, r.stderr[-3000:]
