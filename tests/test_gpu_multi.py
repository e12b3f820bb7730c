"""-m gpu: 2-GPU synchronous data-parallel step vs the oracle on the merged mini-batch (needs >= 2 devices)."""
import os
import subprocess
import sys
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_parity():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (run with gpurun --gpus 2)')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1', '--master-port', '29533',
           os.path.join(ROOT, 'tests', 'mg_worker.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, (out.stdout[-3000:], out.stderr[-3000:])
    assert out.stdout.count('multi-gpu parity ok') == 2
