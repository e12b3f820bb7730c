"""-m gpu: 2-GPU synchronous data-parallel step vs the oracle on the merged mini-batch (needs >= 2 devices)."""
import os
import subprocess
import sys
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world, port):
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world), '--master-addr', '127.0.0.1', '--master-port', str(port),
           os.path.join(ROOT, 'tests', 'mg_worker.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, (out.stdout[-3000:], out.stderr[-3000:])
    return out.stdout


def test_two_gpu_parity():
    """row-sharded in-kernel exchange (3 cases) + replicated NCCL path (2 cases) vs the oracle on the merged mini-batch"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (run with gpurun --gpus 2)')
    out = _run(2, 29533)
    assert out.count('multi-gpu parity ok') == 5, out[-2000:]
    assert out.count('sharded') >= 3


def test_all_gpu_parity():
    """the same on every GPU of the box (8 on an HGX node)"""
    import torch
    n = torch.cuda.device_count()
    if n < 4:
        pytest.skip('needs >= 4 GPUs')
    out = _run(n, 29534)
    assert out.count('multi-gpu parity ok') >= 4, out[-2000:]
