"""Host-side pieces of the multi-GPU (one process per GPU) path: session sharding and step-count agreement.

Sessions are independent (README.md:242 of the reference), so rank r trains on every world-th session of the
time-ordered session list; the ranks advance in lock step (one merged update per mini-batch), therefore they must
run the same number of mini-batches per epoch: the minimum over the ranks."""
import numpy as np


def shard_sessions(session_order, rank, world):
    """Sessions of `session_order` (gru4rec.py:585/593) handled by `rank`: positions rank, rank+world, ..."""
    return np.ascontiguousarray(np.asarray(session_order)[rank::world])


def common_steps(n_steps, dist):
    """Minimum step count over all ranks (torch.distributed, any backend)."""
    import torch
    dev = 'cuda' if dist.get_backend() == 'nccl' else 'cpu'
    t = torch.tensor([int(n_steps)], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return int(t.item())
