"""Host-side pieces of the multi-GPU (one process per GPU) path: session sharding and step-count agreement.

Sessions are independent (README.md:242 of the reference), so rank r trains on every world-th session of the
time-ordered session list; the ranks advance in lock step (one merged update per mini-batch), therefore they must
run the same number of mini-batches per epoch: the minimum over the ranks."""
import numpy as np


def shard_sessions(session_order, rank, world):
    """Sessions of `session_order` (gru4rec.py:585/593) handled by `rank`: positions rank, rank+world, ..."""
    return np.ascontiguousarray(np.asarray(session_order)[rank::world])


def env_world():
    """(world_size, rank, local_rank) a launcher such as torchrun put into the environment, or (1, 0, 0)."""
    import os
    try:
        world = int(os.environ.get('WORLD_SIZE', '1'))
        rank = int(os.environ.get('RANK', '0'))
        local = int(os.environ.get('LOCAL_RANK', str(rank)))
    except ValueError:
        return 1, 0, 0
    return (world, rank, local) if world > 1 else (1, 0, 0)


def init_from_env(backend=None):
    """Join the torch.distributed job the launcher described (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*): one process per
    GPU, NCCL unless `backend` (or G4R_DIST_BACKEND) says otherwise.  Returns (world_size, rank); (1, 0) without a launcher.
    Idempotent: an already initialised process group is left alone."""
    import os
    world, rank, local = env_world()
    if world == 1:
        return 1, 0
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        backend = backend or os.environ.get('G4R_DIST_BACKEND', 'nccl')
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        if torch.cuda.is_available():
            torch.cuda.set_device(local % max(torch.cuda.device_count(), 1))
        if backend == 'nccl':
            dist.init_process_group('nccl', device_id=torch.device('cuda', torch.cuda.current_device()))
        else:
            dist.init_process_group(backend)
    return dist.get_world_size(), dist.get_rank()


def allreduce_sum(values, dist):
    """Element-wise sum over all ranks of a vector of float64 (torch.distributed, any backend); every rank gets the result."""
    import torch
    dev = 'cuda' if dist.get_backend() == 'nccl' else 'cpu'
    t = torch.tensor(np.asarray(values, dtype=np.float64).reshape(-1), dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def shard_eval_sessions(n_sessions, rank, world):
    """Test sessions `rank` scores in a multi-process evaluate_gpu: every world-th session of the sorted session list.
    Sessions are independent in evaluation too (every lane carries its own hidden state, evaluation.py:90-139), so the
    shards need no exchange: the per-cut-off hit / reciprocal-rank sums and the event count are summed over the ranks."""
    return np.arange(int(rank), int(n_sessions), int(world), dtype=np.int64)


def common_steps(n_steps, dist):
    """Minimum step count over all ranks (torch.distributed, any backend)."""
    import torch
    dev = 'cuda' if dist.get_backend() == 'nccl' else 'cpu'
    t = torch.tensor([int(n_steps)], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return int(t.item())


# ---- row-sharded item tables (csrc/g4r_shard.cuh): host restatement of the ownership map and of the merged owner plan ----
def owner_of(item, world):
    """rank that holds row `item` of Wy / By / Wx0 (and of their optimizer state)"""
    return int(item) % int(world)


def local_row(item, world):
    """row index inside the owner's shard"""
    return int(item) // int(world)


def shard_rows(n_items, world, rank):
    """number of rows rank `rank` owns"""
    return (int(n_items) - int(rank) + int(world) - 1) // int(world)


def sort_columns_owner_major(items, n_items, world):
    """order in which a rank processes the score columns of one mini-batch in the sharded layout: by (owner, item, position);
    returns (keys, positions) with key = owner * n_items + item (what k_plan writes to pKey)"""
    items = np.asarray(items, dtype=np.int64)
    keys = (items % world) * n_items + items
    order = np.lexsort((np.arange(len(items)), keys))
    return keys[order], order


def merged_owner_plan(keys_per_rank, n_items, world, me):
    """entries (rank, sorted column index) of all ranks whose item is owned by `me`, in (item, rank, position) order -- the
    list k_mgs_plan builds on the device and the apply CTAs walk; the gradient row of entry (r, j) is slot [r][j] of the inbox"""
    ent = []
    for r, keys in enumerate(keys_per_rank):
        keys = np.asarray(keys)
        lo, hi = np.searchsorted(keys, me * n_items, 'left'), np.searchsorted(keys, (me + 1) * n_items, 'left')
        ent += [(int(keys[j]) - me * n_items, r, j) for j in range(lo, hi)]
    ent.sort()
    return ent
