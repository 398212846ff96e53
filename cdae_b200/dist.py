"""Host-side logic of the data-parallel path: which users a rank trains, and
how a `torch.distributed` process group hands the NCCL unique id to the engine.

The rule is the one `build_plan()` in csrc/api.cu implements: the users [0, U) are cut into
GLOBAL minibatches of `batch_users` consecutive ids; rank r of `world` trains the contiguous
slice  [lo + n*r/world, lo + n*(r+1)/world)  of the minibatch [lo, lo+n).  Every rank therefore
sees every minibatch (one all-reduce each, same count on every rank — no rank can run ahead or
hang), the slices partition the minibatch, and the set of per-user gradients that are summed
does not depend on `world`.
"""
import numpy as np


def minibatch_slices(U, batch_users, rank, world):
    """[(a, b)] — this rank's user range in every global minibatch (may be empty: a == b)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside [0,%d)" % (rank, world))
    B = int(batch_users) if batch_users > 0 else 8192
    out = []
    for lo in range(0, int(U), B):
        n = min(B, U - lo)
        out.append((lo + (n * rank) // world, lo + (n * (rank + 1)) // world))
    return out


def owned_users(U, batch_users, rank, world):
    """Sorted global ids of the users rank `rank` trains (and whose Wu / Uu rows it owns)."""
    sl = minibatch_slices(U, batch_users, rank, world)
    return np.concatenate([np.arange(a, b, dtype=np.int64) for a, b in sl] +
                          [np.zeros(0, np.int64)])


def init_process_group_engine(model, device_index=None):
    """Join `model` (a reset() CDAE) to the current torch.distributed group: rank 0 creates the
    NCCL unique id, everyone receives it through the existing group (any backend), and the
    engine builds its own communicator on its own stream."""
    import torch.distributed as dist
    from .model import CDAE
    rank, world = dist.get_rank(), dist.get_world_size()
    if world == 1:
        return
    box = [CDAE.dist_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    model.dist_init(rank, world, box[0])
