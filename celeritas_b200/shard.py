"""Event sharding and the end-of-run reduction for multi-GPU runs.

Events are independent (RNG streams are reseeded from the *global* event id,
/root/reference/src/celeritas/random/RngReseed.cu:38-41), so rank r simply owns a disjoint
block of global event ids; the only exchange is one sum-reduction of the per-detector
energy deposition and of the step counters at the end of a pass (NCCL on GPUs, gloo in the
CPU tests). The reference itself has no multi-device mode.
"""
import numpy as np


def shard_events(events_per_rank, rank):
    """Global event ids owned by `rank` (weak scaling: fixed work per rank)."""
    first = rank * events_per_rank
    return np.arange(first, first + events_per_rank, dtype=np.uint32)


def split_events(num_events, rank, world):
    """Strong-scaling split of a fixed event list: contiguous, sizes differ by at most one."""
    base, extra = divmod(num_events, world)
    first = rank * base + min(rank, extra)
    count = base + (1 if rank < extra else 0)
    return np.arange(first, first + count, dtype=np.uint32)


def reduce_tallies(calo, counts, dist=None, device='cpu'):
    """Sum per-detector energy deposition (f64) and integer counters over all ranks.

    `counts` = [num_steps, num_step_iterations, num_primaries]; returns numpy arrays.
    """
    import torch
    t_calo = torch.as_tensor(np.asarray(calo, dtype=np.float64)).to(device)
    t_counts = torch.as_tensor(np.asarray(counts, dtype=np.int64)).to(device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t_calo)
        dist.all_reduce(t_counts)
    return t_calo.cpu().numpy(), t_counts.cpu().numpy()
