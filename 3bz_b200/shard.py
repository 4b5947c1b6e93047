"""Host-side sharding of a batch of independent members over ranks / devices.

No collective is involved in the data path: every rank decodes its own members on its own GPU.
`tbz_partition` (C ABI) balances arbitrary member sizes; `rank_slice` is the contiguous split
used when members are equal-sized (the bench workloads).  torch.distributed is only used by the
callers for the barrier and the max-over-ranks of the device time.
"""
import ctypes as C


def rank_slice(n_total, rank, world):
    """Contiguous [lo, hi) of n_total members for `rank` of `world` (sizes differ by at most one)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    return n_total * rank // world, n_total * (rank + 1) // world


def partition(in_lens, world):
    """owner[i] in [0, world): longest-first greedy assignment through the C ABI (no GPU needed)."""
    from . import _ffi
    L = _ffi.lib()
    n = len(in_lens)
    lens = (C.c_uint64 * max(1, n))(*in_lens)
    owner = (C.c_int32 * max(1, n))()
    _ffi.check(L.tbz_partition(lens, n, world, owner))
    return list(owner[:n])


def reduce_max(values, dist=None):
    """max over ranks of a list of floats (device times); identity when not distributed."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return list(values)
    import torch
    t = torch.tensor(list(values), dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]
