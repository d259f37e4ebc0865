"""Multi-GPU plumbing: batch sharding and the per-lap safe-set exchange.

The hot path shards over independent controller instances: rank r owns the contiguous range
``shard_range(B, r, world)`` and no collective sits inside solve()/addPoint().  The only exchange is the
optional *pooled safe-set* mode: once per lap, after LMPC.addTrajectory (PredictiveControllers.py:418-445),
every rank all-gathers the laps its instances just finished so that all ranks can rank laps by LapTime
(PC.py:395) over the pooled set.  torch.distributed is plumbing here (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(batch, rank, world):
    """Contiguous instance range [lo, hi) of `rank`; sizes differ by at most one."""
    base, rem = divmod(int(batch), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_laps(xs, us, Tmax):
    """Lists of laps (x[T,6], u[T,2]) -> padded rows [n, Tmax, 8] and lengths [n] (int32)."""
    n = len(xs)
    rows = np.zeros((n, Tmax, 8))
    lens = np.zeros(n, np.int32)
    for i, (x, u) in enumerate(zip(xs, us)):
        T = x.shape[0]
        if T > Tmax:
            raise ValueError("lap longer than Tmax")
        rows[i, :T, 0:6] = x
        rows[i, :T, 6:8] = u
        lens[i] = T
    return rows, lens


def allgather_laps(rows, lens, group=None):
    """All-gather padded laps over the process group.

    rows: tensor [n_local, Tmax, 8] (CUDA for NCCL, CPU for gloo), lens: int32 tensor [n_local];
    n_local must be equal on all ranks (pad with zero-length laps).  Returns (rows_all [world*n_local, Tmax, 8],
    lens_all [world*n_local]) in rank order, i.e. global instance order under `shard_range`."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return rows, lens
    out_rows = torch.empty((world * rows.shape[0],) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
    out_lens = torch.empty(world * lens.shape[0], dtype=lens.dtype, device=lens.device)
    dist.all_gather_into_tensor(out_rows, rows.contiguous(), group=group)
    dist.all_gather_into_tensor(out_lens, lens.contiguous(), group=group)
    return out_rows, out_lens


def allgather_vec(v, group=None):
    """All-gather a 1-D tensor of equal length on every rank, in rank order."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return v
    out = torch.empty(world * v.shape[0], dtype=v.dtype, device=v.device)
    dist.all_gather_into_tensor(out, v.contiguous(), group=group)
    return out


def pooled_fastest(lens_all, k, valid_min=2):
    """Indices (into the gathered order) of the k fastest finished laps, ties to the lower global index —
    the stable counterpart of np.argsort(LapTime)[:k] (PC.py:395,402)."""
    lens_np = lens_all.cpu().numpy() if torch.is_tensor(lens_all) else np.asarray(lens_all)
    idx = np.where(lens_np >= valid_min)[0]
    order = idx[np.argsort(lens_np[idx], kind="stable")]
    return order[:k]


def max_over_ranks(value, device, group=None):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
