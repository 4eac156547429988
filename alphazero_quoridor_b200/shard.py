"""Sharding of independent games over the GPUs of one box.

Games never interact, so rank r simply owns a contiguous block of GLOBAL game indices and there is no
collective on the hot path.  Every random stream (rollouts, move sampling) is keyed by the global game index,
so a given set of games produces the same results on 1, 2, 4 or 8 GPUs.  The only communication is the
reduction of benchmark scalars (max of the elapsed time, sum of the work) -- `reduce_stats`.
"""
import os

import torch
import torch.distributed as dist


def world():
    """(rank, world_size, local_rank) from the torchrun environment (1 process per GPU)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_range(n_total, rank, world_size):
    """Contiguous block [lo, hi) of global game indices owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(int(n_total), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def owner_of(game_index, n_total, world_size):
    base, rem = divmod(int(n_total), int(world_size))
    cut = rem * (base + 1)
    if game_index < cut:
        return game_index // (base + 1)
    return rem + (game_index - cut) // max(base, 1)


def reduce_stats(elapsed_ms, work, device=None):
    """(max over ranks of elapsed_ms, sum over ranks of each entry of `work`).  Works on gloo (CPU) and nccl."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(elapsed_ms), [float(w) for w in work]
    t = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=device)
    w = torch.tensor([float(x) for x in work], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(w, op=dist.ReduceOp.SUM)
    return t.item(), w.tolist()


def gather_rows(local_rows, n_total, device=None):
    """All-gather per-rank result rows (e.g. chosen moves keyed by global game index) to every rank --
    a convenience for tests and result collection, NOT used on the hot path."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_rows
    ws = dist.get_world_size()
    sizes = [shard_range(n_total, r, ws) for r in range(ws)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((mx,) + tuple(local_rows.shape[1:]), dtype=local_rows.dtype, device=local_rows.device)
    pad[:local_rows.shape[0]] = local_rows
    outs = [torch.zeros_like(pad) for _ in range(ws)]
    dist.all_gather(outs, pad)
    return torch.cat([o[:hi - lo] for o, (lo, hi) in zip(outs, sizes)], 0)
