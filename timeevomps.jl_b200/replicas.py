"""Multi-GPU host plumbing for the paths that do not need a data-path collective: independent trajectories /
parameter sweeps (BASELINE config 4: one h value per GPU; two-site TDVP is sequential in the bond index, so it
scales by replicas only) and the replica mode of the TEBD benchmark.  One process per GPU, torch.distributed for
the barrier and the max-over-ranks timing only."""
from __future__ import annotations

import os
from typing import List, Sequence


def world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_sweep(values: Sequence, rank: int, world_size: int) -> List:
    """Contiguous block partition of a parameter sweep (e.g. the 8 field values of BASELINE config 4)."""
    n = len(values)
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return list(values[lo:hi])


def trajectory_seed(rank: int, base_seed: int = 0) -> int:
    return base_seed + rank


def max_over_ranks(x: float, dist=None, device="cpu") -> float:
    """Device-timed durations are combined as the max over ranks (never by wall clock)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    import torch
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_steps_per_sec(ms_local: float, steps_per_rank: int, dist=None, device="cpu") -> float:
    """Whole-job throughput: all ranks' steps divided by the slowest rank's time."""
    ws = 1 if (dist is None or not dist.is_initialized()) else dist.get_world_size()
    ms = max_over_ranks(ms_local, dist, device)
    return ws * steps_per_rank * 1e3 / ms


def gather_observables(local: Sequence[float], dist=None, device="cpu") -> List[List[float]]:
    """all_gather of per-rank observable vectors (e.g. per-site <Sz> of every trajectory) for the host-side report."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [list(local)]
    import torch
    t = torch.tensor(list(local), dtype=torch.float64, device=device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [o.tolist() for o in out]
