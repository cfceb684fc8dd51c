"""world_size-2 gloo tests (CPU) of the N>1 host plumbing: sweep sharding, max-over-ranks timing, aggregate
throughput and the observable gather used by the replica mode of bench.py (--gpus N)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world_size, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    import tevo_b200  # noqa: F401
    from timeevomps_jl_b200 import replicas as r
    res = {}
    res["world"] = r.world()
    res["shard"] = r.shard_sweep([0.0, 0.2, 0.4, 0.6, 0.8, 1.0, 1.2, 1.4], rank, world_size)
    res["shard5"] = r.shard_sweep(list(range(5)), rank, world_size)
    res["seed"] = r.trajectory_seed(rank, 10)
    ms_local = 100.0 if rank == 0 else 250.0
    res["max"] = r.max_over_ranks(ms_local, dist)
    res["agg"] = r.aggregate_steps_per_sec(ms_local, 2, dist)
    res["gather"] = r.gather_observables([rank + 0.5, rank * 2.0], dist)
    dist.barrier()
    q.put((rank, res))
    dist.destroy_process_group()


def test_replica_plumbing_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out[0]["world"] == (0, 2, 0) and out[1]["world"] == (1, 2, 1)
    assert out[0]["shard"] == [0.0, 0.2, 0.4, 0.6] and out[1]["shard"] == [0.8, 1.0, 1.2, 1.4]
    assert out[0]["shard5"] == [0, 1, 2] and out[1]["shard5"] == [3, 4]
    assert (out[0]["seed"], out[1]["seed"]) == (10, 11)
    for r in (0, 1):
        assert out[r]["max"] == 250.0                              # slowest rank defines the job time
        assert abs(out[r]["agg"] - 2 * 2 * 1e3 / 250.0) < 1e-12   # all ranks' steps / max time
        assert out[r]["gather"] == [[0.5, 0.0], [1.5, 2.0]]


def test_single_process_fallbacks():
    sys.path.insert(0, ROOT)
    import tevo_b200  # noqa: F401
    from timeevomps_jl_b200 import replicas as r
    assert r.max_over_ranks(3.0) == 3.0
    assert r.aggregate_steps_per_sec(500.0, 1) == 2.0
    assert r.gather_observables([1.0, 2.0]) == [[1.0, 2.0]]
    assert r.shard_sweep([1, 2, 3], 0, 1) == [1, 2, 3]
