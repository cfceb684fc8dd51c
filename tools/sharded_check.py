"""Multi-GPU equality check of the chain-segment sharded TEBD (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/sharded_check.py [N chi nsteps]
Every rank builds the same global random state, keeps its segment, evolves it with boundary exchange over NCCL and
the gathered per-site <Sz>, <Sx> are compared with a single-GPU layer-parallel run of the whole chain on rank 0."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import tevo_b200 as t
from timeevomps_jl_b200.sharded import DeviceSegment, ShardedTEBD, segment_bounds

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
chi = int(sys.argv[2]) if len(sys.argv) > 2 else 32
nsteps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = t.default_ctx()
glob = t.MPS.random(N, chi, seed=5).to_bform()
lo, hi = segment_bounds(N, world, rank)
seg = DeviceSegment(glob, lo, hi, ghost=(rank < world - 1))
sh = ShardedTEBD(seg, N, dist, rank, world, device="cuda")
H = t.gates(t.xxz_bondop(N, 1.0, 0.2))
kw = dict(maxdim=chi, mindim=min(chi, 4))
ctx.synchronize()
t0 = time.time()
sh.run(H, 0.05, nsteps, t.time_evo_gates, t.TEBD4(), **kw)
seg.sync()
dt = time.time() - t0
sz, sx = sh.gather_expect("Sz"), sh.gather_expect("Sx")
if rank == 0:
    ref = glob  # evolve the full chain on this GPU
    Ustart, Us, Uend = t.time_evo_gates(0.05, H, t.TEBD4())
    for U in Ustart:
        t.apply_gates_parallel(ref, U, **kw)
    for _ in range(nsteps - 1):
        for U in Us:
            t.apply_gates_parallel(ref, U, **kw)
    for U in Uend:
        t.apply_gates_parallel(ref, U, **kw)
    rz, rx = ref.bform_expect_all("Sz"), ref.bform_expect_all("Sx")
    print("world %d N %d chi %d steps %d: sharded run %.3f s, boundary bytes sent by rank 0: %d" % (world, N, chi, nsteps, dt, sh.bytes_sent))
    print("max |dSz| = %.3e   max |dSx| = %.3e   (sharded vs single-GPU layer-parallel)" % (np.abs(sz - rz).max(), np.abs(sx - rx).max()))
    assert np.abs(sz - rz).max() < 1e-10 and np.abs(sx - rx).max() < 1e-10
    print("SHARDED_CHECK_OK")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
