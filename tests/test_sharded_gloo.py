"""world_size-2 (and 3) gloo tests of the chain-segment sharding schedule (timeevomps.jl_b200/sharded.py): the
boundary-tensor exchange around the cut bonds must reproduce the single-process layer-parallel TEBD exactly.
The per-segment arithmetic is a NumPy B-form segment (the oracle's update rule), so no GPU is needed."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _make_segment_class():
    from oracle import tevo_oracle as o
    from timeevomps_jl_b200.sharded import Segment

    class OracleSegment(Segment):
        def __init__(self, glob, lo, hi, ghost):
            self.nloc = hi - lo
            B = [glob.B[j] for j in range(lo, hi)]
            lam = [glob.lam[j] for j in range(lo, hi + 1)]
            if ghost:
                B.append(glob.B[hi])
                lam.append(glob.lam[hi + 1])
            self.st = o.BForm(B, lam)

        def site_tensor(self, i):
            a = np.asfortranarray(self.st.B[i])
            return torch.from_numpy(a.reshape(-1, order="F").view(np.float64).copy()), a.shape

        def set_site_tensor(self, i, t, shape):
            a = t.numpy().view(np.complex128).reshape(shape, order="F").copy()
            if i < len(self.st.B):
                self.st.B[i] = a
            else:
                self.st.B.append(a)
                self.st.lam.append(np.ones(shape[2]))

        def lam_tensor(self, i):
            return torch.from_numpy(self.st.lam[i].copy())

        def set_lam_tensor(self, i, t):
            self.st.lam[i] = t.numpy().copy()

        def apply_layer(self, local_gates, **kw):
            return [o.apply_gate_bform(self.st, o.BondGate(G, b), **kw) for (b, G) in local_gates]

        def expect_site(self, op, i):
            return self.st.expect_site(op, i).real

    return OracleSegment


def _worker(rank, world, port, N, kw, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import tevo_b200  # noqa: F401
    from oracle import tevo_oracle as o
    from timeevomps_jl_b200.sharded import ShardedTEBD, segment_bounds
    glob = o.to_bform(o.MPS.random(N, 5, seed=11))
    lo, hi = segment_bounds(N, world, rank)
    Seg = _make_segment_class()
    seg = Seg(glob, lo, hi, ghost=(rank < world - 1))
    sh = ShardedTEBD(seg, N, dist, rank, world)
    H = o.xxz_bondop(N, 0.9, 0.2).gates()
    sh.run(H, 0.05, 4, o.time_evo_gates, o.TEBD4(), **kw)
    sz = sh.gather_expect("Sz")
    dims = [seg.st.B[i].shape[2] for i in range(seg.nloc)]
    q.put((rank, sz, dims, sh.bytes_sent))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,kw", [(2, dict(maxdim=6)), (2, dict(cutoff=1e-8)), (3, dict(maxdim=4))])
def test_sharded_equals_single_process(world, kw):
    sys.path.insert(0, ROOT)
    from oracle import tevo_oracle as o
    N = 10
    ref = o.to_bform(o.MPS.random(N, 5, seed=11))
    o.tebd_parallel(ref, o.xxz_bondop(N, 0.9, 0.2), 0.05, 0.2, o.TEBD4(), **kw)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 1500) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, kw, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = dict((r, (sz, dims, nb)) for (r, sz, dims, nb) in (q.get(timeout=180) for _ in range(world)))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        np.testing.assert_allclose(out[r][0], ref.expect_all("Sz"), atol=1e-12)
    dims = sum((out[r][1] for r in range(world)), [])
    assert dims[:N - 1] == ref.bond_dims()
    assert sum(out[r][2] for r in range(world)) > 0     # boundary tensors did travel


def test_segment_bounds_cover_chain():
    sys.path.insert(0, ROOT)
    import tevo_b200  # noqa: F401
    from timeevomps_jl_b200.sharded import segment_bounds, owner_of_site
    for N, world in ((256, 8), (256, 4), (100, 3), (10, 2)):
        cover = []
        for r in range(world):
            lo, hi = segment_bounds(N, world, r)
            cover += list(range(lo, hi))
        assert cover == list(range(N))
        assert owner_of_site(N, world, 0) == 0 and owner_of_site(N, world, N - 1) == world - 1


def test_weighted_segment_bounds_cover_the_chain():
    """The benchmark's cost-weighted partition: contiguous, covering, at least two sites per rank, and balancing the
    estimated cost (the chain ends of a saturated MPS are cheap, so the edge ranks take more sites)."""
    import tevo_b200  # noqa: F401  (registers the package alias)
    from timeevomps_jl_b200.sharded import weighted_segment_bounds
    N, chi = 256, 1024
    w = [float(min(chi, 2 ** min(j + 1, N - j - 1, 40))) ** 3 for j in range(N)]
    for world in (1, 2, 3, 4, 8):
        b = [weighted_segment_bounds(w, world, r) for r in range(world)]
        assert b[0][0] == 0 and b[-1][1] == N
        assert all(b[r][1] == b[r + 1][0] for r in range(world - 1))
        assert all(hi - lo >= 2 for lo, hi in b)
        cost = [sum(w[lo:hi]) for lo, hi in b]
        assert max(cost) <= 1.08 * (sum(w) / world)
    b8 = [weighted_segment_bounds(w, 8, r) for r in range(8)]
    assert b8[0][1] - b8[0][0] > b8[3][1] - b8[3][0]            # edge ranks hold more sites than interior ranks
    assert [weighted_segment_bounds([1.0] * 4, 2, r) for r in range(2)] == [(0, 2), (2, 4)]
