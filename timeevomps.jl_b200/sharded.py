"""Chain-segment sharding of layer-parallel TEBD over the GPUs of one node (BASELINE config 5).

Rank r owns the contiguous sites [lo, hi).  It keeps a local B-form chain of its sites plus one ghost site (a copy
of the right neighbour's first site).  A gate whose two sites straddle the cut (hi-1 | hi) is applied by the left
owner: before the layer it receives the current B_hi from rank r+1, after the update it sends the new B_hi and the
new Schmidt values of the cut bond back - one boundary tensor each way per cut and layer, peer to peer
(torch.distributed send/recv: NCCL over NVLink on the GPUs, gloo in the CPU tests).  No other communication sits on
the data path; observables are gathered with one small all_gather.

The arithmetic of a segment lives behind the tiny `Segment` interface so that the exchange schedule is tested on
CPU with a NumPy segment and runs on the device with `DeviceSegment` (libtevo).  Layer-parallel TEBD is not the
reference's sequential algorithm (see tebd.py / DESIGN.md section 8)."""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import numpy as np


def segment_bounds(N: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous balanced partition of the sites."""
    return (rank * N) // world, ((rank + 1) * N) // world


def weighted_segment_bounds(weights: Sequence[float], world: int, rank: int) -> Tuple[int, int]:
    """Contiguous partition of the sites with (nearly) equal total weight per rank.  Used by the benchmark with
    weight_j ~ chi_j^3 (the cost of a two-site update next to site j): the first and last ~log_d(chi) sites of a
    saturated chain carry small bond dimensions, so the edge ranks take a few more sites than the interior ones.
    Every rank gets at least two sites."""
    N = len(weights)
    tot = float(sum(weights))
    cuts = [0]
    acc, j = 0.0, 0
    for r in range(1, world):
        target = tot * r / world
        while j < N and acc + weights[j] * 0.5 < target:
            acc += weights[j]
            j += 1
        j = max(j, cuts[-1] + 2)
        j = min(j, N - 2 * (world - r))
        cuts.append(j)
        acc = float(sum(weights[:j]))
    cuts.append(N)
    return cuts[rank], cuts[rank + 1]


def owner_of_site(N: int, world: int, j: int) -> int:
    for r in range(world):
        lo, hi = segment_bounds(N, world, r)
        if lo <= j < hi:
            return r
    raise ValueError(j)


class Segment:
    """Interface of a rank-local piece of a B-form chain.  Local site i = global site lo + i; local index nloc is the
    ghost site.  Tensors cross the interface as torch tensors living where the communication backend wants them."""

    def site_tensor(self, i):            # -> (torch tensor of float64 pairs, (chil, d, chir))
        raise NotImplementedError

    def set_site_tensor(self, i, t, shape):
        raise NotImplementedError

    def lam_tensor(self, i):             # Schmidt values on the bond left of local site i
        raise NotImplementedError

    def set_lam_tensor(self, i, t):
        raise NotImplementedError

    def apply_layer(self, local_gates, **kw):
        raise NotImplementedError

    def expect_site(self, op, i) -> float:
        raise NotImplementedError

    def sync(self):
        pass


class ShardedTEBD:
    def __init__(self, segment: Segment, N: int, dist, rank: int, world: int, device="cpu", bounds=None):
        self.seg, self.N, self.dist, self.rank, self.world, self.device = segment, N, dist, rank, world, device
        self.lo, self.hi = bounds if bounds is not None else segment_bounds(N, world, rank)
        self.nloc = self.hi - self.lo
        self.bytes_sent = 0
        self.exchange_s = 0.0     # host wall time spent in the boundary exchange (sends, receives, their syncs)

    # ---- point-to-point helpers ----
    def _exchange(self, send, recv_from):
        """One batched neighbour exchange: `send` = (dst, [tensors]) or None, `recv_from` = src rank or None.  Two
        rounds of batch_isend_irecv (sizes, then payload), so that a rank's send and receive are posted together:
        with blocking send-then-recv the transfers of a layer form a serial chain from the last rank to the first.
        Returns the list of received float64 tensors (same split as the sender's list)."""
        import torch
        dist = self.dist
        NH = 8
        ops, hdr_in = [], None
        if send is not None:
            dst, parts = send
            sizes = [int(p.numel()) for p in parts]
            hdr = torch.tensor([len(parts)] + sizes + [0] * (NH - 1 - len(sizes)), dtype=torch.int64, device=self.device)
            ops.append(dist.P2POp(dist.isend, hdr, dst))
        if recv_from is not None:
            hdr_in = torch.zeros(NH, dtype=torch.int64, device=self.device)
            ops.append(dist.P2POp(dist.irecv, hdr_in, recv_from))
        for r in dist.batch_isend_irecv(ops):
            r.wait()
        ops, buf, sizes_in = [], None, []
        if send is not None:
            dst, parts = send
            payload = torch.cat([p.reshape(-1).to(torch.float64) for p in parts]) if len(parts) > 1 else parts[0].reshape(-1)
            payload = payload.contiguous()
            ops.append(dist.P2POp(dist.isend, payload, dst))
            self.bytes_sent += payload.numel() * payload.element_size()
        if recv_from is not None:
            h = hdr_in.tolist()
            sizes_in = [int(x) for x in h[1:1 + int(h[0])]]
            buf = torch.empty(sum(sizes_in), dtype=torch.float64, device=self.device)
            ops.append(dist.P2POp(dist.irecv, buf, recv_from))
        for r in dist.batch_isend_irecv(ops):
            r.wait()
        if buf is None:
            return []
        out, o = [], 0
        for n in sizes_in:
            out.append(buf[o:o + n])
            o += n
        return out

    # ---- one Trotter layer ----
    def apply_layer(self, gates: Sequence, **kw):
        """gates: BondGate-like objects with global bond .b and matrix .G, all on disjoint bonds."""
        import time as _time
        import torch
        bonds = sorted(g.b for g in gates)
        if any(b2 - b1 < 2 for b1, b2 in zip(bonds, bonds[1:])):
            # e.g. a TEBD2Sweep layer (all bonds): the cut gate would read a ghost copy of a site that the right
            # neighbour updates in the same layer
            raise ValueError("ShardedTEBD.apply_layer: the gates of a layer must act on pairwise disjoint bonds "
                             "(even/odd layers of TEBD2 / TEBD4); TEBD2Sweep cannot be sharded")
        mine = [g for g in gates if self.lo <= g.b < self.hi]
        right_cut = self.hi - 1          # bond (hi-1 | hi), owned by this rank
        left_cut = self.lo - 1           # bond (lo-1 | lo), owned by rank-1
        has_right = self.rank < self.world - 1 and any(g.b == right_cut for g in gates)
        has_left = self.rank > 0 and any(g.b == left_cut for g in gates)
        self.seg.sync()
        _t0 = _time.perf_counter()
        # 1. boundary tensors travel left: my first site goes to the rank that updates the cut bond
        if has_left or has_right:
            send = None
            if has_left:
                t, shape = self.seg.site_tensor(0)
                send = (self.rank - 1, [torch.tensor(list(shape) + [0], dtype=torch.float64, device=self.device), t])
            got = self._exchange(send, self.rank + 1 if has_right else None)
            if has_right:
                shape = tuple(int(x) for x in got[0].tolist()[:3])   # (header padded to 32 bytes: payload stays aligned)
                self.seg.set_site_tensor(self.nloc, got[1], shape)
        self.exchange_s += _time.perf_counter() - _t0
        # 2. local updates (all independent)
        out = self.seg.apply_layer([(g.b - self.lo, g.G) for g in mine], **kw)
        self.seg.sync()
        _t0 = _time.perf_counter()
        # 3. updated boundary tensor and cut-bond Schmidt values travel back right
        if has_left or has_right:
            send = None
            if has_right:
                t, shape = self.seg.site_tensor(self.nloc)
                lt = self.seg.lam_tensor(self.nloc)
                send = (self.rank + 1, [torch.tensor(list(shape) + [0], dtype=torch.float64, device=self.device), t, lt])
            got = self._exchange(send, self.rank - 1 if has_left else None)
            if has_left:
                shape = tuple(int(x) for x in got[0].tolist()[:3])
                self.seg.set_site_tensor(0, got[1], shape)
                self.seg.set_lam_tensor(0, got[2])
        self.exchange_s += _time.perf_counter() - _t0
        return out

    def run(self, H_gates, dt, nsteps: int, time_evo_gates, alg, **kw):
        """The schedule of tebd! (bunched half steps, no callbacks) over the sharded chain."""
        Ustart, Us, Uend = time_evo_gates(dt, H_gates, alg)
        for U in Ustart:
            self.apply_layer(U, **kw)
        for _ in range(nsteps - 1 if len(Uend) else nsteps):
            for U in Us:
                self.apply_layer(U, **kw)
        for U in Uend:
            self.apply_layer(U, **kw)

    def gather_expect(self, op) -> np.ndarray:
        """Per-site expectation values of the whole chain on every rank (one all_gather of N doubles)."""
        import torch
        local = np.zeros(self.N)
        for i in range(self.nloc):
            local[self.lo + i] = self.seg.expect_site(op, i)
        t = torch.tensor(local, dtype=torch.float64, device=self.device)
        if self.world > 1:
            self.dist.all_reduce(t)
        return t.cpu().numpy()


class _DevView:
    """__cuda_array_interface__ wrapper so torch can view library-owned device memory (no copy)."""

    def __init__(self, ptr: int, ndoubles: int):
        self.__cuda_array_interface__ = {"shape": (ndoubles,), "typestr": "<f8", "data": (ptr, False), "version": 2}


class DeviceSegment(Segment):
    """A rank-local B-form chain held by libtevo (sites lo..hi-1 plus the ghost site)."""

    def __init__(self, global_bform, lo: int, hi: int, ghost: bool):
        """Copies the segment out of a full B-form chain that lives on this rank's GPU."""
        from . import _lib
        from .mps import MPS
        self._lib = _lib
        lib = _lib.load()
        g = global_bform
        self.ctx, self.d = g.ctx, g.d
        self.nloc = hi - lo
        nl = self.nloc + 1                  # always keep the ghost slot (unused on the last rank)
        self.psi = MPS(max(nl, 2), g.d, g.ctx)
        for i in range(nl):
            j = lo + i
            if j < g.N and (i < self.nloc or ghost):
                ptr, l, r = C.c_void_p(), C.c_int(), C.c_int()
                _lib.check(lib.tevo_mps_site_devptr(g.h, j, C.byref(ptr), C.byref(l), C.byref(r)), g.ctx.h)
                _lib.check(lib.tevo_mps_set_site_from_device(self.psi.h, i, l.value, r.value, ptr), g.ctx.h)
            else:
                break
        for i in range(max(nl, 2) + 1):
            j = lo + i
            if j <= g.N and i <= self.nloc + (1 if ghost else 0):
                ptr, n = C.c_void_p(), C.c_int()
                _lib.check(lib.tevo_mps_lambda_devptr(g.h, j, C.byref(ptr), C.byref(n)), g.ctx.h)
                _lib.check(lib.tevo_mps_set_lambda_from_device(self.psi.h, i, n.value, ptr), g.ctx.h)
            else:
                one = np.ones(1)
                _lib.check(lib.tevo_mps_set_lambda(self.psi.h, i, 1, one.ctypes.data_as(C.POINTER(C.c_double))),
                           g.ctx.h)
        _lib.check(lib.tevo_mps_mark_bform(self.psi.h, 1), g.ctx.h)
        self.ctx.synchronize()

    @classmethod
    def random(cls, N: int, chi: int, lo: int, hi: int, ghost: bool, seed: int = 0, d: int = 2):
        """The segment [lo, hi) (+ ghost site) of the random saturated benchmark chain (tevo_mps_random with the same
        seed on every rank, brought to B-form); the full chain is dropped again once the segment has been copied."""
        from .mps import MPS
        glob = MPS.random(N, chi, d=d, seed=seed).to_bform()
        seg = cls(glob, lo, hi, ghost)
        del glob
        return seg

    def sync(self):
        import torch
        self.ctx.synchronize()
        torch.cuda.synchronize()

    def site_tensor(self, i):
        import torch
        lib = self._lib.load()
        ptr, l, r = C.c_void_p(), C.c_int(), C.c_int()
        self._lib.check(lib.tevo_mps_site_devptr(self.psi.h, i, C.byref(ptr), C.byref(l), C.byref(r)), self.ctx.h)
        n = l.value * self.d * r.value * 2
        return torch.as_tensor(_DevView(ptr.value, n), device="cuda"), (l.value, self.d, r.value)

    def set_site_tensor(self, i, t, shape):
        import torch
        torch.cuda.synchronize()
        self._lib.check(self._lib.load().tevo_mps_set_site_from_device(self.psi.h, i, shape[0], shape[2],
                                                                       C.c_void_p(t.data_ptr())), self.ctx.h)
        self.ctx.synchronize()

    def lam_tensor(self, i):
        import torch
        ptr, n = C.c_void_p(), C.c_int()
        self._lib.check(self._lib.load().tevo_mps_lambda_devptr(self.psi.h, i, C.byref(ptr), C.byref(n)), self.ctx.h)
        return torch.as_tensor(_DevView(ptr.value, n.value), device="cuda")

    def set_lam_tensor(self, i, t):
        import torch
        torch.cuda.synchronize()
        self._lib.check(self._lib.load().tevo_mps_set_lambda_from_device(self.psi.h, i, t.numel(),
                                                                         C.c_void_p(t.data_ptr())), self.ctx.h)
        self.ctx.synchronize()

    def apply_layer(self, local_gates, **kw):
        from .bondop import BondGate
        from .tebd import apply_gates_parallel
        return apply_gates_parallel(self.psi, [BondGate(G, b) for (b, G) in local_gates], **kw)

    def expect_site(self, op, i) -> float:
        return float(self.psi.bform_expect([op], i)[0].real)
