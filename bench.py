#!/usr/bin/env python
"""bench.py - TEBD4 time-steps/s of the two-site-update hot path at N=256, chi=1024 (BASELINE.json configs[4] on
one B200; the metric is quoted on that configuration and it fits one GPU: 8 GiB of MPS).

One "step" = one bulk TEBD4 time step of tebd! (src/tebd.jl:138-146): the 10 gate layers of `Us`
(5 even + 5 odd = 1275 sequential two-site updates at N=256), applied through the public host API
(apply_gates -> tevo_apply_layer) to a device-resident random saturated MPS with mindim=maxdim=chi.

    python bench.py [--gpus N --steps K --warmup W] [--impl reference]

N>1 (torchrun): one independent trajectory per GPU (different seed), no data-path collective; `value` is the
aggregate over ranks.  `--impl reference` times the CPU restatement of the reference path (oracle/) on the host
cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "tebd4_time_steps_per_sec"
UNIT = "steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nsites", type=int, default=256)
    ap.add_argument("--chi", type=int, default=1024)
    ap.add_argument("--dt", type=float, default=0.05)
    ap.add_argument("--model", default="xxz", choices=["xxz", "tfim"])
    ap.add_argument("--cpu-sample-gates", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-1thread", action="store_true",
                    help="also time one update with the BLAS pool limited to one thread (adds ~1 min at chi=1024)")
    ap.add_argument("--mode", default="auto", choices=["auto", "sequential", "parallel"],
                    help="sequential: the reference's algorithm (centre moves between gates; default at 1 GPU); "
                         "parallel: layer-parallel (Hastings) TEBD, chain-segment sharded over the ranks with NCCL "
                         "boundary exchange (default at >1 GPU, strong scaling of one N-site chain)")
    ap.add_argument("--workload", default="tebd", choices=["tebd", "tdvp"],
                    help="tebd: the headline TEBD4 step (default); tdvp: two-site TDVP step (BASELINE config 4 shape)")
    return ap.parse_args()


def gates_per_step(N):
    n_odd = len(range(0, N - 1, 2))   # 0-based even indices = reference's odd bonds
    n_even = len(range(1, N - 1, 2))
    return 5 * n_odd + 5 * n_even


def algorithmic_flops_per_gate(chi, d=2):
    """SURVEY.md 8(d) conservative numerator: F_theta + F_gate + F_split (complex MAC = 8 flops)."""
    return 8 * d * d * chi ** 3 + 8 * d ** 4 * chi ** 2 + 8 * (d * chi) * (d * chi) * chi


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING recipe)."""

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def model_bondop(t, name, N):
    return t.xxz_bondop(N, 1.0) if name == "xxz" else t.tfi_bondop(N, 1.0, 0.5)


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (restated reference path) on the host cores, bounded sample
# ---------------------------------------------------------------------------------------------------------
class CpuArm:
    """The oracle (restated reference path, oracle/tevo_oracle.py) on the host cores.  The state is a window of the
    benchmark chain long enough for the bond dimension to saturate at chi in the middle; each sample applies
    interior two-site updates exactly as oracle.apply_gate = src/bondgate.jl:46-53 (QR centre move + theta GEMM +
    gate + LAPACK SVD split + truncation) and is extrapolated by the update count of one TEBD4 step."""

    def __init__(self, args):
        from oracle import tevo_oracle as o
        self.o = o
        self.args = args
        chi = args.chi
        self.nw = min(2 * int(np.ceil(np.log2(chi))) + 10, args.nsites)
        self.psi = o.MPS.random(self.nw, chi, seed=0)
        H = (o.xxz_bondop(self.nw, 1.0) if args.model == "xxz" else o.tfi_bondop(self.nw, 1.0, 0.5)).gates()
        tau1 = 1 / (4 - 4 ** (1 / 3)) * args.dt
        self.lo = self.nw // 2 - 3
        self.gates = {b: o.gate_exp(H[b], -1j * tau1) for b in range(self.lo, self.lo + 6)}
        self.next = self.lo
        self.psi.orthogonalize(self.lo)

    def sample(self, ngates):
        o, chi = self.o, self.args.chi
        t0 = time.perf_counter()
        for _ in range(ngates):
            b = self.next
            o.apply_gate(self.psi, self.gates[b], maxdim=chi, mindim=chi)
            self.next = b + 2 if b + 2 < self.lo + 6 else self.lo
        dt = time.perf_counter() - t0
        per_gate = dt / ngates
        sps = 1.0 / (per_gate * gates_per_step(self.args.nsites))
        try:
            import threadpoolctl
            cores = max([p.get("num_threads", 1) for p in threadpoolctl.threadpool_info()] or [1])
        except Exception:
            cores = os.cpu_count() or 1
        return {"value": sps, "unit": UNIT, "cores": int(cores), "kind": "port",
                "sample": "%d interior two-site updates (QR centre move + theta + gate + LAPACK SVD split) at chi=%d on "
                          "a %d-site window, %.2f s/update, extrapolated x%d updates per TEBD4 step"
                          % (ngates, chi, self.nw, per_gate, gates_per_step(self.args.nsites))}


def _cpu_arm_one_thread(self):
    """One more interior update with the BLAS pool limited to a single thread (bounded: one update)."""
    try:
        import threadpoolctl
    except Exception:
        return {"value_1thread": None, "sample_1thread": "threadpoolctl not available"}
    with threadpoolctl.threadpool_limits(limits=1):
        one = self.sample(1)
    return {"value_1thread": one["value"], "sample_1thread": "1 interior update, BLAS limited to 1 thread"}


CpuArm.sample_one_thread = _cpu_arm_one_thread


def workload_string(args):
    return ("TEBD4 time step (10 layers, %d sequential two-site updates), %s S=1/2 chain N=%d chi=%d d=2, random "
            "saturated canonical MPS, mindim=maxdim=chi, dt=%g"
            % (gates_per_step(args.nsites), args.model, args.nsites, args.chi, args.dt))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    arm = CpuArm(args)
    vals = []
    cb = None
    for _ in range(args.warmup + args.steps):
        cb = arm.sample(1)
        vals.append(cb["value"])
    timed = vals[args.warmup:] or vals[-1:]
    v = float(1.0 / np.mean([1.0 / x for x in timed]))
    cb["value"] = v
    cb["sample"] = "each step = " + cb["sample"]
    line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "c128",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_string(args), "mode": "sequential (reference-faithful centre moves)",
                       "parallelism": "host CPU, all BLAS threads"},
            "cpu_baseline": cb,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_sharded(args):
    """TEBD4 steps/s of ONE N-site chain sharded by contiguous segments over the ranks (BASELINE config 5):
    layer-parallel (Hastings, B-form) updates, one boundary tensor exchanged each way per cut bond and layer over
    NCCL P2P.  Strong scaling: the chain is fixed, the ranks split it.  Not the reference's sequential algorithm:
    identical without truncation, O(discarded weight) apart with it."""
    import torch
    import tevo_b200 as t
    from timeevomps_jl_b200 import replicas
    from timeevomps_jl_b200.sharded import DeviceSegment, ShardedTEBD, segment_bounds
    rank, world, local = replicas.world()
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = t.default_ctx()
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
    N, chi = args.nsites, args.chi
    glob = t.MPS.random(N, chi, seed=0).to_bform()      # same state on every rank; each keeps its segment
    lo, hi = segment_bounds(N, world, rank)
    seg = DeviceSegment(glob, lo, hi, ghost=(rank < world - 1))
    del glob
    sh = ShardedTEBD(seg, N, dist, rank, world, device="cuda")
    Ustart, Us, Uend = t.time_evo_gates(args.dt, t.gates(model_bondop(t, args.model, N)), t.TEBD4())
    kw = dict(maxdim=chi, mindim=chi)

    def step():
        for U in Us:
            sh.apply_layer(U, **kw)

    def barrier():
        seg.sync()
        if dist is not None:
            dist.barrier()

    for U in Ustart:
        sh.apply_layer(U, **kw)
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t.profile_enable(True)
    t.profile_read(reset=True)
    n0 = ctx.launch_count()
    b0 = sh.bytes_sent
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    wall = time.perf_counter() - t0
    ms = max(e0.elapsed_time(e1), 0.0)
    ms = replicas.max_over_ranks(ms, dist, device="cuda")
    launches = ctx.launch_count() - n0
    prof = t.profile_read(reset=True)
    t.profile_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms / args.steps
    value = 1e3 / ms_per_step
    # end to end: segment uploaded from pinned host memory, one step, segment downloaded - every step
    e2e = None
    if not args.no_e2e:
        import ctypes as C
        from timeevomps_jl_b200._lib import load as _load, check as _check
        nl = seg.nloc + (1 if rank < world - 1 else 0)
        host = [seg.psi.site(i) for i in range(nl)]
        lams = [seg.psi.schmidt_values(i) for i in range(nl + 1)]
        pinned = []
        for a in host:
            buf = torch.empty(a.size * 2, dtype=torch.float64, pin_memory=True)
            v = buf.numpy().view(np.complex128).reshape(a.shape, order="F")
            v[...] = a
            pinned.append(v)
        nbytes = sum(a.nbytes for a in pinned) + sum(l.nbytes for l in lams)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            for i, a in enumerate(pinned):
                seg.psi.set_site(i, a)
            for i, l in enumerate(lams):
                _check(_load().tevo_mps_set_lambda(seg.psi.h, i, len(l), l.ctypes.data_as(C.POINTER(C.c_double))), ctx.h)
            _check(_load().tevo_mps_mark_bform(seg.psi.h, 1), ctx.h)
            step()
            for i in range(nl):
                out = seg.psi.site(i)
                if out.shape == pinned[i].shape:
                    pinned[i][...] = out
            lams = [seg.psi.schmidt_values(i) for i in range(nl + 1)]
        barrier()
        dt_e2e = replicas.max_over_ranks(time.perf_counter() - t0, dist, device="cuda")
        e2e = {"value": args.steps / dt_e2e, "unit": UNIT, "h2d_bytes_per_step": int(nbytes),
               "d2h_bytes_per_step": int(nbytes)}
    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    total_prof = sum(v["ms"] for v in prof.values()) or 1.0
    roofline = None
    if "tridiag_panel" in prof:
        d = 2
        dims = [min(d ** min(j, N - j, 40), chi) for j in range(N + 1)]
        sizes = [d * dims[g.b + 2] for U in Us for g in U if lo <= g.b < hi]    # rho = theta^+ theta is n x n
        bytes_alg = sum(16.0 * s ** 3 / 3.0 for s in sizes) * args.steps
        tk = prof["tridiag_panel"]
        ach = bytes_alg / (tk["ms"] * 1e-3) / 1e9
        roofline = {"kernel": "tridiag_panel_kernel (rank 0)", "bound": "hbm", "achieved": ach, "peak": hbm_peak,
                    "unit": "GB/s", "frac": ach / hbm_peak, "traffic": None, "launches": tk["count"],
                    "avg_launch_ms": tk["ms"] / tk["count"], "share_of_step": tk["ms"] / total_prof}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "c128",
        "data": "synthetic",
        "config": {"workload": workload_string(args),
                   "mode": "layer-parallel (Hastings B-form) TEBD: NOT the sequential reference algorithm; equal to it "
                           "without truncation, O(discarded weight) apart with it",
                   "parallelism": ("1 GPU" if world == 1 else
                                   "chain split into %d contiguous segments, one per GPU; boundary tensor + Schmidt "
                                   "values exchanged per cut bond and layer by NCCL send/recv" % world),
                   "l2_policy": "inputs larger than L2 (segment of an 8 GiB MPS streamed per step)"},
        "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e, "roofline": roofline,
        "boundary_bytes_sent_per_step_rank0": int((sh.bytes_sent - b0) / max(args.steps, 1)),
        "wall_ms_per_step": wall * 1e3 / args.steps,
        "phases_ms": {k: round(v["ms"], 3) for k, v in prof.items()},
    }
    print(json.dumps(line))


def run_tdvp(args):
    """Secondary workload: one two-site TDVP step (right + left sweep, src/tdvp.jl:54-95) of the XXZ chain at the
    BASELINE config-4 shape (default N=64 via --nsites, chi=1024, krylovdim=30, exp_tol=1e-12); one independent
    h value per GPU (replicas).  Reported as an extra JSON line for profiles/, not the headline metric."""
    import ctypes as C
    import torch
    import tevo_b200 as t
    from timeevomps_jl_b200 import replicas
    from timeevomps_jl_b200._lib import load as _load
    rank, world, local = replicas.world()
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist = None
    ctx = t.default_ctx()
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
    N, chi = args.nsites, args.chi
    hz = replicas.shard_sweep(list(np.linspace(0.0, 1.4, max(world, 1))), rank, world)[0]
    psi = t.MPS.random(N, chi, seed=replicas.trajectory_seed(rank))
    H = t.xxz_mpo(N, 1.0, float(hz))
    kw = dict(maxdim=chi, mindim=chi, exp_tol=1e-12, krylovdim=30)
    for _ in range(args.warmup):
        t.tdvp(psi, H, args.dt, args.dt, **kw)
    ctx.synchronize()
    if dist is not None:
        dist.barrier()
    out0 = (C.c_double * 12)()
    _load().tevo_debug_counters.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
    _load().tevo_debug_counters(ctx.h, out0, 12)
    t.profile_enable(True)
    t.profile_read(reset=True)
    st = {}
    n0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        t.tdvp(psi, H, args.dt, args.dt, stats=st, **kw)
    e1.record(stream)
    ctx.synchronize()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    ms = replicas.max_over_ranks(e0.elapsed_time(e1), dist, device="cuda")
    prof = t.profile_read(reset=True)
    out1 = (C.c_double * 12)()
    _load().tevo_debug_counters(ctx.h, out1, 12)
    if rank != 0:
        return
    fp64 = {}
    try:
        fp64 = json.load(open(os.path.join(ROOT, "profiles", "r01_fp64_peaks.json")))
    except Exception:
        pass
    zpeak = float(fp64.get("zgemm_tflops_sustained", 36.8))
    heff_flops = out1[11] - out0[11]
    hk = prof.get("heff_apply", {"ms": 0.0, "count": 0})
    ach = heff_flops / (hk["ms"] * 1e-3) / 1e12 if hk["ms"] else None
    line = {"metric": "tdvp_time_steps_per_sec", "value": world * args.steps * 1e3 / ms, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "c128", "data": "synthetic",
            "config": {"workload": "two-site TDVP step (right+left sweep), XXZ S=1/2 chain N=%d chi=%d, w=5 MPO, dt=%g, "
                                   "krylovdim=30, exp_tol=1e-12, mindim=maxdim=chi" % (N, chi, args.dt),
                       "parallelism": "1 GPU" if world == 1 else "%d independent field values (replicas)" % world},
            "gpu_launches": int(ctx.launch_count() - n0), "matvecs": st,
            "roofline": {"kernel": "zgemm_dmma_kernel (H_eff legs) + mid_contract_kernel", "bound": "tensor",
                         "achieved": ach, "peak": zpeak, "unit": "TFLOP/s", "frac": (ach / zpeak) if ach else None,
                         "traffic": None, "peak_source": "measured cuBLAS ZGEMM 4096^3 (profiles/r01_fp64_peaks.json)",
                         "share_of_step": hk["ms"] / (sum(v["ms"] for v in prof.values()) or 1.0)},
            "phases_ms": {k: round(v["ms"], 3) for k, v in prof.items()}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.workload == "tdvp":
        run_tdvp(args)
        return
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    if args.mode == "parallel" or (args.mode == "auto" and world_env > 1):
        run_sharded(args)
        return
    import torch
    import tevo_b200 as t
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = t.default_ctx()
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
    N, chi = args.nsites, args.chi

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    from timeevomps_jl_b200 import replicas
    psi = t.MPS.random(N, chi, seed=replicas.trajectory_seed(rank))
    ctx.synchronize()
    Ustart, Us, Uend = t.time_evo_gates(args.dt, t.gates(model_bondop(t, args.model, N)), t.TEBD4())
    kw = dict(maxdim=chi, mindim=chi)
    state = {"rev": False}

    def layer(U):
        t.apply_gates(psi, U, reversed=state["rev"], **kw)
        state["rev"] = not state["rev"]

    def step():
        for U in Us:
            layer(U)

    for U in Ustart:
        layer(U)
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t.profile_enable(True)
    t.profile_read(reset=True)
    n0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    launches = ctx.launch_count() - n0
    ms = e0.elapsed_time(e1)
    prof = t.profile_read(reset=True)
    t.profile_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    ms = replicas.max_over_ranks(ms, dist if world > 1 else None, device="cuda")
    ms_per_step = ms / args.steps
    value = world * 1e3 / ms_per_step

    # ---- end-to-end through the public API with host buffers (upload + step + download every step) ----
    e2e = None
    if not args.no_e2e:
        host = psi.tensors()
        pinned = []
        for a in host:
            buf = torch.empty(a.size * 2, dtype=torch.float64, pin_memory=True)
            v = buf.numpy().view(np.complex128).reshape(a.shape, order="F")
            v[...] = a
            pinned.append(v)
        nbytes = sum(a.nbytes for a in pinned)
        llim, rlim = psi.limits()
        del psi
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            psi = t.MPS.from_tensors(pinned, llim=llim, rlim=rlim)
            step()
            out = psi.tensors()
            llim, rlim = psi.limits()
            for dst, src in zip(pinned, out):
                if dst.shape == src.shape:
                    dst[...] = src
            del out
        barrier()
        dt_e2e = time.perf_counter() - t0
        dt_e2e = replicas.max_over_ranks(dt_e2e, dist if world > 1 else None, device="cuda")
        e2e = {"value": world * args.steps / dt_e2e, "unit": UNIT, "h2d_bytes_per_step": int(nbytes),
               "d2h_bytes_per_step": int(nbytes)}

    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    fp64 = {}
    try:
        fp64 = json.load(open(os.path.join(ROOT, "profiles", "r01_fp64_peaks.json")))
    except Exception:
        pass
    zgemm_peak = float(fp64.get("zgemm_tflops_sustained", 36.8))
    total_prof = sum(v["ms"] for v in prof.values()) or 1.0
    dom = max(prof.items(), key=lambda kv: kv[1]["ms"])[0] if prof else None
    roofline = None
    if "tridiag_panel" in prof:
        # dominant kernel: Householder tridiagonalisation panel (HEMV-bound).  Algorithmic bytes of one launch =
        # one read of the trailing matrix per column = 64 columns x (n - p0)^2 x 16 B; summed over the launches of
        # the timed region analytically: every factorisation of size n reads 16 n^3 / 3 bytes.
        d = 2
        sizes = []
        dims = [min(d ** min(j, N - j, 40), chi) for j in range(N + 1)]
        for U in Us:
            for g in U:
                b = g.b
                m, n = dims[b] * d, d * dims[b + 2]
                sizes.append(m)                                  # split: rho is m x m (left) / n x n (right)
                sizes.append(min(dims[b + 1] * d, dims[b + 2]))  # centre move: Gram matrix of the small side
        bytes_alg = sum(16.0 * s ** 3 / 3.0 for s in sizes) * args.steps
        tk = prof["tridiag_panel"]
        ach = bytes_alg / (tk["ms"] * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_panel_traffic.json")))["dram_bytes_per_launch"]
        except Exception:
            pass
        roofline = {"kernel": "tridiag_panel_kernel", "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                    "frac": ach / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": bytes_alg / max(tk["count"], 1),
                    "note": "latency-bound panel kernel: the trailing matrix is L2-resident within a launch, so DRAM "
                            "traffic (ncu, profiles/r01_panel_traffic.json) is ~1 matrix per launch although 64 "
                            "matrix-vector products read it algorithmically",
                    "launches": tk["count"], "avg_launch_ms": tk["ms"] / tk["count"],
                    "share_of_step": tk["ms"] / total_prof}
    flops_step = algorithmic_flops_per_gate(chi) * gates_per_step(N)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "c128",
        "data": "synthetic",
        "config": {"workload": workload_string(args),
                   "mode": "sequential (reference-faithful centre moves)",
                   "parallelism": "1 GPU" if world == 1 else "%d independent trajectories (replicas), no collective" % world,
                   "l2_policy": "inputs larger than L2 (8 GiB of MPS streamed per step)"},
        "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e, "roofline": roofline,
        "fp64_tensor": {"algorithmic_tflops": flops_step / (ms_per_step * 1e-3) / 1e12, "peak_tflops": zgemm_peak,
                        "frac": flops_step / (ms_per_step * 1e-3) / 1e12 / zgemm_peak,
                        "note": "SURVEY 8(d) conservative numerator (theta + gate + factor GEMMs) over measured cuBLAS "
                                "ZGEMM peak"},
        "phases_ms": {k: round(v["ms"], 3) for k, v in prof.items()}, "dominant_phase": dom,
    }
    try:
        import ctypes as C
        from timeevomps_jl_b200._lib import load as _load
        out = (C.c_double * 11)()
        _load().tevo_debug_counters.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
        _load().tevo_debug_counters(ctx.h, out, 11)
        line["centre_moves"] = {"cholqr_single_pass": int(out[0]), "cholqr_two_pass": int(out[1]),
                                "eigen_fallback": int(out[2]), "note": "cumulative since context creation"}
    except Exception:
        pass
    if not args.no_cpu:
        arm = CpuArm(args)
        line["cpu_baseline"] = arm.sample(args.cpu_sample_gates)
        if args.cpu_1thread:
            # the baseline is stated for all BLAS threads (value, cores) and for one thread
            line["cpu_baseline"].update(arm.sample_one_thread())
    print(json.dumps(line))


if __name__ == "__main__":
    main()
