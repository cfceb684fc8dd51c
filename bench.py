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

T_PROCESS_START = time.time()

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "tebd4_time_steps_per_sec"
UNIT = "steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2, help="upper bound on the timed steps (see --budget)")
    ap.add_argument("--warmup", type=int, default=3, help="upper bound on the warm-up steps (see --budget)")
    ap.add_argument("--budget", type=float, default=float(os.environ.get("TEVO_BENCH_BUDGET_S", "540")),
                    help="wall-clock budget of the whole run in seconds (env TEVO_BENCH_BUDGET_S).  A TEBD4 step at "
                         "N=256, chi=1024 takes tens of seconds, so --steps/--warmup are upper bounds: the first "
                         "warm-up step is timed and the counts are cut so that the run (warm-up, timed steps, one "
                         "end-to-end step, CPU sample) ends inside the budget.  The JSON line reports the counts "
                         "actually used (steps, warmup) and the requested ones (steps_requested, warmup_requested)")
    ap.add_argument("--no-parallel-mode", action="store_true",
                    help="at 1 GPU: skip the extra `parallel_mode` object (layer-parallel mode on the same chain)")
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


class Budget:
    """Wall-clock budget of the whole process (--budget / TEVO_BENCH_BUDGET_S).  --steps / --warmup are upper bounds."""

    def __init__(self, seconds):
        self.total = float(seconds)

    def elapsed(self):
        return time.time() - T_PROCESS_START

    def left(self):
        return self.total - self.elapsed()

    def plan(self, t_step, warm_done, warm_req, steps_req, reserve):
        """How many more warm-up steps and how many timed steps fit: at least one timed step; three warm-up steps in
        total (the timing rule) before a second timed step is added; never more than requested."""
        avail = self.left() - reserve
        n = max(int(avail // max(t_step, 1e-9)), 1)            # steps that still fit
        warm_more = max(0, min(warm_req - warm_done, max(min(3 - warm_done, n - 1), n - steps_req)))
        timed = max(1, min(steps_req, n - warm_more))
        return warm_more, timed


def debug_counters(ctx):
    import ctypes as C
    from timeevomps_jl_b200._lib import load as _load
    out = (C.c_double * 16)()
    _load().tevo_debug_counters.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
    _load().tevo_debug_counters(ctx.h, out, 16)
    return list(out)


def read_peaks():
    peaks, fp64 = {}, {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    try:
        fp64 = json.load(open(os.path.join(ROOT, "profiles", "r01_fp64_peaks.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    return hbm, src, float(fp64.get("zgemm_tflops_sustained", 36.8))


def panel_roofline(prof, c0, c1, hbm_peak, peak_src, label="tridiag_panel_kernel"):
    """Roofline object of the dominant kernel.  Numerator = the library's own count of the algorithmic bytes of the
    tridiagonalisation panels it launched inside the timed region (one read of the trailing matrix per column:
    sum over launches of pw x (n - p0)^2 x 16 B, tevo_debug_counters[13]); denominator = the CUDA-event time of those
    launches (in-library per-phase events on the library stream)."""
    if "tridiag_panel" not in prof:
        return None
    tk = prof["tridiag_panel"]
    bytes_alg = c1[13] - c0[13]
    nfact = int(c1[12] - c0[12])
    nlaunch = int(c1[14] - c0[14])
    ach = bytes_alg / (tk["ms"] * 1e-3) / 1e9
    traffic, tsrc = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_panel_traffic.json")))
        traffic, tsrc = tj["dram_bytes_per_launch"], tj.get("source")
    except Exception:
        pass
    total_prof = sum(v["ms"] for v in prof.values()) or 1.0
    return {"kernel": label, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
            "frac": ach / hbm_peak, "traffic": traffic, "traffic_source": tsrc, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": bytes_alg / max(nlaunch, 1), "launches": nlaunch,
            "factorisations": nfact, "avg_launch_ms": tk["ms"] / max(tk["count"], 1),
            "share_of_step": tk["ms"] / total_prof,
            "limiter": "latency and L2 traffic, not HBM bandwidth: every column of the Householder tridiagonalisation is a "
                       "chain of two grid-wide barriers and dependent L2 round trips (72 % of its time at n = 2048), and "
                       "the matrix-vector product runs at ~60 % of the L2 throughput cap; the trailing matrix stays L2-resident inside a "
                       "launch, so DRAM sees it about once per 64 columns (traffic) while the matrix-vector products "
                       "read it once per column (algorithmic bytes).  The HBM roof is the contract's denominator; the "
                       "kernel cannot reach it by construction of the one-stage algorithm"}


_REAL_STDOUT_FD = None


def quiet_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write banners there (NCCL prints its version on the first
    communicator), so file descriptor 1 is pointed at stderr for the duration of the run and `emit` writes the line
    to the saved descriptor."""
    global _REAL_STDOUT_FD
    if _REAL_STDOUT_FD is None:
        sys.stdout.flush()
        _REAL_STDOUT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT_FD, data)


def run_reference(args):
    """Reference arm: the CPU restatement of the reference path (oracle/) on all host BLAS threads.  One "step" of this
    arm = REF_UPDATES interior two-site updates at the benchmark's chi (a bounded sample of the 1275 updates of one
    TEBD4 step), extrapolated by the update count; --steps/--warmup are cut to the wall-clock budget."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    REF_UPDATES = 2
    budget = Budget(min(args.budget, 300.0))
    arm = CpuArm(args)
    first = arm.sample(REF_UPDATES)
    t_step = REF_UPDATES / (first["value"] * gates_per_step(args.nsites))
    warm_more, timed = budget.plan(t_step, 1, max(args.warmup, 1), args.steps, reserve=5.0)
    for _ in range(warm_more):
        arm.sample(REF_UPDATES)
    vals, cb = [], first
    for _ in range(timed):
        cb = arm.sample(REF_UPDATES)
        vals.append(cb["value"])
    v = float(1.0 / np.mean([1.0 / x for x in vals]))
    cb["value"] = v
    cb["sample"] = "each step = " + cb["sample"]
    cb["per_step_values"] = [float(x) for x in vals]
    line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": timed, "warmup": 1 + warm_more,
            "steps_requested": args.steps, "warmup_requested": args.warmup,
            "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "c128",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_string(args), "mode": "sequential (reference-faithful centre moves)",
                       "parallelism": "host CPU, all BLAS threads"},
            "cpu_baseline": cb,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def pinned_copy(torch, arrays):
    out = []
    for a in arrays:
        buf = torch.empty(max(a.size, 1) * 2, dtype=torch.float64, pin_memory=True)
        v = buf.numpy()[:a.size * 2].view(np.complex128).reshape(a.shape, order="F")
        v[...] = a
        out.append(v)
    return out


def run_sharded(args):
    """TEBD4 steps/s of ONE N-site chain sharded by contiguous segments over the ranks (BASELINE config 5):
    layer-parallel (Hastings, B-form) updates, one boundary tensor exchanged each way per cut bond and layer over
    NCCL P2P.  Strong scaling: the chain is fixed, the ranks split it.  Not the reference's sequential algorithm:
    identical without truncation, O(discarded weight) apart with it."""
    import torch
    import tevo_b200 as t
    from timeevomps_jl_b200 import replicas
    from timeevomps_jl_b200.sharded import DeviceSegment, ShardedTEBD, weighted_segment_bounds
    budget = Budget(args.budget)
    rank, world, local = replicas.world()
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = t.default_ctx()
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
    N, chi = args.nsites, args.chi
    # contiguous segments of equal estimated cost (~ chi_j^3 per site: the chain ends carry small bonds)
    wts = [float(min(chi, 2 ** min(j + 1, N - j - 1, 40))) ** 3 for j in range(N)]
    lo, hi = weighted_segment_bounds(wts, world, rank)
    seg = DeviceSegment.random(N, chi, lo, hi, ghost=(rank < world - 1), seed=0)
    sh = ShardedTEBD(seg, N, dist, rank, world, device="cuda", bounds=(lo, hi))
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
    barrier()
    t0 = time.perf_counter()
    step()
    barrier()
    t_step = replicas.max_over_ranks(time.perf_counter() - t0, dist, device="cuda")
    elapsed = replicas.max_over_ranks(budget.elapsed(), dist, device="cuda")
    budget.total -= max(0.0, elapsed - budget.elapsed())       # all ranks plan with the slowest rank's clock
    reserve = 15.0 + (0.0 if args.no_e2e else 1.3 * t_step + 10.0)
    warm_more, timed = budget.plan(t_step, 1, max(args.warmup, 1), args.steps, reserve)
    if dist is not None:      # identical counts on every rank
        tt = torch.tensor([warm_more, timed], dtype=torch.int64, device="cuda")
        dist.broadcast(tt, 0)
        warm_more, timed = int(tt[0].item()), int(tt[1].item())
    for _ in range(warm_more):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t.profile_enable(True)
    t.profile_read(reset=True)
    c0 = debug_counters(ctx)
    n0 = ctx.launch_count()
    b0 = sh.bytes_sent
    x0 = sh.exchange_s
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    t0 = time.perf_counter()
    for _ in range(timed):
        step()
    e1.record(stream)
    barrier()
    wall = time.perf_counter() - t0
    # the boundary exchange runs on torch's stream, ordered with the library stream by events, so the device time of
    # the region is the larger of the event span on the library stream and the synchronised wall clock of the slowest rank
    ms = max(e0.elapsed_time(e1), 0.0)
    ms = replicas.max_over_ranks(ms, dist, device="cuda")
    launches = ctx.launch_count() - n0
    prof = t.profile_read(reset=True)
    t.profile_enable(False)
    c1 = debug_counters(ctx)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms / timed
    value = 1e3 / ms_per_step
    bytes_per_step = int((sh.bytes_sent - b0) / max(timed, 1))
    # wall time this rank spent inside the boundary exchange (includes waiting for the neighbour = load imbalance)
    exch_ms = replicas.max_over_ranks((sh.exchange_s - x0) * 1e3 / max(timed, 1), dist, device="cuda")
    # end to end: segment uploaded from pinned host memory, one step, segment downloaded
    e2e = None
    if not args.no_e2e:
        import ctypes as C
        from timeevomps_jl_b200._lib import load as _load, check as _check
        nl = seg.nloc + (1 if rank < world - 1 else 0)
        pinned = pinned_copy(torch, [seg.psi.site(i) for i in range(nl)])
        lams = [seg.psi.schmidt_values(i) for i in range(nl + 1)]
        nbytes = sum(a.nbytes for a in pinned) + sum(l.nbytes for l in lams)
        barrier()
        t0 = time.perf_counter()
        for i, a in enumerate(pinned):
            seg.psi.set_site(i, a)
        for i, l in enumerate(lams):
            _check(_load().tevo_mps_set_lambda(seg.psi.h, i, len(l), l.ctypes.data_as(C.POINTER(C.c_double))), ctx.h)
        _check(_load().tevo_mps_mark_bform(seg.psi.h, 1), ctx.h)
        step()
        for i in range(nl):
            seg.psi.site(i, out=pinned[i])         # straight into the pinned host buffers
        lams = [seg.psi.schmidt_values(i) for i in range(nl + 1)]
        barrier()
        dt_e2e = replicas.max_over_ranks(time.perf_counter() - t0, dist, device="cuda")
        e2e = {"value": 1.0 / dt_e2e, "unit": UNIT, "h2d_bytes_per_step": int(nbytes),
               "d2h_bytes_per_step": int(nbytes), "steps": 1}
    if rank != 0:
        return
    hbm_peak, peak_src, zpeak = read_peaks()
    roofline = panel_roofline(prof, c0, c1, hbm_peak, peak_src, "tridiag_panel_kernel (rank 0)")
    flops_step = algorithmic_flops_per_gate(chi) * gates_per_step(N)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": timed, "warmup": 1 + warm_more,
        "steps_requested": args.steps, "warmup_requested": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "c128",
        "data": "synthetic",
        "config": {"workload": workload_string(args),
                   "mode": "layer-parallel (Hastings B-form) TEBD: NOT the sequential reference algorithm; equal to it "
                           "without truncation, O(discarded weight) apart with it.  The 1-GPU line reports this mode "
                           "as `parallel_mode` beside the sequential `value`: scaling efficiency is value(N) / "
                           "(N x parallel_mode.value)",
                   "parallelism": ("1 GPU" if world == 1 else
                                   "chain split into %d contiguous segments, one per GPU; boundary tensor + Schmidt "
                                   "values exchanged per cut bond and layer by NCCL send/recv" % world),
                   "l2_policy": "inputs larger than L2 (segment of an 8 GiB MPS streamed per step)"},
        "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e, "roofline": roofline,
        "fp64_tensor": {"algorithmic_tflops": flops_step / (ms_per_step * 1e-3) / 1e12, "peak_tflops": zpeak * world,
                        "frac": flops_step / (ms_per_step * 1e-3) / 1e12 / (zpeak * world)},
        "segment_rank0": [int(lo), int(hi)],
        "boundary_bytes_sent_per_step_rank0": bytes_per_step,
        "boundary_exchange_ms_per_step_max_rank": exch_ms,
        "boundary_transfer_ms_per_step_at_770GBs": bytes_per_step / 770e9 * 1e3,
        "wall_ms_per_step": wall * 1e3 / timed,
        "phases_ms": {k: round(v["ms"], 3) for k, v in prof.items()},
        "budget_s": args.budget, "wall_s": round(budget.elapsed(), 1),
    }
    emit(line)


def run_tdvp(args):
    """Secondary workload: one two-site TDVP step (right + left sweep, src/tdvp.jl:54-95) of the XXZ chain at the
    BASELINE config-4 shape (default N=64 via --nsites, chi=1024, krylovdim=30, exp_tol=1e-12); one independent
    h value per GPU (replicas).  Reported as an extra JSON line for profiles/, not the headline metric."""
    import torch
    import tevo_b200 as t
    from timeevomps_jl_b200 import replicas
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
    c0 = debug_counters(ctx)
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
    c1 = debug_counters(ctx)
    if rank != 0:
        return
    _, _, zpeak = read_peaks()
    heff_flops = c1[11] - c0[11]
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
                         "note": "effective rate in the 8*M*N*K convention: the 64x32 / 32x32 ZGEMM kernels form a "
                                 "complex product from three real multiplications (csrc/zgemm.cu), cuBLAS ZGEMM from four",
                         "share_of_step": hk["ms"] / (sum(v["ms"] for v in prof.values()) or 1.0)},
            "phases_ms": {k: round(v["ms"], 3) for k, v in prof.items()}}
    emit(line)


def run_sequential(args):
    """The headline: the reference's own algorithm (gates of a layer applied sequentially, centre moved between gates)
    on one GPU; with torchrun and --mode sequential, one independent trajectory per GPU (replicas, no collective)."""
    import torch
    import tevo_b200 as t
    from timeevomps_jl_b200 import replicas
    budget = Budget(args.budget)
    rank, world, local = replicas.world()
    torch.cuda.set_device(local)
    dist = None
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

    psi = t.MPS.random(N, chi, seed=replicas.trajectory_seed(rank))
    ctx.synchronize()
    Ustart, Us, Uend = t.time_evo_gates(args.dt, t.gates(model_bondop(t, args.model, N)), t.TEBD4())
    kw = dict(maxdim=chi, mindim=chi)
    state = {"rev": False, "psi": psi}

    def layer(U):
        t.apply_gates(state["psi"], U, reversed=state["rev"], **kw)
        state["rev"] = not state["rev"]

    def step():
        for U in Us:
            layer(U)

    for U in Ustart:
        layer(U)
    barrier()
    t0 = time.perf_counter()
    step()                                     # warm-up step 1, wall-timed: sizes the rest of the run
    barrier()
    t_step = replicas.max_over_ranks(time.perf_counter() - t0, dist, device="cuda")
    want_par = (world == 1) and not args.no_parallel_mode
    reserve = 20.0
    if not args.no_e2e:
        reserve += 1.2 * t_step + 15.0         # one step + 2 x 8 GiB over PCIe
    if not args.no_cpu:
        reserve += 15.0
    if want_par:
        reserve += 2 * 0.75 * t_step + 15.0    # one warm-up + one timed layer-parallel step + the B-form conversion
    warm_more, timed = budget.plan(t_step, 1, max(args.warmup, 1), args.steps, reserve)
    if dist is not None:
        tt = torch.tensor([warm_more, timed], dtype=torch.int64, device="cuda")
        dist.broadcast(tt, 0)
        warm_more, timed = int(tt[0].item()), int(tt[1].item())
    for _ in range(warm_more):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t.profile_enable(True)
    t.profile_read(reset=True)
    c0 = debug_counters(ctx)
    n0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(timed):
        step()
    e1.record(stream)
    barrier()
    launches = ctx.launch_count() - n0
    ms = e0.elapsed_time(e1)
    prof = t.profile_read(reset=True)
    t.profile_enable(False)
    c1 = debug_counters(ctx)
    clocks = sampler.stop() if rank == 0 else None
    ms = replicas.max_over_ranks(ms, dist, device="cuda")
    ms_per_step = ms / timed
    value = world * 1e3 / ms_per_step

    # ---- end-to-end through the public API with host buffers: upload, one step, download ----
    e2e = None
    if not args.no_e2e:
        psi = state["psi"]
        pinned = pinned_copy(torch, psi.tensors())
        nbytes = sum(a.nbytes for a in pinned)
        llim, rlim = psi.limits()
        state["psi"] = None
        del psi
        barrier()
        t0 = time.perf_counter()
        state["psi"] = t.MPS.from_tensors(pinned, llim=llim, rlim=rlim)
        step()
        out = state["psi"].tensors(out=pinned)     # the device copies straight into the pinned host buffers
        del out
        barrier()
        dt_e2e = replicas.max_over_ranks(time.perf_counter() - t0, dist, device="cuda")
        e2e = {"value": world / dt_e2e, "unit": UNIT, "h2d_bytes_per_step": int(nbytes),
               "d2h_bytes_per_step": int(nbytes), "steps": 1}
        del pinned

    # ---- the same chain in the layer-parallel mode (what the >1-GPU runs shard): reference point for scaling ----
    parallel_mode = None
    if want_par and budget.left() > 2 * 0.75 * t_step + 30.0:
        from timeevomps_jl_b200.tebd import apply_gates_parallel
        psi = state["psi"]
        psi.to_bform()

        def pstep():
            for U in Us:
                apply_gates_parallel(psi, U, **kw)

        barrier()
        t0 = time.perf_counter()
        pstep()
        barrier()
        tp1 = time.perf_counter() - t0
        npar = 1 if budget.left() < 2.2 * tp1 + 30.0 else 2
        nl0 = ctx.launch_count()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record(stream)
        for _ in range(npar):
            pstep()
        p1.record(stream)
        barrier()
        pms = p0.elapsed_time(p1) / npar
        parallel_mode = {"value": 1e3 / pms, "unit": UNIT, "ms_per_step": pms, "steps": npar, "warmup": 1,
                         "gpu_launches": int(ctx.launch_count() - nl0),
                         "mode": "layer-parallel (Hastings B-form), the mode `bench.py --gpus N>1` shards; divide "
                                 "value(N) by N x this value for the scaling efficiency in the same mode"}
    if rank != 0:
        return
    hbm_peak, peak_src, zgemm_peak = read_peaks()
    roofline = panel_roofline(prof, c0, c1, hbm_peak, peak_src)
    total_prof = sum(v["ms"] for v in prof.values()) or 1.0
    dom = max(prof.items(), key=lambda kv: kv[1]["ms"])[0] if prof else None
    flops_step = algorithmic_flops_per_gate(chi) * gates_per_step(N)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": timed, "warmup": 1 + warm_more,
        "steps_requested": args.steps, "warmup_requested": args.warmup,
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
        "parallel_mode": parallel_mode,
        "phases_ms": {k: round(v["ms"], 3) for k, v in prof.items()}, "dominant_phase": dom,
        "phase_share": {k: round(v["ms"] / total_prof, 4) for k, v in prof.items()},
        "centre_moves": {"cholqr_single_pass": int(c1[0] - c0[0]), "cholqr_two_pass": int(c1[1] - c0[1]),
                         "eigen_fallback": int(c1[2] - c0[2]), "note": "inside the timed region"},
        "budget_s": args.budget,
    }
    if not args.no_cpu:
        arm = CpuArm(args)
        one = arm.sample(1)
        per_gate = 1.0 / (one["value"] * gates_per_step(N))
        more = int(max(0, min(args.cpu_sample_gates - 1, (min(budget.left(), 40.0) - 5.0) // max(per_gate, 1e-3))))
        line["cpu_baseline"] = arm.sample(more) if more > 0 else one
        if args.cpu_1thread:
            # the baseline is stated for all BLAS threads (value, cores) and for one thread
            line["cpu_baseline"].update(arm.sample_one_thread())
    line["wall_s"] = round(budget.elapsed(), 1)
    emit(line)


# ---------------------------------------------------------------------------------------------------------
def main():
    args = parse()
    quiet_stdout()
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
    run_sequential(args)


if __name__ == "__main__":
    main()
