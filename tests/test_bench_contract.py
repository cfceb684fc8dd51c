"""bench.py contract checks that need no GPU: the reference arm's JSON line, and that the product arm refuses to run
(instead of falling back to the CPU) when no CUDA device is present."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, env=e, capture_output=True,
                          text=True, timeout=600)


def test_reference_arm_line():
    p = _run(["--impl", "reference", "--chi", "16", "--nsites", "16", "--steps", "2", "--warmup", "1"])
    assert p.returncode == 0, p.stderr
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "tebd4_time_steps_per_sec" and d["unit"] == "steps/s"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and abs(d["ms_per_step"] * d["value"] - 1e3) < 1e-6 * 1e3
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_do_nothing():
    p = _run(["--impl", "reference", "--chi", "8", "--nsites", "12", "--steps", "1", "--warmup", "1", "--gpus", "2"],
             env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert p.returncode == 0, p.stderr
    assert not [l for l in p.stdout.splitlines() if l.startswith("{")]


def test_product_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = _run(["--chi", "8", "--nsites", "12", "--steps", "1", "--warmup", "1", "--no-cpu", "--no-e2e"])
    assert p.returncode != 0
    assert not [l for l in p.stdout.splitlines() if l.startswith("{")]
    msg = p.stderr + p.stdout
    assert "CUDA" in msg or "NVIDIA" in msg or "no CPU fallback" in msg


def test_budget_plan_cuts_counts_to_the_wall_clock_budget():
    """--steps/--warmup are upper bounds: the counts are cut so the run ends inside the budget (round 1's bench ran
    warmup + 2*steps full steps and was killed by the driver)."""
    sys.path.insert(0, ROOT)
    import importlib
    bench = importlib.import_module("bench")
    b = bench.Budget(540.0)
    b.elapsed = lambda: 95.0              # set-up + Ustart + the wall-timed first warm-up step
    # 52 s per step, driver asks for 20 + 5: three warm-up steps in total, then what still fits
    warm_more, timed = b.plan(52.0, 1, 5, 20, reserve=205.0)
    assert warm_more == 2 and timed == 2
    assert 95.0 + (warm_more + timed) * 52.0 + 205.0 <= 540.0
    # a fast configuration runs exactly what was asked for
    warm_more, timed = b.plan(0.5, 1, 5, 20, reserve=20.0)
    assert warm_more == 4 and timed == 20
    # nothing fits: still one timed step, no extra warm-up
    b.elapsed = lambda: 530.0
    assert b.plan(52.0, 1, 5, 20, reserve=50.0) == (0, 1)
    # room for two steps only: one more warm-up, one timed
    b.elapsed = lambda: 100.0
    warm_more, timed = b.plan(100.0, 1, 5, 20, reserve=200.0)
    assert (warm_more, timed) == (1, 1)
