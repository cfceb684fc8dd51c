#!/bin/bash
# layer-parallel mode: lanes x panel-grid cap (N=48, chi=1024, one TEBD4 step each, ~5 s per point)
cd "$(dirname "$0")/.."
run() {
  TEVO_LANES=$1 TEVO_PANEL_MAX_CTAS=$2 TEVO_BENCH_BUDGET_S=100 timeout 200 python bench.py --mode parallel --nsites 48 --steps 1 --warmup 1 --no-e2e --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); p=d['phases_ms']; print('lanes', $1, 'cap', $2, 'ms_per_step', round(d['ms_per_step'],1), 'main-lane panel', round(p.get('tridiag_panel',0)), 'her2k', round(p.get('tridiag_her2k',0)), 'stedc', round(p.get('stedc',0)))"
}
run 3 0
run 2 0
run 2 60
run 3 40
run 3 32
run 4 0
run 4 28
run 5 24
run 6 20
echo "== e2e path with direct pinned download (N=64 chi=1024 sequential, 1 step)"
TEVO_BENCH_BUDGET_S=200 timeout 300 python bench.py --nsites 64 --steps 1 --warmup 1 --no-cpu --no-parallel-mode 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', d['value'], 'e2e', d['e2e'])"
timeout 100 python -m pytest tests/test_gpu_tebd.py -m gpu -x -q -k roundtrip 2>&1 | tail -2
