#!/bin/bash
# GPU box: parity tests with the row-register k x k solve, then class times of a MAL iteration with the tile
# Cholesky (YCNR_SOLVE_ROWS=0) and the row-register LDL^T (1)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
export YCNR_DUAL_WARP=${YCNR_DUAL_WARP:-0}
timeout 900 python -m pytest tests/test_gpu_als.py -m gpu -q -x > gpurun_out/sr_pytest.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/sr_pytest.log
: > gpurun_out/sr_bins.jsonl
for v in 0 1; do
  YCNR_SOLVE_ROWS=$v timeout 600 python scripts/quick_bench.py mal 100 3 >> gpurun_out/sr_bins.jsonl 2>> gpurun_out/sr_bins.err
done
for k in 64 96 128; do
  YCNR_SOLVE_ROWS=1 timeout 600 python scripts/quick_bench.py mal $k 2 >> gpurun_out/sr_bins.jsonl 2>> gpurun_out/sr_bins.err
done
python - <<'PY'
import json
for l in open('gpurun_out/sr_bins.jsonl'):
    d=json.loads(l); print(d['k'], round(d['wall_ms_per_step'],2), d['classes'], d['rmse'])
PY
