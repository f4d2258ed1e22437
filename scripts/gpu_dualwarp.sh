#!/bin/bash
# GPU box: parity tests of the dual kernels, then per-bin times of a MAL iteration for the tile kernels (mask 0)
# and the warp-per-system kernel (all bins)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
timeout 900 python -m pytest tests/test_gpu_als.py -m gpu -q -x > gpurun_out/dw_pytest.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/dw_pytest.log
: > gpurun_out/dw_bins.jsonl
for mask in 0 ffffff; do
  YCNR_DUAL_WARP=$mask timeout 600 python scripts/quick_bench.py mal 100 3 >> gpurun_out/dw_bins.jsonl 2>> gpurun_out/dw_bins.err
done
cat gpurun_out/dw_bins.jsonl
