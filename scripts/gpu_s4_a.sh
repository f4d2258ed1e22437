#!/bin/bash
# Session-4 baseline at HEAD: GPU parity suite, smoke, default bench line (what the driver runs).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
echo "== pytest gpu"
SECONDS=0
timeout 900 python -m pytest tests -m gpu -q -x --durations=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $? (${SECONDS}s)"; tail -18 gpurun_out/pytest_gpu.log
echo "== smoke"
SECONDS=0
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $? (${SECONDS}s)"; tail -3 gpurun_out/smoke.log
echo "== bench default"
SECONDS=0
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "exit $? (${SECONDS}s)"; tail -c 4500 gpurun_out/bench_default.json; tail -5 gpurun_out/bench_default.err
