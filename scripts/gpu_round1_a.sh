#!/bin/bash
# First contact with the B200: sanitizer on tiny cases, the GPU parity suite, a first bench line.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu_info.txt 2>&1
nproc >> gpurun_out/gpu_info.txt; free -g | head -2 >> gpurun_out/gpu_info.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
echo "== sanitizer"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_als.py -q -x \
  -k "tiny or zero_cols or gather or rmse_portion" > gpurun_out/sanitizer.log 2>&1
echo "sanitizer exit $?"; tail -15 gpurun_out/sanitizer.log
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -40 gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
echo "== bench ml-1m"
timeout 600 python bench.py --workload ml-1m --steps 3 --warmup 3 --cpu-seconds 6 > gpurun_out/bench_ml1m.json 2> gpurun_out/bench_ml1m.err; echo "exit $?"; tail -c 3000 gpurun_out/bench_ml1m.json; tail -5 gpurun_out/bench_ml1m.err
echo "== bench mal"
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_mal.json 2> gpurun_out/bench_mal.err; echo "exit $?"; tail -c 4000 gpurun_out/bench_mal.json; tail -5 gpurun_out/bench_mal.err
