#!/bin/bash
# Re-entry baseline: GPU parity suite, smoke, bench (auto = tcgen05 path, ffma path), launch list.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
echo "== bench mal auto"
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_mal_auto.json 2> gpurun_out/bench_mal_auto.err; echo "exit $?"; tail -c 6000 gpurun_out/bench_mal_auto.json; tail -5 gpurun_out/bench_mal_auto.err
echo "== bench mal ffma"
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --gram ffma > gpurun_out/bench_mal_ffma.json 2> gpurun_out/bench_mal_ffma.err; echo "exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_mal_ffma.json')); print(d['ms_per_step'], d['rmse']); print({k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()})"; tail -3 gpurun_out/bench_mal_ffma.err
echo "== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_mal_auto.csv python bench.py --workload mal --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/launches_bench.log 2>&1
echo "exit $?"; tail -2 gpurun_out/launches_bench.log | cut -c1-300
