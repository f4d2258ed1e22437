#!/bin/bash
# Round evidence: GPU parity suite, smoke, full bench line (value + e2e + cpu_baseline), reference arm,
# launch list and one ncu --set full capture of every kernel class of the timed step.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
echo "== pytest gpu"
timeout 400 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
echo "== bench (default flags)"
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "exit $?"; tail -c 5000 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
echo "== bench reference arm"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "exit $?"; tail -c 1500 gpurun_out/bench_reference.json
echo "== launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 16 -c 16 --csv --log-file gpurun_out/launches_mal.csv python bench.py --workload mal --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/launches_bench.log 2>&1; echo "exit $?"
echo "== ncu full (timed step, all 16 launches)"
timeout 600 ncu --set full --clock-control none --import-source on -s 16 -c 16 -o gpurun_out/prof_mal_step -f python bench.py --workload mal --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_step.log 2>&1; echo "exit $?"
ls -la gpurun_out | head -30
