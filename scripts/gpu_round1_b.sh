#!/bin/bash
# ncu: launch list of one bench step + full captures of the top kernels (MAL shape).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
BENCH="python bench.py --workload mal --steps 1 --warmup 1 --no-e2e --no-cpu"
echo "== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_mal.csv $BENCH > gpurun_out/launches_bench.log 2>&1
echo "exit $?"; tail -2 gpurun_out/launches_bench.log | cut -c1-300
echo "== full capture (timed step only: skip the warm-up step's launches)"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'als_primal_kernel|als_dual_kernel|rmse_rows' -s 13 -c 13 -o gpurun_out/prof_mal_v1 -f $BENCH > gpurun_out/full_bench.log 2>&1
echo "exit $?"; tail -3 gpurun_out/full_bench.log | cut -c1-300
ls -la gpurun_out/
