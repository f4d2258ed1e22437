#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
BENCH="python bench.py --workload mal --steps 1 --warmup 1 --no-e2e --no-cpu"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gram_tc -s 2 -c 2 -o gpurun_out/prof_tc -f $BENCH > gpurun_out/ncu_tc.log 2>&1; echo "exit $?"
ls -la gpurun_out | head
