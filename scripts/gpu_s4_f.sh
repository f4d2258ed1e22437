#!/bin/bash
# ncu --set full (+source) of selected launches: args TAG KERNEL_REGEX SKIP COUNT
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=$1; RX=$2; SKIP=$3; CNT=$4
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
SECONDS=0
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$RX" -s $SKIP -c $CNT \
  -o gpurun_out/prof_$TAG -f python bench.py --workload mal --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu exit $? (${SECONDS}s)"
python scripts/ncu_summary.py gpurun_out/prof_$TAG.ncu-rep > gpurun_out/prof_$TAG.txt
ncu -i gpurun_out/prof_$TAG.ncu-rep --page source --csv > gpurun_out/prof_${TAG}_source.csv 2>/dev/null
grep -E "^==|time_duration" gpurun_out/prof_$TAG.txt
