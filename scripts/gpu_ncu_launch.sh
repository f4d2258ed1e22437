#!/bin/bash
# ncu --set full (+source) of single launches: args TAG REGEX SKIP [SKIP...]; exports txt + source csv, drops the .ncu-rep
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=$1; RX=$2; shift 2
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
for SKIP in "$@"; do
  SECONDS=0
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$RX" -s $SKIP -c 1 \
    -o gpurun_out/prof_${TAG}_$SKIP -f python bench.py --workload mal --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_${TAG}_$SKIP.log 2>&1; echo "ncu exit $? (${SECONDS}s)"
  python scripts/ncu_summary.py gpurun_out/prof_${TAG}_$SKIP.ncu-rep > gpurun_out/prof_${TAG}_$SKIP.txt
  ncu -i gpurun_out/prof_${TAG}_$SKIP.ncu-rep --page source --csv > gpurun_out/prof_${TAG}_${SKIP}_source.csv 2>/dev/null
  rm -f gpurun_out/prof_${TAG}_$SKIP.ncu-rep
  cat gpurun_out/prof_${TAG}_$SKIP.txt
done
du -sh gpurun_out
