#!/bin/bash
# GPU box: mbarrier wait flavours of gram_tc (suspend-time hint / nanosleep between polls)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
: > gpurun_out/tune_wait.jsonl
for flags in "-DYCNR_TC_WAIT_HINT=0" "-DYCNR_TC_WAIT_HINT=1000" "-DYCNR_TC_WAIT_HINT=20000" "-DYCNR_TC_WAIT_SLEEP=32" "-DYCNR_TC_WAIT_SLEEP=128"; do
  for k in 100 32; do
    YCNR_NVCC_FLAGS="$flags" python scripts/quick_bench.py mal $k 3 >> gpurun_out/tune_wait.jsonl 2>> gpurun_out/tune_wait.err
  done
done
touch you_can_not_recommend_b200/csrc/ycnr_als.cu
python -c "from you_can_not_recommend_b200 import build; build.build_cuda()"
python - <<'PY'
import json
for l in open('gpurun_out/tune_wait.jsonl'):
    d=json.loads(l); print(d['flags'], d['k'], round(d['wall_ms_per_step'],2), d['classes']['gram_tc'], d['rmse']['rmseTestShift'])
PY
