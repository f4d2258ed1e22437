#!/bin/bash
# GPU box: k x k solve tiles per thread at k = 32 / 64 (BASELINE configs[4])
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
: > gpurun_out/tune_k.jsonl
for flags in "-DYCNR_REDUCE_TPT=3" "-DYCNR_REDUCE_TPT=2" "-DYCNR_REDUCE_TPT=1"; do
  for k in 32 64; do
    YCNR_NVCC_FLAGS="$flags" python scripts/quick_bench.py mal $k 3 >> gpurun_out/tune_k.jsonl 2>> gpurun_out/tune_k.err
  done
done
touch you_can_not_recommend_b200/csrc/ycnr_als.cu
python -c "from you_can_not_recommend_b200 import build; build.build_cuda()"
python - <<'PY'
import json
for l in open('gpurun_out/tune_k.jsonl'):
    d=json.loads(l); print(d['flags'], d['k'], round(d['wall_ms_per_step'],2), d['classes'])
PY
