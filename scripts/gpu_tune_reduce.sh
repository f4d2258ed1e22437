#!/bin/bash
# GPU box: k x k solve (als_primal_kernel MODE_REDUCE) under different register caps (min CTAs per SM)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
: > gpurun_out/tune_reduce.jsonl
for flags in "-DYCNR_DUAL_REGS=0" "-DYCNR_DUAL_REGS=112" "-DYCNR_DUAL_REGS=100" "-DYCNR_DUAL_REGS=88" "-DYCNR_DUAL_REGS=0" "-DYCNR_DUAL_REGS=100"; do
  YCNR_NVCC_FLAGS="$flags -Xptxas -v" python scripts/quick_bench.py mal 100 3 >> gpurun_out/tune_reduce.jsonl 2>> gpurun_out/tune_reduce.err
done
touch you_can_not_recommend_b200/csrc/ycnr_als.cu
python -c "from you_can_not_recommend_b200 import build; build.build_cuda()"
python - <<'PY'
import json
for l in open('gpurun_out/tune_reduce.jsonl'):
    d=json.loads(l); print(d['flags'], round(d['wall_ms_per_step'],2), d['classes'], {k:v[0] for k,v in d['dual_bins'].items() if int(k)>=13})
PY
