#!/bin/bash
# GPU box: does the 128-byte alignment of the gathered factor rows matter for gram_tc?  k = 96 (384-byte rows, every
# row starts on a 128-byte line), 100 (400 bytes), 104 (416 bytes), 112 (448), 128 (512)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
: > gpurun_out/kalign.jsonl
for k in 96 100 104 112 124; do
  YCNR_DUAL_WARP=0 timeout 600 python scripts/quick_bench.py mal $k 3 >> gpurun_out/kalign.jsonl 2>> gpurun_out/kalign.err
done
python - <<'PY'
import json
for l in open('gpurun_out/kalign.jsonl'):
    d=json.loads(l); print(d['k'], d['wall_ms_per_step'], d['classes'])
PY
