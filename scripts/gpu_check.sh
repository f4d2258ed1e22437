#!/bin/bash
# 1-GPU check of a change: parity suite, then the default bench line (value + e2e), optional extra bench args.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-x}; shift || true
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
SECONDS=0
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $? (${SECONDS}s)"; tail -4 gpurun_out/pytest_gpu_$TAG.log
SECONDS=0
timeout 600 python bench.py --no-cpu "$@" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $? (${SECONDS}s)"; tail -5 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_$TAG.json') if l.startswith('{')][-1])
print('ms/step', round(d['ms_per_step'],2), 'value', round(d['value']/1e9,3), 'roof', round(d['roofline']['frac'],3), d['roofline']['kernel'])
print({k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()})
print('e2e', d['e2e'] and (round(d['e2e']['ms_per_step'],1), {k:round(v,1) for k,v in d['e2e'].get('phase_ms',{}).items()}))
print(d['rmse'])
PY
