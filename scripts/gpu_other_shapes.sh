#!/bin/bash
# The other BASELINE shapes through the same bench (not bench lines of record: parity-test shapes)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
for W in netflix ml-1m; do
  SECONDS=0
  timeout 600 python bench.py --workload $W --steps 5 --warmup 3 > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; echo "== $W exit $? (${SECONDS}s)"; tail -2 gpurun_out/bench_$W.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$W.json').read())
print('ms/step', round(d['ms_per_step'],3), 'value', round(d['value']/1e9,3), 'roof', d['roofline']['kernel'], round(d['roofline']['frac'],3))
print({k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})
print('e2e', round(d['e2e']['ms_per_step'],2), 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value']/1e6,3), d['rmse'])
PY
done
