#!/bin/bash
# Last check of the tree as committed: GPU parity suite, smoke, default bench line.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/sanity_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/sanity_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/sanity_bench.json 2> gpurun_out/sanity_bench.err; echo "bench exit $? lines $(wc -l < gpurun_out/sanity_bench.json)"
python - <<PY
import json
d=json.loads(open('gpurun_out/sanity_bench.json').read())
print('ms/step', round(d['ms_per_step'],2), 'value', round(d['value']/1e9,3), 'roof', d['roofline']['kernel'], round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['ms_per_step'],1), 'cpu', round(d['cpu_baseline']['value']/1e6,2))
PY
