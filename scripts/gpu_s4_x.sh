#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
for W in mal ml-1m; do for S in 0 1; do
  if [ $S = 1 ]; then export YCNR_SPREAD_BULK=1; else unset YCNR_SPREAD_BULK; fi
  python bench.py --workload $W --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$W spread=$S ms/step', round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})"
done; done
