#!/bin/bash
# tcgen05 Gram diagnostics + e2e path check + bench.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
echo "== tc diag"
timeout 120 python scripts/tc_diag.py 2>&1 | tail -20
echo "== pytest gpu (quick subset)"
timeout 900 python -m pytest tests -m gpu -q -x -k "not full_size" 2>&1 | tail -15
echo "== bench mal ffma (new e2e path)"
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/bench_mal_ffma.json 2> gpurun_out/bench_mal_ffma.err; echo "exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_mal_ffma.json')); print(d['ms_per_step'], d['e2e']); print({k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()})"; tail -3 gpurun_out/bench_mal_ffma.err
echo "== bench mal tc"
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --gram tc > gpurun_out/bench_mal_tc.json 2> gpurun_out/bench_mal_tc.err; echo "exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_mal_tc.json')); print(d['ms_per_step'], d['rmse']); print({k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()})"; tail -3 gpurun_out/bench_mal_tc.err
