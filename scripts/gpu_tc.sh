#!/bin/bash
# TC kernel iteration: diag, GPU parity suite, MAL bench (device leg), launch list.  Tight timeouts: a hung
# kernel must not eat the GPU budget.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
echo "== tc diag"; timeout 90 python scripts/tc_diag.py 2>&1 | tail -5; [ ${PIPESTATUS[0]} -eq 0 ] || { echo "TC DIAG FAILED/HUNG"; exit 1; }
echo "== pytest gpu"
timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
echo "== bench mal auto"
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_mal_auto.json 2> gpurun_out/bench_mal_auto.err; echo "exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_mal_auto.json')); print(d['ms_per_step'], d['rmse']); print({k:(round(v['ms_per_step'],2), v['launches_per_step']) for k,v in d['kernels'].items()}); print(d['roofline'])"; tail -3 gpurun_out/bench_mal_auto.err
echo "== launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 16 -c 16 --csv --log-file gpurun_out/launches_mal_auto.csv python bench.py --workload mal --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/launches_bench.log 2>&1
grep -E "gram_tc|als_primal" gpurun_out/launches_mal_auto.csv | cut -d, -f5,14- | head
