#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
echo "== tc diag"; timeout 120 python scripts/tc_diag.py 2>&1 | tail -5
echo "== pytest gpu"
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
for MIN in 0 512; do
echo "== bench mal tc (tc_min_cols=$MIN)"
YCNR_TC_MIN=$MIN timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --gram tc > gpurun_out/bench_mal_tc_$MIN.json 2> gpurun_out/bench_mal_tc_$MIN.err; echo "exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_mal_tc_$MIN.json')); print(d['ms_per_step'], d['rmse']); print({k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()}); print(d['roofline'])"; tail -3 gpurun_out/bench_mal_tc_$MIN.err
done
echo "== ncu gram_tc"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gram_tc' -s 2 -c 2 -o gpurun_out/prof_mal_tc -f python bench.py --workload mal --steps 1 --warmup 1 --no-e2e --no-cpu --gram tc > gpurun_out/ncu_tc.log 2>&1; echo "exit $?"
