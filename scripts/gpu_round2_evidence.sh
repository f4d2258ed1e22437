#!/bin/bash
# Round-2 evidence (one B200): GPU parity suite, smoke, the default bench line (value + e2e legs + cpu_baseline), the
# reference arm, the launch list of one timed iteration and ncu --set full captures of every kernel of it.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total,driver_version --format=csv > $O/r2_gpu_info.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > $O/r2_build.log 2>&1 || tail -20 $O/r2_build.log
echo "== pytest gpu"; SECONDS=0
timeout 1200 python -m pytest tests -m gpu -q --durations=8 > $O/r2_pytest_gpu.log 2>&1; echo "pytest exit $? (${SECONDS}s)"; tail -12 $O/r2_pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $O/r2_smoke.log | cut -c1-300
echo "== bench default"; SECONDS=0
timeout 900 python bench.py > $O/r2_bench_mal_default.json 2> $O/r2_bench_mal_default.err; echo "exit $? (${SECONDS}s) lines $(wc -l < $O/r2_bench_mal_default.json)"; tail -3 $O/r2_bench_mal_default.err
echo "== bench reference arm"; SECONDS=0
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_bench_mal_reference.json 2> $O/r2_bench_mal_reference.err; echo "exit $? (${SECONDS}s)"; cut -c1-300 $O/r2_bench_mal_reference.json
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r2_launches_mal.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $O/r2_launches.log 2>&1; echo "exit $?"
echo "== ncu full: gram_tc, k x k solve, rmse"; SECONDS=0
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:gram_tc|als_primal|rmse_rows' -s 6 -c 6 -o $O/r2_main -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $O/r2_ncu_main.log 2>&1; echo "exit $? (${SECONDS}s)"
echo "== ncu full: dual bins"; SECONDS=0
timeout 900 ncu --set full --clock-control none -k 'regex:als_dual' -s 21 -c 21 -o $O/r2_dual -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $O/r2_ncu_dual.log 2>&1; echo "exit $? (${SECONDS}s)"
python scripts/ncu_summary.py $O/r2_main.ncu-rep > $O/r2_ncu_mal_main.txt
python scripts/ncu_summary.py $O/r2_dual.ncu-rep > $O/r2_ncu_mal_dual.txt
python scripts/ncu_traffic.py $O/r2_traffic_mal.json $O/r2_main.ncu-rep $O/r2_dual.ncu-rep | cut -c1-600
rm -f $O/r2_main.ncu-rep $O/r2_dual.ncu-rep
python - <<PY
import json
d=json.loads(open('$O/r2_bench_mal_default.json').read())
print('ms/step', round(d['ms_per_step'],2), 'value', round(d['value']/1e9,3), 'roof', d['roofline']['kernel'], round(d['roofline']['frac'],3))
print({k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()})
print('e2e', round(d['e2e']['ms_per_step'],1), round(d['e2e']['value']/1e9,3), 'cpu', round(d['cpu_baseline']['value']/1e6,3), 'M/s on', d['cpu_baseline']['cores'], 'cores')
PY
du -sh $O
