#!/bin/bash
# Round evidence: GPU parity suite, smoke, the default bench line (value + e2e + cpu_baseline), the reference arm,
# the launch list of one timed iteration and ncu --set full captures of every kernel of it (summaries + traffic).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total,driver_version --format=csv > $O/gpu_info.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || tail -20 $O/build.log
echo "== pytest gpu"; SECONDS=0
timeout 900 python -m pytest tests -m gpu -q --durations=5 > $O/final_pytest_gpu.log 2>&1; echo "pytest exit $? (${SECONDS}s)"; tail -9 $O/final_pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $O/final_smoke.log
echo "== bench default"; SECONDS=0
timeout 900 python bench.py > $O/final_bench_default.json 2> $O/final_bench_default.err; echo "exit $? (${SECONDS}s) lines $(wc -l < $O/final_bench_default.json)"; tail -3 $O/final_bench_default.err
echo "== bench reference arm"; SECONDS=0
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/final_bench_reference.json 2> $O/final_bench_reference.err; echo "exit $? (${SECONDS}s)"; cut -c1-400 $O/final_bench_reference.json
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/final_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $O/final_launches.log 2>&1; echo "exit $?"
echo "== ncu full: gram_tc, k x k solve, rmse"; SECONDS=0
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:gram_tc|als_primal|rmse_rows' -s 7 -c 7 -o $O/final_main -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $O/final_ncu_main.log 2>&1; echo "exit $? (${SECONDS}s)"
echo "== ncu full: dual bins"; SECONDS=0
timeout 900 ncu --set full --clock-control none -k 'regex:als_dual' -s 21 -c 21 -o $O/final_dual -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $O/final_ncu_dual.log 2>&1; echo "exit $? (${SECONDS}s)"
python scripts/ncu_summary.py $O/final_main.ncu-rep > $O/final_ncu_main.txt
python scripts/ncu_summary.py $O/final_dual.ncu-rep > $O/final_ncu_dual.txt
python scripts/ncu_traffic.py $O/final_traffic_mal.json $O/final_main.ncu-rep $O/final_dual.ncu-rep | cut -c1-600
ncu -i $O/final_main.ncu-rep --page source --csv > $O/final_main_source.csv 2>/dev/null
python scripts/ncu_waits.py $O/final_main_source.csv 0 > $O/final_gram_tc_waits.txt 2>&1
rm -f $O/final_main.ncu-rep $O/final_dual.ncu-rep $O/final_main_source.csv
python - <<PY
import json
d=json.loads(open('$O/final_bench_default.json').read())
print('ms/step', round(d['ms_per_step'],2), 'value', round(d['value']/1e9,3), 'roof', d['roofline']['kernel'], round(d['roofline']['frac'],3))
print({k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()})
print('e2e', round(d['e2e']['ms_per_step'],1), round(d['e2e']['value']/1e9,3), 'cpu', round(d['cpu_baseline']['value']/1e6,3), 'M/s on', d['cpu_baseline']['cores'], 'cores')
PY
du -sh $O
