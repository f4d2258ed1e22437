#!/bin/bash
# Quick iteration: GPU parity suite + MAL bench (device-resident leg only) with the per-kernel table.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu.log
for G in ${GRAMS:-auto}; do
echo "== bench mal $G"
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --gram $G > gpurun_out/bench_mal_$G.json 2> gpurun_out/bench_mal_$G.err; echo "exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_mal_$G.json')); print(d['ms_per_step'], d['rmse']); print({k:(round(v['ms_per_step'],2), v['launches_per_step']) for k,v in d['kernels'].items()}); print(d['roofline'])"; tail -3 gpurun_out/bench_mal_$G.err
done
