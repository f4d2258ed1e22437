#!/bin/bash
# 2 x B200: the two-rank NCCL parity test, the default bench line at N = 2, and (on one of the GPUs) the launch list
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || tail -20 $O/build.log
timeout 600 python -m pytest tests/test_gpu_parity_configs.py -m gpu -q -k two_rank > $O/r2_pytest_two_rank.log 2>&1; echo "two-rank test exit $?"; tail -3 $O/r2_pytest_two_rank.log
SECONDS=0
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --no-cpu --e2e-python-steps 0 --e2e-large-portion 0 > $O/r2_bench_mal_n2.json 2> $O/r2_bench_mal_n2.err; echo "n2 bench exit $? (${SECONDS}s)"
python - <<PY
import json
d=json.loads(open('$O/r2_bench_mal_n2.json').read())
print(d['n_gpus'], 'ms/step', round(d['ms_per_step'],2), 'value', round(d['value']/1e9,3), {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()})
print('e2e', round(d['e2e']['ms_per_step'],1), d['e2e']['phase_ms'], d['rmse'])
PY
CUDA_VISIBLE_DEVICES=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r2_launches_mal.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $O/r2_launches.log 2>&1; echo "launch list exit $?"
