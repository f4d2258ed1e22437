#!/bin/bash
# N-GPU default bench only (as the driver launches it), short.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-2}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n${N}_final.json 2> gpurun_out/bench_n${N}_final.err
echo "exit $? lines $(wc -l < gpurun_out/bench_n${N}_final.json)"; grep -v -E "^W|OMP_NUM|^\*+$|^$|NCCL version" gpurun_out/bench_n${N}_final.err | tail -5
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n${N}_final.json').read())
print(d['n_gpus'], 'ms/step', round(d['ms_per_step'],2), 'value', round(d['value']/1e9,3), 'e2e', round(d['e2e']['ms_per_step'],1), d['rmse'])
PY
