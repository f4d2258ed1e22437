#!/bin/bash
# N-GPU bench: NCCL exchange (with e2e) and fused peer stores, launched as the driver does.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-8}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
run() {  # name, extra args...
  local name=$1; shift
  SECONDS=0
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N "$@" > gpurun_out/bench_n${N}_${name}.json 2> gpurun_out/bench_n${N}_${name}.err
  echo "== $name exit $? (${SECONDS}s)"; grep -v -E "^W|OMP_NUM|^\*+$|^$|NCCL version" gpurun_out/bench_n${N}_${name}.err | tail -6
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_n${N}_${name}.json') if l.startswith('{')][-1])
    print(d['n_gpus'], 'ms/step', round(d['ms_per_step'],2), 'wall', round(d['wall_ms_per_step'],2), 'value', round(d['value']/1e9,3), d['config']['replica_refresh'])
    print({k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()})
    print('e2e', d['e2e'] and (round(d['e2e']['ms_per_step'],1), {k:round(v,1) for k,v in d['e2e'].get('phase_ms',{}).items()}))
    print(d['rmse'])
except Exception as e:
    print('parse error', e)
PY
}
run fused --steps 5 --warmup 3 --fused-peers --no-e2e --no-cpu
run nccl --steps 5 --warmup 3 --no-cpu
