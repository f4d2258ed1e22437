#!/bin/bash
# N-GPU bench with the DEFAULT flags, launched exactly as the driver does; then the NCCL-exchange variant (no e2e).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-2}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
run() {
  local name=$1; shift
  SECONDS=0
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N "$@" > gpurun_out/bench_n${N}_${name}.json 2> gpurun_out/bench_n${N}_${name}.err
  echo "== $name exit $? (${SECONDS}s) stdout lines: $(wc -l < gpurun_out/bench_n${N}_${name}.json)"; grep -v -E "^W|OMP_NUM|^\*+$|^$|NCCL version" gpurun_out/bench_n${N}_${name}.err | tail -6
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n${N}_${name}.json').read())
    print(d['n_gpus'], 'ms/step', round(d['ms_per_step'],2), 'value', round(d['value']/1e9,3), d['config']['replica_refresh'], 'launches', d['gpu_launches'])
    print({k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()})
    print('e2e', d['e2e'] and (round(d['e2e']['ms_per_step'],1), round(d['e2e']['value']/1e9,3), {k:round(v,1) for k,v in d['e2e'].get('phase_ms',{}).items()}))
    print(d['rmse'])
except Exception as e:
    print('parse error', e)
PY
}
run default --steps 5 --warmup 3
run nccl --steps 5 --warmup 3 --nccl-exchange --no-e2e --no-cpu
