#!/bin/bash
# N-GPU bench runs launched exactly as the driver does (torchrun, one rank per GPU).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-2}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
run() {  # name, extra args...
  local name=$1; shift
  SECONDS=0
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N "$@" > gpurun_out/bench_n${N}_${name}.json 2> gpurun_out/bench_n${N}_${name}.err
  echo "== $name exit $? (${SECONDS}s)"; tail -c 3000 gpurun_out/bench_n${N}_${name}.json; grep -v -E "^W|OMP_NUM|^\*+$|^$" gpurun_out/bench_n${N}_${name}.err | tail -8
}
run nccl --steps 5 --warmup 3
run fused --steps 5 --warmup 3 --fused-peers --no-e2e --no-cpu
run reference --impl reference --steps 2 --warmup 1
