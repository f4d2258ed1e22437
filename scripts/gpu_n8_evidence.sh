#!/bin/bash
# Round-2 multi-GPU evidence (8 x B200): MAL default line, Netflix shape, wide systems — one torchrun each.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
N=${1:-8}
run() {  # name, extra bench args...
  local name=$1; shift
  SECONDS=0
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N "$@" > $O/r2_bench_${name}_n$N.json 2> $O/r2_bench_${name}_n$N.err
  echo "== $name: exit $? (${SECONDS}s)"; python - <<PY
import json
try:
    d=json.loads(open("$O/r2_bench_${name}_n$N.json").read())
    print(" ms/step", round(d["ms_per_step"],2), "G ratings/s", round(d["value"]/1e9,2), {k:round(v["ms_per_step"],2) for k,v in d["kernels"].items()})
    if d.get("e2e"): print(" e2e", round(d["e2e"]["ms_per_step"],2), d["e2e"]["phase_ms"])
    print(" rmse", d["rmse"])
except Exception as e:
    print(" no line:", e)
PY
}
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || tail -20 $O/build.log
run mal --no-cpu --e2e-python-steps 0
run netflix --workload netflix --no-cpu --e2e-python-steps 0 --e2e-large-portion 0
if [ "${2:-}" = "wide" ]; then
  run mal_k256 --factors 256 --no-cpu --no-e2e --steps 3 --warmup 2
  run mal_k128 --factors 128 --no-cpu --no-e2e --steps 3 --warmup 2
fi
