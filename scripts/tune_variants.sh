#!/bin/bash
# Tuning experiment (GPU box): rebuild libycnr_als.so with different compile-time knobs and print the
# per-class / per-dual-bin device times of a MAL iteration for each.  Output: gpurun_out/tune_<tag>.jsonl
tag=${1:-r2}
shift
out=gpurun_out/tune_${tag}.jsonl
: > $out
for flags in "$@"; do
  YCNR_NVCC_FLAGS="$flags" python scripts/quick_bench.py mal 100 3 >> $out 2>> gpurun_out/tune_${tag}.err
done
# leave the default build behind
touch you_can_not_recommend_b200/csrc/ycnr_als.cu
python -c "from you_can_not_recommend_b200 import build; build.build_cuda()"
cat $out
