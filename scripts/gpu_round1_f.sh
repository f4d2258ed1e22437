#!/bin/bash
# pytest + ncu full capture of the timed step's kernels (MAL, auto path).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu.log
echo "== ncu full"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'als_dual_kernel<8|gram_tc|als_primal|rmse_rows' -s 8 -c 6 -o gpurun_out/prof_mal_r1e -f python bench.py --workload mal --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "exit $?"; tail -3 gpurun_out/ncu_full.log | cut -c1-200
ls -la gpurun_out
