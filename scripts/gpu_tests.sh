#!/bin/bash
# GPU parity suite only (optionally a -k filter)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
timeout 600 python -m pytest tests -m gpu -q -x "$@" > gpurun_out/pytest_gpu_t.log 2>&1; echo "pytest exit $?"; tail -30 gpurun_out/pytest_gpu_t.log
