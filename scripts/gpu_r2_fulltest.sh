#!/bin/bash
mkdir -p gpurun_out
export WN_TIMEOUT_MS=${WN_TIMEOUT_MS:-4000}
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout ${PYTEST_TIMEOUT:-2700} python -m pytest tests -m gpu -q --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -45 gpurun_out/pytest_gpu.log
