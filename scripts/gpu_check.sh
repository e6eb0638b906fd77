#!/bin/bash
# One gpurun call: smoke, GPU parity tests, short bench.  Everything is logged under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
export WN_TIMEOUT_MS=${WN_TIMEOUT_MS:-3000}
echo "== smoke" | tee gpurun_out/smoke.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
echo "== pytest gpu"
timeout ${PYTEST_TIMEOUT:-900} python -m pytest tests -m gpu -q -x ${PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
echo "== bench"
timeout 600 python bench.py --steps ${BENCH_STEPS:-3} --warmup 3 ${BENCH_ARGS} > gpurun_out/bench.log 2>&1; echo "bench rc=$?" | tee -a gpurun_out/bench.log
tail -5 gpurun_out/bench.log
