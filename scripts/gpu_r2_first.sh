#!/bin/bash
# First GPU contact of the cluster engine: exchange microbenchmark, smoke, golden parity, timing sweep.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
export WN_TIMEOUT_MS=${WN_TIMEOUT_MS:-3000}
echo "== xbench_cluster"
timeout 120 ./scripts/xbench_cluster > gpurun_out/xbench_cluster.log 2>&1; echo "xbench rc=$?"
head -16 gpurun_out/xbench_cluster.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
echo "== pytest golden"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "golden" > gpurun_out/pytest_golden.log 2>&1; echo "pytest golden rc=$?" | tee -a gpurun_out/pytest_golden.log
tail -25 gpurun_out/pytest_golden.log
echo "== sweep"
timeout 900 python scripts/sweep.py cfg2:T=2000,WN_PROF=1 cfg2:T=2000,WN_ENGINE=5 cfg2:T=2000,B=8 cfg2:T=2000,B=4 cfg2:T=2000,B=2 cfg1:T=2000 cfg3:T=2000 cfg5:T=2000 cfg2:T=2000,WN_CLUSTER=16 cfg2:T=2000,WN_CLUSTER=4 > gpurun_out/sweep_r2a.log 2>&1; echo "sweep rc=$?"
cat gpurun_out/sweep_r2a.log
echo "== pytest all"
timeout ${PYTEST_TIMEOUT:-900} python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
