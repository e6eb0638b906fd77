#!/bin/bash
# GPU contact of engine v7: smoke, golden parity, timing sweep over the exchange knobs, then the full GPU suite.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
export WN_TIMEOUT_MS=${WN_TIMEOUT_MS:-3000}
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
echo "== pytest golden"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "golden or block_count or batch_tiles" > gpurun_out/pytest_golden.log 2>&1; echo "pytest golden rc=$?" | tee -a gpurun_out/pytest_golden.log
tail -25 gpurun_out/pytest_golden.log
echo "== sweep"
timeout 1500 python scripts/sweep.py cfg2:T=2000,WN_PROF=1 cfg2:T=2000,WN_ENGINE=5 \
  cfg2:T=2000,WN_EX_SPREAD=0 cfg2:T=2000,WN_EX_SPREAD=3 cfg2:T=2000,WN_EX_SPREAD=2 \
  cfg2:T=2000,WN_POLL_WARPS=8 cfg2:T=2000,WN_POLL_WARPS=8,WN_PROF=1 cfg2:T=2000,WN_POLL_WARPS=4 \
  cfg2:T=2000,WN_GATE_CYCLES=600 cfg2:T=2000,WN_GATE_CYCLES=1000 cfg2:T=2000,WN_GATE_CYCLES=1400 \
  cfg2:T=2000,WN_BACKOFF_NS=100 cfg2:T=2000,WN_RING_SLOTS=3 cfg2:T=2000,WN_RING_SLOTS=2 \
  cfg2:T=2000,B=8 cfg2:T=2000,B=4 cfg2:T=2000,B=2 cfg2:T=2000,B=8,WN_POLL_WARPS=8 cfg1:T=2000 cfg3:T=2000 cfg5:T=2000 > gpurun_out/sweep_r2b.log 2>&1; echo "sweep rc=$?"
cat gpurun_out/sweep_r2b.log
echo "== pytest all"
timeout ${PYTEST_TIMEOUT:-1200} python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
