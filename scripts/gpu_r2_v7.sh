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
tail -5 gpurun_out/pytest_golden.log
echo "== pytest golden v5 (warp remap)"
WN_ENGINE=5 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "golden_teacher or golden_free or block_count" > gpurun_out/pytest_golden_v5.log 2>&1; echo "pytest golden v5 rc=$?" | tee -a gpurun_out/pytest_golden_v5.log
tail -5 gpurun_out/pytest_golden_v5.log
echo "== sweep"
timeout 1500 python scripts/sweep.py cfg2:T=2000,WN_PROF=1 cfg2:T=2000,WN_WARP_REVERSE=0 cfg2:T=2000,WN_DEFER_GATE=0 cfg2:T=2000,WN_WARP_REVERSE=0,WN_DEFER_GATE=0 \
  cfg2:T=2000,WN_ENGINE=5,WN_PROF=1 cfg2:T=2000,WN_ENGINE=5,WN_WARP_REVERSE=0 cfg2:T=4000,WN_ENGINE=5 \
  cfg2:T=2000,WN_EX_SPREAD=0 cfg2:T=2000,WN_POLL_WARPS=8 cfg2:T=2000,WN_POLL_WARPS=4 \
  cfg2:T=2000,B=8 cfg2:T=2000,B=8,WN_POLL_WARPS=8 cfg2:T=2000,B=8,WN_ENGINE=5 cfg2:T=2000,B=4,WN_ENGINE=5 \
  cfg1:T=2000 cfg3:T=2000 cfg5:T=2000 cfg1:T=2000,WN_ENGINE=5 cfg3:T=2000,WN_ENGINE=5 cfg5:T=2000,WN_ENGINE=5 > gpurun_out/sweep_r2c.log 2>&1; echo "sweep rc=$?"
grep -v "WN_PROF -" gpurun_out/sweep_r2c.log
