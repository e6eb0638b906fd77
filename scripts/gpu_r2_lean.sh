#!/bin/bash
mkdir -p gpurun_out
export WN_TIMEOUT_MS=${WN_TIMEOUT_MS:-3000}
timeout 600 python scripts/lean_check.py > gpurun_out/lean_check.log 2>&1; echo "lean_check rc=$?"; tail -8 gpurun_out/lean_check.log
timeout 1200 python scripts/sweep.py cfg2:T=3000 cfg2:T=3000,WN_LEAN=0 cfg2:T=3000,WN_FAST_GATE=1 cfg2:T=3000,WN_PROF=1 \
   cfg5:T=3000 cfg5:T=3000,WN_LEAN=0 cfg5:T=3000,WN_FAST_GATE=1 cfg2:T=3000,WN_L2_PERSIST=2 > gpurun_out/sweep_r2e.log 2>&1; echo "sweep rc=$?"
cat gpurun_out/sweep_r2e.log | cut -c1-220 | head -60
timeout 600 python -m pytest tests/test_upsample.py tests/test_gpu_parity.py -m gpu -q -x -k "upsampler or concurrent or golden" > gpurun_out/pytest_lean.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_lean.log
