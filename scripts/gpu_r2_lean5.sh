#!/bin/bash
mkdir -p gpurun_out
export WN_TIMEOUT_MS=${WN_TIMEOUT_MS:-3000}
timeout 600 python scripts/lean_check.py > gpurun_out/lean_check.log 2>&1; echo "lean_check rc=$?"; tail -4 gpurun_out/lean_check.log
timeout 1500 python scripts/sweep.py cfg2:T=3000,WN_LEAN=0 cfg2:T=3000,WN_LEAN=1,WN_FAST_GATE=1 cfg2:T=3000,WN_LEAN=1,WN_FAST_GATE=1,WN_PROF=1 \
   cfg2:T=3000,WN_LEAN=1,WN_FAST_GATE=1,WN_GATE_CYCLES=600 cfg5:T=3000,WN_LEAN=1,WN_FAST_GATE=1 > gpurun_out/sweep_r2i.log 2>&1; echo "sweep rc=$?"
grep -v "WN_PROF" gpurun_out/sweep_r2i.log | cut -c1-170
grep -A12 "WN_PROF=1" gpurun_out/sweep_r2i.log | grep -v "^cfg" | head -13 | cut -c1-120
