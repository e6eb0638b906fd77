#!/bin/bash
mkdir -p gpurun_out
export WN_TIMEOUT_MS=${WN_TIMEOUT_MS:-3000}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,temperature.gpu,power.draw --format=csv > gpurun_out/gpu2.txt 2>&1
timeout 600 python scripts/lean_check.py > gpurun_out/lean_check.log 2>&1; echo "lean_check rc=$?"; tail -4 gpurun_out/lean_check.log
PREV=wavenet_vocoder_b200/libwn_prev.so
timeout 1500 python scripts/sweep.py cfg2:T=3000,WN_LIB_PATH=$PREV cfg2:T=3000,WN_LEAN=0 cfg2:T=3000 cfg2:T=3000,WN_LIB_PATH=$PREV \
   cfg2:T=3000,WN_PROF=1 cfg2:T=3000,WN_GATE_CYCLES=400 cfg2:T=3000,WN_GATE_CYCLES=800 cfg2:T=3000,WN_FAST_GATE=1 \
   cfg5:T=3000,WN_LIB_PATH=$PREV cfg5:T=3000 cfg5:T=3000,WN_FAST_GATE=1 cfg2:T=3000,WN_LIB_PATH=$PREV,WN_PROF=1 > gpurun_out/sweep_r2f.log 2>&1; echo "sweep rc=$?"
grep -v "WN_PROF" gpurun_out/sweep_r2f.log | cut -c1-150
grep -A12 "WN_PROF=1" gpurun_out/sweep_r2f.log | grep -v "^cfg" | head -26 | cut -c1-120
cat gpurun_out/gpu2.txt
