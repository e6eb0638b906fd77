#!/bin/bash
mkdir -p gpurun_out
export WN_TIMEOUT_MS=20000
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wn7_kernel -c 1 -o gpurun_out/r2_v7_prof python scripts/ncu_target.py 400 1 > gpurun_out/r2_v7_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r2_v7_ncu.log
ncu -i gpurun_out/r2_v7_prof.ncu-rep --page raw --csv > gpurun_out/r2_v7_prof_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_v7_prof.ncu-rep --page source --csv > gpurun_out/r2_v7_prof_source.csv 2>/dev/null
ls -la gpurun_out/r2_v7_prof*
