#!/bin/bash
mkdir -p gpurun_out
export WN_TIMEOUT_MS=4000
timeout 900 python scripts/sweep.py cfg2:T=3000 cfg2:T=3000,WN_LIB_PATH=wavenet_vocoder_b200/libwn_tool.so cfg2:T=3000,WN_LIB_PATH=wavenet_vocoder_b200/libwn_c.so cfg2:T=3000 cfg5:T=3000 > gpurun_out/sweep_r2k.log 2>&1; echo "sweep rc=$?"; cut -c1-170 gpurun_out/sweep_r2k.log
export WN_TIMEOUT_MS=120000
WN_LIB_PATH=wavenet_vocoder_b200/libwn_tool.so timeout 300 compute-sanitizer --tool synccheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_synccheck_smoke.log 2>&1; echo "synccheck smoke (single-site build) rc=$?"
tail -3 gpurun_out/r2_sanitizer_synccheck_smoke.log
WN_LIB_PATH=wavenet_vocoder_b200/libwn_tool.so timeout 300 compute-sanitizer --tool synccheck --print-limit 20 python scripts/ncu_target.py 100 1 > gpurun_out/r2_sanitizer_synccheck_cfg2.log 2>&1; echo "synccheck cfg2 (single-site build) rc=$?"
tail -3 gpurun_out/r2_sanitizer_synccheck_cfg2.log | cut -c1-200
