#!/bin/bash
mkdir -p gpurun_out
export WN_TIMEOUT_MS=4000
timeout 600 python scripts/sweep.py cfg2:T=3000 cfg5:T=3000 cfg2:T=2000,B=4 > gpurun_out/sweep_r2j.log 2>&1; echo "sweep rc=$?"; cut -c1-170 gpurun_out/sweep_r2j.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_upsample.py -m gpu -q -x -k "golden or concurrent or softmax or upsampler" > gpurun_out/pytest_c.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_c.log
export WN_TIMEOUT_MS=120000
timeout 300 compute-sanitizer --tool synccheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_synccheck_smoke.log 2>&1; echo "synccheck smoke rc=$?"
grep -v "Host Frame\|Saved host" gpurun_out/r2_sanitizer_synccheck_smoke.log | head -12 | cut -c1-200; tail -3 gpurun_out/r2_sanitizer_synccheck_smoke.log
timeout 300 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_racecheck_smoke.log 2>&1; echo "racecheck smoke rc=$?"
grep "Race reported\|SUMMARY\|smoke:" gpurun_out/r2_sanitizer_racecheck_smoke.log | cut -c1-200
