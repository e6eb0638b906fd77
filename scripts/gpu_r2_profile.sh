#!/bin/bash
# Evidence run (round 2): bench lines, ncu launch list + full-set capture of the synthesis kernel, sanitizer logs.
mkdir -p gpurun_out
export WN_TIMEOUT_MS=${WN_TIMEOUT_MS:-4000}
echo "== bench N=1"
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
tail -1 gpurun_out/r2_bench_n1.json | cut -c1-1500
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"
tail -1 gpurun_out/r2_bench_ref.json | cut -c1-600
echo "== ncu launch list (shares, not absolutes)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-config4 --T 4000 > gpurun_out/r2_bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
echo "== ncu full set on the synthesis kernel (T=2000)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:wn_persistent -c 1 -o gpurun_out/r2_prof \
    python scripts/ncu_target.py 2000 1 > gpurun_out/r2_ncu_full.log 2>&1; echo "ncu full rc=$?"
if [ -f gpurun_out/r2_prof.ncu-rep ]; then
  ncu -i gpurun_out/r2_prof.ncu-rep --page raw --csv > gpurun_out/r2_prof_raw.csv 2>/dev/null
  ncu -i gpurun_out/r2_prof.ncu-rep --page details > gpurun_out/r2_prof_details.txt 2>/dev/null
fi
echo "== new tests"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "concurrent or philox or lean" > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_new.log
echo "== compute-sanitizer (smoke case, then config 2 T=300)"
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"
tail -4 gpurun_out/r2_sanitizer_memcheck_smoke.log
export WN_TIMEOUT_MS=120000
timeout 300 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_racecheck_smoke.log 2>&1; echo "racecheck smoke rc=$?"
tail -4 gpurun_out/r2_sanitizer_racecheck_smoke.log
timeout 300 compute-sanitizer --tool synccheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_synccheck_smoke.log 2>&1; echo "synccheck smoke rc=$?"
tail -4 gpurun_out/r2_sanitizer_synccheck_smoke.log
timeout 300 compute-sanitizer --tool racecheck --print-limit 20 python scripts/ncu_target.py 100 1 > gpurun_out/r2_sanitizer_racecheck_cfg2.log 2>&1; echo "racecheck cfg2 rc=$?"
tail -4 gpurun_out/r2_sanitizer_racecheck_cfg2.log
timeout 300 compute-sanitizer --tool synccheck --print-limit 20 python scripts/ncu_target.py 100 1 > gpurun_out/r2_sanitizer_synccheck_cfg2.log 2>&1; echo "synccheck cfg2 rc=$?"
tail -4 gpurun_out/r2_sanitizer_synccheck_cfg2.log
