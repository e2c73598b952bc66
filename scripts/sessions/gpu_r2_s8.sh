#!/usr/bin/env bash
# Round 2, GPU session 8: validate run-based seam costs + DP L2 prefetch + D2H turns; seam-stage laps on a device-resident run.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s8_build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/s8_pytest_gpu.log 2>&1
echo "pytest gpu: exit $?" | tee gpurun_out/s8_status.txt
tail -4 gpurun_out/s8_pytest_gpu.log
timeout 300 python scripts/dp_bench.py > gpurun_out/s8_dp_bench.log 2>&1
echo "dp bench: exit $?" | tee -a gpurun_out/s8_status.txt
cut -c1-260 gpurun_out/s8_dp_bench.log | head -5
timeout 600 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/s8_bench_c2.json 2> gpurun_out/s8_bench_c2.err
echo "bench c2: exit $?" | tee -a gpurun_out/s8_status.txt
python scripts/bench_brief.py gpurun_out/s8_bench_c2.json
IS_SEAM_DEBUG=1 IS_DEBUG_PLAN_TIMING=1 timeout 300 python scripts/prof_step.py 6 > gpurun_out/s8_seam_laps.log 2>&1
echo "seam laps: exit $?" | tee -a gpurun_out/s8_status.txt
tail -60 gpurun_out/s8_seam_laps.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s8_sanitizer_memcheck.log 2>&1
echo "memcheck smoke: exit $?" | tee -a gpurun_out/s8_status.txt
tail -5 gpurun_out/s8_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s8_sanitizer_racecheck.log 2>&1
echo "racecheck smoke: exit $?" | tee -a gpurun_out/s8_status.txt
tail -5 gpurun_out/s8_sanitizer_racecheck.log
