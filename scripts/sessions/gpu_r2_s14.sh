#!/usr/bin/env bash
# Round 2, GPU session 14: host-only validation of the seam waves (verdicts cross-checked against the device validation on the whole suite).
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s14_build.log 2>&1
IS_SEAM_CHECK_BOTH=1 timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/s14_pytest_both.log 2>&1
echo "pytest gpu, host + device validation cross-checked: exit $?" | tee gpurun_out/s14_status.txt
tail -4 gpurun_out/s14_pytest_both.log
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/s14_pytest_gpu.log 2>&1
echo "pytest gpu: exit $?" | tee -a gpurun_out/s14_status.txt
tail -4 gpurun_out/s14_pytest_gpu.log
for wl in c2 c3; do
  timeout 600 python bench.py --no-cpu-baseline --steps 8 --workload $wl > gpurun_out/s14_bench_$wl.json 2> gpurun_out/s14_bench_$wl.err
  echo "bench $wl: exit $?" | tee -a gpurun_out/s14_status.txt
  python scripts/bench_brief.py gpurun_out/s14_bench_$wl.json 4
done
IS_SEAM_DEBUG=1 timeout 300 python scripts/prof_step.py 6 > gpurun_out/s14_seam_laps.log 2>&1
grep "seam batch" gpurun_out/s14_seam_laps.log | tail -9
