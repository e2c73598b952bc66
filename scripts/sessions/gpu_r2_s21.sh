#!/usr/bin/env bash
# Round 2, GPU session 21: four-channel seam costs, level blend without the no-op clamps / divisions, warp kernel two rows ahead.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s21_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/s21_pytest_gpu.log 2>&1
echo "pytest gpu: exit $?" | tee gpurun_out/s21_status.txt
tail -4 gpurun_out/s21_pytest_gpu.log
for wl in c2 c3; do
  timeout 600 python bench.py --no-cpu-baseline --steps 10 --workload $wl > gpurun_out/s21_bench_$wl.json 2> gpurun_out/s21_bench_$wl.err
  echo "bench $wl: exit $?" | tee -a gpurun_out/s21_status.txt
  python scripts/bench_brief.py gpurun_out/s21_bench_$wl.json 12
done
