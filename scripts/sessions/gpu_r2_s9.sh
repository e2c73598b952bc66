#!/usr/bin/env bash
# Round 2, GPU session 9: control transfers by kernel (no copy engine), 2 / 3 panoramas in flight end to end.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s9_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/s9_pytest_gpu.log 2>&1
echo "pytest parity+fullsize: exit $?" | tee gpurun_out/s9_status.txt
tail -3 gpurun_out/s9_pytest_gpu.log
for lanes in 3 2 4; do
  timeout 600 python bench.py --no-cpu-baseline --steps 8 --e2e-lanes $lanes > gpurun_out/s9_bench_c2_l$lanes.json 2> gpurun_out/s9_bench_c2_l$lanes.err
  echo "bench c2 lanes=$lanes: exit $?" | tee -a gpurun_out/s9_status.txt
  python scripts/bench_brief.py gpurun_out/s9_bench_c2_l$lanes.json 3
done
IS_COPY_ENGINE_SMALL=1 timeout 600 python bench.py --no-cpu-baseline --steps 8 --e2e-lanes 3 > gpurun_out/s9_bench_c2_ce.json 2> gpurun_out/s9_bench_c2_ce.err
echo "bench c2 lanes=3, control transfers on the copy engines: exit $?" | tee -a gpurun_out/s9_status.txt
python scripts/bench_brief.py gpurun_out/s9_bench_c2_ce.json 3
IS_PIPELINE_DEBUG=1 timeout 600 python bench.py --no-cpu-baseline --steps 4 --e2e-lanes 3 > gpurun_out/s9_bench_dbg.json 2> gpurun_out/s9_bench_dbg.err
grep "pipeline" gpurun_out/s9_bench_dbg.err | tail -64
