#!/usr/bin/env bash
# Round 2, GPU session 1: new batched seam path + DP formulations.  gpurun --timeout 1500 -- 'bash scripts/gpu_r2_s1.sh'
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s1_build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_seam_more.py -x -q -k "dp_formulations or pair_loop_paths" > gpurun_out/s1_pytest_seam_new.log 2>&1
echo "new seam tests: exit $?" | tee gpurun_out/s1_status.txt
tail -15 gpurun_out/s1_pytest_seam_new.log
timeout 300 python scripts/dp_bench.py > gpurun_out/s1_dp_bench.log 2>&1
echo "dp bench: exit $?" | tee -a gpurun_out/s1_status.txt
cat gpurun_out/s1_dp_bench.log | tail -12
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s1_pytest_gpu.log 2>&1
echo "pytest -m gpu: exit $?" | tee -a gpurun_out/s1_status.txt
tail -15 gpurun_out/s1_pytest_gpu.log
for v in 1 0; do
  IS_DP_VARIANT=$v IS_SEAM_DEBUG=1 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/s1_bench_dbg_v$v.json 2> gpurun_out/s1_bench_dbg_v$v.err
  echo "bench debug v$v: exit $?" | tee -a gpurun_out/s1_status.txt
  grep "seam batch" gpurun_out/s1_bench_dbg_v$v.err | tail -12
  IS_DP_VARIANT=$v timeout 400 python bench.py --no-cpu-baseline > gpurun_out/s1_bench_v$v.json 2> gpurun_out/s1_bench_v$v.err
  echo "bench v$v: exit $?" | tee -a gpurun_out/s1_status.txt
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/s1_bench_v$v.json").read().strip().splitlines()[-1])
    print("v$v", {k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")}, d.get("e2e"), d.get("stage_ms"))
    for k in d.get("top_kernels", [])[:14]: print("   ", k)
except Exception as e:
    print("no bench line", e)
PY
done
