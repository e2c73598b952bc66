#!/usr/bin/env bash
# Round 2, GPU session 3.  gpurun --timeout 1500 -- 'bash scripts/gpu_r2_s3.sh'
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s3_build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_seam_more.py -x -q > gpurun_out/s3_pytest_seam_more.log 2>&1
echo "seam_more tests: exit $?" | tee gpurun_out/s3_status.txt
tail -12 gpurun_out/s3_pytest_seam_more.log
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_seam_more.py > gpurun_out/s3_pytest_gpu.log 2>&1
echo "pytest -m gpu: exit $?" | tee -a gpurun_out/s3_status.txt
tail -25 gpurun_out/s3_pytest_gpu.log
IS_SEAM_DEBUG=1 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/s3_bench_dbg.json 2> gpurun_out/s3_bench_dbg.err
grep "seam batch" gpurun_out/s3_bench_dbg.err | tail -14
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err
echo "bench: exit $?" | tee -a gpurun_out/s3_status.txt
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/s3_bench.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")}, d.get("e2e"), d.get("stage_ms"))
    for k in d.get("all_kernels", []): print("   ", k)
except Exception as e:
    print("no bench line", e)
PY
timeout 400 python bench.py --no-cpu-baseline --workload c3 --steps 5 > gpurun_out/s3_bench_c3.json 2> gpurun_out/s3_bench_c3.err
echo "bench c3: exit $?" | tee -a gpurun_out/s3_status.txt
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/s3_bench_c3.json").read().strip().splitlines()[-1])
    print("c3", {k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")}, d.get("e2e"), d.get("stage_ms"))
except Exception as e:
    print("no bench line", e)
PY
