#!/usr/bin/env bash
# Round 2, GPU session 4: new cost / special-point kernels, new bench.py.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s4_build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/s4_pytest_gpu.log 2>&1
echo "pytest -m gpu: exit $?" | tee gpurun_out/s4_status.txt
tail -12 gpurun_out/s4_pytest_gpu.log
IS_SEAM_DEBUG=1 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/s4_bench_dbg.json 2> gpurun_out/s4_bench_dbg.err
grep "seam batch" gpurun_out/s4_bench_dbg.err | tail -9
for wl in c2 c3 c2_8k c1; do
  timeout 600 python bench.py --no-cpu-baseline --workload $wl > gpurun_out/s4_bench_$wl.json 2> gpurun_out/s4_bench_$wl.err
  echo "bench $wl: exit $?" | tee -a gpurun_out/s4_status.txt
  tail -3 gpurun_out/s4_bench_$wl.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/s4_bench_$wl.json").read().strip().splitlines()[-1])
    print("$wl", {k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")}, d.get("stage_ms"))
    e = d.get("e2e") or {}
    print("   e2e", e.get("value"), e.get("ms_per_step"), e.get("one_panorama_at_a_time"), e.get("two_panoramas_in_flight"))
    r = d.get("roofline") or {}
    print("   roofline", r.get("frac"), r.get("kernel_ms_per_step"), r.get("whole_step"))
    if "$wl" == "c2":
        for k in r.get("kernels", []): print("      ", k)
    if d.get("parity"): print("   parity", d["parity"], d.get("cpu_baseline"))
except Exception as e:
    print("no bench line", e)
PY
done
