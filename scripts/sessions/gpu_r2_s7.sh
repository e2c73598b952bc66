#!/usr/bin/env bash
# Round 2, GPU session 7: reworked fused warp kernel, host link duplex probe, ncu launch list + full captures.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s7_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/s7_pytest_gpu.log 2>&1
echo "pytest parity+fullsize: exit $?" | tee gpurun_out/s7_status.txt
tail -6 gpurun_out/s7_pytest_gpu.log
timeout 120 python scripts/pcie_duplex.py 2>&1 | tee gpurun_out/s7_pcie_duplex.log
IS_PIPELINE_DEBUG=1 timeout 600 python bench.py --no-cpu-baseline --steps 6 > gpurun_out/s7_bench_c2.json 2> gpurun_out/s7_bench_c2.err
echo "bench c2: exit $?" | tee -a gpurun_out/s7_status.txt
grep "pipeline" gpurun_out/s7_bench_c2.err | tail -42
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/s7_bench_c2.json").read().strip().splitlines()[-1])
    print("c2", {k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")}, d.get("stage_ms"))
    e = d.get("e2e") or {}
    print("   e2e", e.get("value"), e.get("ms_per_step"), e.get("one_panorama_at_a_time"), e.get("two_panoramas_in_flight"))
    r = d.get("roofline") or {}
    print("   roofline", r.get("frac"), r.get("kernel_ms_per_step"), r.get("whole_step"))
    for k in r.get("kernels", [])[:12]: print("      ", k)
except Exception as e:
    print("no bench line", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s7_launches.csv python scripts/prof_step.py 3 > gpurun_out/s7_ncu_list.log 2>&1
echo "ncu launch list: exit $?" | tee -a gpurun_out/s7_status.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_warp_g1|k_seam_fwd|k_blend_level_quad|k_blend_l0_tiled|k_pyrdown_images_batch|k_cost_pq_walk|k_row_toggles_batch|k_special_points_batch|k_label_window|k_pyrdown_l0_tiled|k_pyrdown_weights" -s 60 -c 45 -o gpurun_out/s7_full python scripts/prof_step.py 3 > gpurun_out/s7_ncu_full.log 2>&1
echo "ncu full: exit $?" | tee -a gpurun_out/s7_status.txt
ls -la gpurun_out/s7_full.ncu-rep
