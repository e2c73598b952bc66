#!/usr/bin/env bash
# Round 2, GPU session 20: full ncu capture of ONE whole C2 step of the library's kernels (per-kernel DRAM traffic), source-level capture
# of the four heaviest kernels, full-size parity incl. the new C3-vs-oracle test.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s20_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/s20_pytest_gpu.log 2>&1
echo "pytest fullsize + parity: exit $?" | tee gpurun_out/s20_status.txt
tail -3 gpurun_out/s20_pytest_gpu.log
timeout 1200 ncu --set full --clock-control none -k regex:^k_ -s 110 -c 55 -o gpurun_out/s20_full_step python scripts/prof_step.py 3 > gpurun_out/s20_ncu_full.log 2>&1
echo "ncu full (one step, library kernels): exit $?" | tee -a gpurun_out/s20_status.txt
ncu -i gpurun_out/s20_full_step.ncu-rep --page raw --csv > gpurun_out/s20_full_step_raw.csv 2> gpurun_out/s20_export.err
rm -f gpurun_out/s20_full_step.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^(k_warp_g1|k_seam_fwd_cluster|k_blend_level_quad|k_blend_l0_tiled)" -s 12 -c 12 -o gpurun_out/s20_top4 python scripts/prof_step.py 3 > gpurun_out/s20_ncu_top4.log 2>&1
echo "ncu source-level, top 4 kernels: exit $?" | tee -a gpurun_out/s20_status.txt
ls -la gpurun_out/s20_top4.ncu-rep
du -sh gpurun_out
