#!/usr/bin/env bash
# Round 2, GPU session 19: full GPU suite, the bench lines kept under profiles/ (C2 with the CPU legs, C1, C3, 8K, reference arm),
# ncu launch list and one full capture of a whole C2 step (per-kernel DRAM traffic), compute-sanitizer on smoke.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s19_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/s19_pytest_gpu.log 2>&1
echo "pytest gpu: exit $?" | tee gpurun_out/s19_status.txt
tail -3 gpurun_out/s19_pytest_gpu.log
timeout 900 python bench.py --steps 10 > gpurun_out/s19_bench_c2.json 2> gpurun_out/s19_bench_c2.err
echo "bench c2 (with CPU legs): exit $?" | tee -a gpurun_out/s19_status.txt
python scripts/bench_brief.py gpurun_out/s19_bench_c2.json 20
for wl in c1 c3 c2_8k; do
  timeout 600 python bench.py --no-cpu-baseline --steps 8 --workload $wl > gpurun_out/s19_bench_$wl.json 2> gpurun_out/s19_bench_$wl.err
  echo "bench $wl: exit $?" | tee -a gpurun_out/s19_status.txt
  python scripts/bench_brief.py gpurun_out/s19_bench_$wl.json 2
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s19_bench_reference.json 2> gpurun_out/s19_bench_reference.err
echo "bench --impl reference: exit $?" | tee -a gpurun_out/s19_status.txt
cut -c1-600 gpurun_out/s19_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s19_launches.csv python scripts/prof_step.py 3 > gpurun_out/s19_ncu_list.log 2>&1
echo "ncu launch list: exit $?" | tee -a gpurun_out/s19_status.txt
timeout 1200 ncu --set full --clock-control none --import-source on -s 112 -c 58 -o gpurun_out/s19_full_step python scripts/prof_step.py 3 > gpurun_out/s19_ncu_full.log 2>&1
echo "ncu full (one step): exit $?" | tee -a gpurun_out/s19_status.txt
ls -la gpurun_out/s19_full_step.ncu-rep
# the report itself is too large to travel back (64 MiB limit on gpurun_out): keep its pages as CSV / text and drop it
ncu -i gpurun_out/s19_full_step.ncu-rep --page raw --csv > gpurun_out/s19_full_step_raw.csv 2> /dev/null
for k in k_warp_g1 k_seam_fwd_cluster k_blend_level_quad k_blend_l0_tiled k_cost_pq_walk k_pyrdown_images_batch; do
  ncu -i gpurun_out/s19_full_step.ncu-rep -k regex:$k -c 1 --page source --csv --print-source cuda,sass > gpurun_out/s19_src_$k.csv 2> /dev/null
  ncu -i gpurun_out/s19_full_step.ncu-rep -k regex:$k -c 1 --page details > gpurun_out/s19_details_$k.txt 2> /dev/null
done
rm -f gpurun_out/s19_full_step.ncu-rep
du -sh gpurun_out
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s19_sanitizer_memcheck.log 2>&1
echo "memcheck smoke: exit $?" | tee -a gpurun_out/s19_status.txt
tail -3 gpurun_out/s19_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s19_sanitizer_racecheck.log 2>&1
echo "racecheck smoke: exit $?" | tee -a gpurun_out/s19_status.txt
tail -6 gpurun_out/s19_sanitizer_racecheck.log
