#!/usr/bin/env bash
# Round 2, GPU session 18: full GPU suite, the bench lines kept under profiles/ (C2 with the CPU legs, C1, C3, 8K, reference arm),
# ncu launch list and one full capture of a whole C2 step (per-kernel DRAM traffic), compute-sanitizer on smoke.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s18_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/s18_pytest_gpu.log 2>&1
echo "pytest gpu: exit $?" | tee gpurun_out/s18_status.txt
tail -3 gpurun_out/s18_pytest_gpu.log
timeout 900 python bench.py --steps 10 > gpurun_out/s18_bench_c2.json 2> gpurun_out/s18_bench_c2.err
echo "bench c2 (with CPU legs): exit $?" | tee -a gpurun_out/s18_status.txt
python scripts/bench_brief.py gpurun_out/s18_bench_c2.json 20
for wl in c1 c3 c2_8k; do
  timeout 600 python bench.py --no-cpu-baseline --steps 8 --workload $wl > gpurun_out/s18_bench_$wl.json 2> gpurun_out/s18_bench_$wl.err
  echo "bench $wl: exit $?" | tee -a gpurun_out/s18_status.txt
  python scripts/bench_brief.py gpurun_out/s18_bench_$wl.json 2
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s18_bench_reference.json 2> gpurun_out/s18_bench_reference.err
echo "bench --impl reference: exit $?" | tee -a gpurun_out/s18_status.txt
cut -c1-600 gpurun_out/s18_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s18_launches.csv python scripts/prof_step.py 3 > gpurun_out/s18_ncu_list.log 2>&1
echo "ncu launch list: exit $?" | tee -a gpurun_out/s18_status.txt
timeout 1200 ncu --set full --clock-control none --import-source on -s 112 -c 58 -o gpurun_out/s18_full_step python scripts/prof_step.py 3 > gpurun_out/s18_ncu_full.log 2>&1
echo "ncu full (one step): exit $?" | tee -a gpurun_out/s18_status.txt
ls -la gpurun_out/s18_full_step.ncu-rep
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s18_sanitizer_memcheck.log 2>&1
echo "memcheck smoke: exit $?" | tee -a gpurun_out/s18_status.txt
tail -3 gpurun_out/s18_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s18_sanitizer_racecheck.log 2>&1
echo "racecheck smoke: exit $?" | tee -a gpurun_out/s18_status.txt
tail -6 gpurun_out/s18_sanitizer_racecheck.log
