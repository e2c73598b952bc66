#!/usr/bin/env bash
# One box session that checks the committed state the way the driver does: build, smoke, the GPU suite, the default bench line, the
# reference arm and an ncu launch list of three device-resident steps.
#   gpurun --timeout 1800 -- 'bash scripts/gpu_round_open.sh'
# Outputs land in gpurun_out/ (merged back by gpurun).  The per-session scripts of round 2 are kept under scripts/sessions/.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/open_build_smoke.log 2>&1
echo "build + smoke: exit $?" | tee gpurun_out/open_status.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/open_pytest_gpu.log 2>&1
echo "pytest -m gpu: exit $?" | tee -a gpurun_out/open_status.txt
tail -3 gpurun_out/open_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/open_bench.json 2> gpurun_out/open_bench.err
echo "bench: exit $?" | tee -a gpurun_out/open_status.txt
python scripts/bench_brief.py gpurun_out/open_bench.json 10
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/open_bench_reference.json 2> gpurun_out/open_bench_reference.err
echo "bench --impl reference: exit $?" | tee -a gpurun_out/open_status.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/open_launches.csv python scripts/prof_step.py 3 > gpurun_out/open_ncu_list.log 2>&1
echo "ncu launch list: exit $?" | tee -a gpurun_out/open_status.txt
python scripts/ncu_launch_list.py gpurun_out/open_launches.csv gpurun_out/open_launch_list.md "three device-resident C2 steps" k_ | head -20
