#!/usr/bin/env bash
# First GPU call of a round: everything that was written without hardware access gets checked in one box session.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_round_open.sh'
# Outputs land in gpurun_out/ (merged back by gpurun): test log, COLOR_GRAD parity log, bench lines, ncu launch list.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/open_build.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_zz_reports.py > gpurun_out/open_pytest_gpu.log 2>&1   # the report scripts run directly below
echo "pytest -m gpu: exit $?" | tee gpurun_out/open_status.txt
timeout 600 python tests/tools/check_color_grad.py > gpurun_out/open_color_grad.log 2>&1
echo "COLOR_GRAD device parity: exit $?" | tee -a gpurun_out/open_status.txt
timeout 600 python tests/tools/check_seam_edge_cases.py > gpurun_out/open_seam_edge_cases.log 2>&1
echo "seam edge cases on the device: exit $?" | tee -a gpurun_out/open_status.txt
timeout 300 python tests/tools/check_linblend_exact.py > gpurun_out/open_linblend_exact.log 2>&1
echo "pair blend exactness report: exit $?" | tee -a gpurun_out/open_status.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/open_smoke.log 2>&1
echo "smoke: exit $?" | tee -a gpurun_out/open_status.txt
timeout 600 python bench.py > gpurun_out/open_bench.json 2> gpurun_out/open_bench.err
echo "bench: exit $?" | tee -a gpurun_out/open_status.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/open_launches.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/open_ncu_bench.log 2>&1
echo "ncu launch list: exit $?" | tee -a gpurun_out/open_status.txt
tail -3 gpurun_out/open_pytest_gpu.log
tail -8 gpurun_out/open_color_grad.log
tail -2 gpurun_out/open_linblend_exact.log
grep -v ': ok' gpurun_out/open_seam_edge_cases.log | tail -12
cat gpurun_out/open_bench.json
