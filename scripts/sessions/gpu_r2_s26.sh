#!/usr/bin/env bash
# Round 2, GPU session 26: the committed state as the driver will run it -- build, smoke, full GPU suite, default bench line (with CPU legs),
# reference arm.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/s26_build_smoke.log 2>&1
echo "build + smoke: exit $?" | tee gpurun_out/s26_status.txt
tail -1 gpurun_out/s26_build_smoke.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/s26_pytest_gpu.log 2>&1
echo "pytest gpu: exit $?" | tee -a gpurun_out/s26_status.txt
tail -3 gpurun_out/s26_pytest_gpu.log
/usr/bin/time -v timeout 900 python bench.py > gpurun_out/s26_bench_default.json 2> gpurun_out/s26_bench_default.err
echo "bench (defaults): exit $?" | tee -a gpurun_out/s26_status.txt
grep "Elapsed (wall clock)" gpurun_out/s26_bench_default.err
python scripts/bench_brief.py gpurun_out/s26_bench_default.json 4
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s26_bench_reference.json 2> gpurun_out/s26_bench_reference.err
echo "bench --impl reference: exit $?" | tee -a gpurun_out/s26_status.txt
cut -c1-200 gpurun_out/s26_bench_reference.json
