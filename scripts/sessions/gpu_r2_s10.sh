#!/usr/bin/env bash
# Round 2, GPU session 10: DP forward pass over thread-block clusters (DSMEM halo exchange): micro-benchmark, parity, C2 / C3 / 8K lines.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s10_build.log 2>&1
timeout 300 python scripts/dp_bench.py > gpurun_out/s10_dp_bench.log 2>&1
echo "dp bench: exit $?" | tee gpurun_out/s10_status.txt
python - <<PY
import json
for l in open("gpurun_out/s10_dp_bench.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in d.items() if k != "reached"})
    else:
        print(l.rstrip()[:300])
PY
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/s10_pytest_gpu.log 2>&1
echo "pytest gpu: exit $?" | tee -a gpurun_out/s10_status.txt
tail -4 gpurun_out/s10_pytest_gpu.log
for wl in c2 c3 c2_8k; do
  timeout 600 python bench.py --no-cpu-baseline --steps 8 --workload $wl > gpurun_out/s10_bench_$wl.json 2> gpurun_out/s10_bench_$wl.err
  echo "bench $wl: exit $?" | tee -a gpurun_out/s10_status.txt
  python scripts/bench_brief.py gpurun_out/s10_bench_$wl.json 8
done
