#!/usr/bin/env bash
# Round 2, GPU session 5 (2 GPUs): strip path of the sharded stitcher on NCCL, e2e pipelining diagnosis.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s5_build.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR scripts/sharded_check.py 800 1200 3 > gpurun_out/s5_check_small.log 2>&1
echo "sharded check small: exit $?" | tee gpurun_out/s5_status.txt
grep -E "sharded:|SHARDED|differs|Error|error" gpurun_out/s5_check_small.log | tail -8
timeout 600 $TR scripts/sharded_check.py 4000 6000 6 > gpurun_out/s5_check_c2.log 2>&1
echo "sharded check 12x(4000x6000): exit $?" | tee -a gpurun_out/s5_status.txt
grep -E "sharded:|SHARDED|differs|Error|error" gpurun_out/s5_check_c2.log | tail -8
timeout 300 $TR scripts/sharded_check.py 800 1200 4 2 > gpurun_out/s5_check_mosaic.log 2>&1
echo "sharded check mosaic: exit $?" | tee -a gpurun_out/s5_status.txt
grep -E "sharded:|SHARDED|differs|Error|error" gpurun_out/s5_check_mosaic.log | tail -8
IS_SHARD_DEBUG=1 timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/s5_bench_n2.json 2> gpurun_out/s5_bench_n2.err
echo "bench N=2: exit $?" | tee -a gpurun_out/s5_status.txt
grep "shard rank 0" gpurun_out/s5_bench_n2.json gpurun_out/s5_bench_n2.err | tail -3
tail -3 gpurun_out/s5_bench_n2.err
python - <<PY
import json
try:
    lines = [l for l in open("gpurun_out/s5_bench_n2.json").read().strip().splitlines() if l.startswith("{")]
    d = json.loads(lines[-1])
    print("N=2", {k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches", "sharded_matches")}, (d.get("e2e") or {}).get("ms_per_step"), d["config"].get("seam_pairs"))
except Exception as e:
    print("no bench line", e)
PY
IS_PIPELINE_DEBUG=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s5_bench_pipe.json 2> gpurun_out/s5_bench_pipe.err
echo "bench pipeline debug: exit $?" | tee -a gpurun_out/s5_status.txt
grep "pipeline" gpurun_out/s5_bench_pipe.err | tail -60
