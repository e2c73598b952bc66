#!/usr/bin/env bash
# Round 2, GPU session 25 (8 GPUs): C5 (48 x (8000 x 12000) mosaic) again -- general seam path on the clustered DP kernels, waves that follow
# the conflicts, no speculation once a plan's pairs are known to depend on each other.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s25_build.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571"
IS_SHARD_DEBUG=1 timeout 900 $TR bench.py --gpus 8 --steps 3 --warmup 3 --workload c5 > gpurun_out/s25_bench_c5_n8.json 2> gpurun_out/s25_bench_c5_n8.err
echo "bench c5 N=8: exit $?" | tee gpurun_out/s25_status.txt
python scripts/bench_brief.py gpurun_out/s25_bench_c5_n8.json 6
grep "shard rank 3" gpurun_out/s25_bench_c5_n8.json | tail -2
grep -i "error" gpurun_out/s25_bench_c5_n8.err | head -3 | cut -c1-300
