#!/usr/bin/env bash
# Round 2, GPU session 17 (8 GPUs): BASELINE.json configs[3] (24 x (6000 x 8000) strip) and configs[4] (48 x (8000 x 12000) 4 x 12 mosaic)
# sharded by column strip over 8 GPUs, each with the sharded-vs-single-GPU parity flag (seam masks + panorama strip, bit for bit);
# the default weak-scaling line (48 x (4000 x 6000)) beside them.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s17_build.log 2>&1
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541"
for wl in c4 c2 c5; do
  IS_SHARD_DEBUG=1 timeout 900 $TR bench.py --gpus 8 --steps 3 --warmup 3 --workload $wl > gpurun_out/s17_bench_${wl}_n8.json 2> gpurun_out/s17_bench_${wl}_n8.err
  echo "bench $wl N=8: exit $?" | tee -a gpurun_out/s17_status.txt
  python scripts/bench_brief.py gpurun_out/s17_bench_${wl}_n8.json 3
  grep "shard rank 3" gpurun_out/s17_bench_${wl}_n8.json | tail -1
  grep -i "error" gpurun_out/s17_bench_${wl}_n8.err | head -3 | cut -c1-300
done
