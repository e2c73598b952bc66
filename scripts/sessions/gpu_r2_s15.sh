#!/usr/bin/env bash
# Round 2, GPU session 15 (2 GPUs): rehearsal of the C4 / C5 workloads (6000x8000 strip, 8000x12000 4-row mosaic) on the sharded path,
# sharded multi-GPU parity test, default weak-scaling line at N = 2.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s15_build.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521"
for wl in c4 c5 c2; do
  IS_SHARD_DEBUG=1 timeout 900 $TR bench.py --gpus 2 --steps 4 --warmup 3 --workload $wl > gpurun_out/s15_bench_${wl}_n2.json 2> gpurun_out/s15_bench_${wl}_n2.err
  echo "bench $wl N=2: exit $?" | tee -a gpurun_out/s15_status.txt
  python scripts/bench_brief.py gpurun_out/s15_bench_${wl}_n2.json 3
  grep "shard rank 0" gpurun_out/s15_bench_${wl}_n2.err | tail -2
  tail -2 gpurun_out/s15_bench_${wl}_n2.err | cut -c1-300
done
timeout 600 python -m pytest tests -m gpu -q -x -k "shard or multi" > gpurun_out/s15_pytest_multi.log 2>&1
echo "pytest multi-GPU tests: exit $?" | tee -a gpurun_out/s15_status.txt
tail -3 gpurun_out/s15_pytest_multi.log
