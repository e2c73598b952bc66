#!/usr/bin/env bash
# Round 2, GPU session 23 (8 GPUs): weak-scaling strip (48 x (4000 x 6000)) and C4 again after the host-pool sizing by cores per rank
# and the deferred verdict; N = 4 beside it.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s23_build.log 2>&1
python -c "import os; print('host cpus', os.cpu_count(), 'affinity', len(os.sched_getaffinity(0)))" | tee gpurun_out/s23_status.txt
for spec in "8 c2" "8 c4" "4 c2"; do
  set -- $spec
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29551"
  timeout 600 $TR bench.py --gpus $1 --steps 5 --warmup 3 --workload $2 > gpurun_out/s23_bench_$2_n$1.json 2> gpurun_out/s23_bench_$2_n$1.err
  echo "bench $2 N=$1: exit $?" | tee -a gpurun_out/s23_status.txt
  python scripts/bench_brief.py gpurun_out/s23_bench_$2_n$1.json 2
done
IS_SHARD_DEBUG=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --steps 3 --warmup 3 --workload c2 --no-parity-check > gpurun_out/s23_bench_c2_n8_laps.json 2> gpurun_out/s23_bench_c2_n8_laps.err
grep "shard rank [034]\]" gpurun_out/s23_bench_c2_n8_laps.json | tail -3
