#!/usr/bin/env bash
# Round 2, GPU session 16 (2 GPUs): pipelined warp kernel + zero-weight skip in the level blend (parity, bench), C5 mosaic rehearsal at N = 2
# after the function-attribute race fix, multi-GPU parity tests.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s16_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/s16_pytest_gpu.log 2>&1
echo "pytest gpu (2 devices visible): exit $?" | tee gpurun_out/s16_status.txt
tail -4 gpurun_out/s16_pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline --steps 8 --workload c2 > gpurun_out/s16_bench_c2.json 2> gpurun_out/s16_bench_c2.err
echo "bench c2: exit $?" | tee -a gpurun_out/s16_status.txt
python scripts/bench_brief.py gpurun_out/s16_bench_c2.json 12
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521"
IS_SHARD_DEBUG=1 timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 --workload c5 > gpurun_out/s16_bench_c5_n2.json 2> gpurun_out/s16_bench_c5_n2.err
echo "bench c5 N=2: exit $?" | tee -a gpurun_out/s16_status.txt
python scripts/bench_brief.py gpurun_out/s16_bench_c5_n2.json 3
grep "shard rank 0" gpurun_out/s16_bench_c5_n2.json | tail -2
grep -i "error" gpurun_out/s16_bench_c5_n2.err | head -5 | cut -c1-300
