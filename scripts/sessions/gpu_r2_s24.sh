#!/usr/bin/env bash
# Round 2, GPU session 24 (2 GPUs): general per-pair seam path on the clustered DP kernels: full suite (2 devices), C5 share at N = 2.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s24_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/s24_pytest_gpu.log 2>&1
echo "pytest gpu (2 devices): exit $?" | tee gpurun_out/s24_status.txt
tail -4 gpurun_out/s24_pytest_gpu.log
IS_SEAM_PATH=seq timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_seam_more.py -m gpu -q -x > gpurun_out/s24_pytest_seq.log 2>&1
echo "pytest parity, every pair through the general path (IS_SEAM_PATH=seq): exit $?" | tee -a gpurun_out/s24_status.txt
tail -3 gpurun_out/s24_pytest_seq.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561"
IS_SHARD_DEBUG=1 timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 --workload c5 > gpurun_out/s24_bench_c5_n2.json 2> gpurun_out/s24_bench_c5_n2.err
echo "bench c5 N=2: exit $?" | tee -a gpurun_out/s24_status.txt
python scripts/bench_brief.py gpurun_out/s24_bench_c5_n2.json 6
grep "shard rank 0" gpurun_out/s24_bench_c5_n2.json | tail -1
