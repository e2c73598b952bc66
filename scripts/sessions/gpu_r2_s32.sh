#!/usr/bin/env bash
# Round 2, GPU session 32 (the last 4.8 GPU-minutes of the round, spent as eight calls of 25-60 s): the rows widened around the path.
# What each call ran and what it showed:
#   1  pytest tests/test_gpu_projectors.py tests/test_imgio.py -m gpu      8 passed (plane / fisheye / stereographic warps and pipelines,
#                                                                           remap on extreme maps, bitmap files through the C ABI)
#   2  pytest tests/test_gpu_orb.py -m gpu                                  first case FAILED: 351 key points against the oracle's 337
#   3  python scripts/orb_debug.py gpu                                      pyramid / blur / row sums equal to the host emulation, score map
#                                                                           not: ptxas folded max(min d, -(max d)) into VIMNMX3 without the
#                                                                           negation (profiles/r2_ptxas_vimnmx3_neg.md)
#   4  pytest tests/test_gpu_orb.py -m gpu   (score rewritten)              4 of 5 passed; the fifth image had no corners (test fixed)
#   5  pytest tests/test_gpu_zz_orb.py tests/test_gpu_parity.py -k "orb or warp or pipeline_end_to_end"      12 passed
#   6  python scripts/orb_bench.py                                          141 ms per 24 MP image, == cv2.ORB; kernels 2.6 ms
#   7  python scripts/orb_bench.py   (counting sort, 8-byte records, laps)  66.7 ms
#   8  python scripts/orb_bench.py   (per-level work on the host pool)      29.1 ms (profiles/r2_orb_bench.json); cv2.ORB 370 ms
# To repeat the lot on one box:
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s32_build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_projectors.py tests/test_gpu_zz_orb.py tests/test_imgio.py -m gpu -q -x > gpurun_out/s32_pytest_widened.log 2>&1
echo "pytest widened rows: exit $?" | tee gpurun_out/s32_status.txt
tail -3 gpurun_out/s32_pytest_widened.log
IS_ORB_LAPS=1 timeout 200 python scripts/orb_bench.py > gpurun_out/s32_orb_bench.json 2> gpurun_out/s32_orb_laps.log
tail -1 gpurun_out/s32_orb_bench.json
