#!/usr/bin/env bash
# Round 2, GPU session 12: where the clustered DP forward pass stalls (ncu source counters, one launch).
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s12_build.log 2>&1
IS_DP_CLUSTER=4 IS_DP_CL_RING=16 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_seam_fwd -s 1 -c 1 -o gpurun_out/s12_dp_cl4_r16 python scripts/dp_one.py 1500 4029 5 > gpurun_out/s12_ncu.log 2>&1
echo "ncu cl4 r16: exit $?" | tee gpurun_out/s12_status.txt
IS_DP_CLUSTER=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_seam_fwd -s 1 -c 1 -o gpurun_out/s12_dp_cl1 python scripts/dp_one.py 1500 4029 5 >> gpurun_out/s12_ncu.log 2>&1
echo "ncu cl1: exit $?" | tee -a gpurun_out/s12_status.txt
tail -3 gpurun_out/s12_ncu.log
