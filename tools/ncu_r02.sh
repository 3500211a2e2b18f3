#!/bin/bash
# Round-2 profile refresh (run under gpurun, ONE GPU): launch list of one K3 step + ncu --set full of each hot kernel.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
K3="python tools/time_stages.py 1000000 1920 1080 256 1"
echo "default" | timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_k3.csv $K3 > gpurun_out/ncu_launch.log 2>&1
for name in blend_forward_tma chain_backward_warp dfeature_gemm alpha_pass DeviceRadixSortOnesweep; do
  echo "default" | timeout 240 ncu --set full --clock-control none --import-source on -k regex:$name -s 3 -c 1 -o gpurun_out/prof_r02_$name -f $K3 > gpurun_out/ncu_$name.log 2>&1
  tail -1 gpurun_out/ncu_$name.log
done
timeout 240 ncu --set full --clock-control none --import-source on -k regex:blend_forward_kernel -s 3 -c 1 -o gpurun_out/prof_r02_k2_blend_forward -f python tools/time_k2.py > gpurun_out/ncu_k2.log 2>&1
timeout 240 ncu --set full --clock-control none --import-source on -k regex:fusion_gather_sorted -s 6 -c 1 -o gpurun_out/prof_r02_k5_fusion_gather -f python tools/time_fusion.py > gpurun_out/ncu_k5.log 2>&1
ls -la gpurun_out/*.ncu-rep
