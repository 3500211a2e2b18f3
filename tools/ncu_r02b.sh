#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
K3="python tools/time_stages.py 1000000 1920 1080 256 1"
for name in blend_forward_tma chain_backward_warp dfeature_gemm; do
  echo "default" | timeout 200 ncu --set full --clock-control none --import-source on -k regex:$name -s 2 -c 1 -o gpurun_out/prof_r02_$name -f $K3 > gpurun_out/ncu_$name.log 2>&1
  tail -1 gpurun_out/ncu_$name.log
done
ls -la gpurun_out/*.ncu-rep
