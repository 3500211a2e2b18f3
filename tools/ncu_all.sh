#!/bin/bash
# Round profile refresh (run under gpurun, one GPU): launch list of one K3 step + ncu --set full of each hot kernel.
mkdir -p gpurun_out
echo "default" | ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_k3.csv python tools/time_stages.py 1000000 1920 1080 256 1 > gpurun_out/ncu_launch.log 2>&1
for name in blend_forward_tma chain_backward dfeature alpha_pass; do
  echo "default" | timeout 200 ncu --set full --clock-control none --import-source on -k regex:$name -s 2 -c 1 -o gpurun_out/prof_$name -f python tools/time_stages.py 1000000 1920 1080 256 1 > gpurun_out/ncu_$name.log 2>&1
done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:semantic_head -s 3 -c 1 -o gpurun_out/prof_head -f python tools/head_only.py > gpurun_out/ncu_head.log 2>&1
ls -la gpurun_out/*.ncu-rep
