#!/bin/bash
# ncu captures for the K3 workload (run under gpurun, one GPU).  Outputs under gpurun_out/.
#   1. launch list with per-launch device time (shares, not absolutes)
#   2. --set full capture of the forward and backward blend kernels
set -x
mkdir -p gpurun_out
echo "default" | ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_k3.csv python tools/time_stages.py 1000000 1920 1080 256 1 > gpurun_out/ncu_launch.log 2>&1
echo "default" | ncu --set full --clock-control none --import-source on -k regex:blend_forward -s 2 -c 1 -o gpurun_out/prof_fwd -f python tools/time_stages.py 1000000 1920 1080 256 1 > gpurun_out/ncu_fwd.log 2>&1
echo "default" | ncu --set full --clock-control none --import-source on -k regex:blend_backward -s 2 -c 1 -o gpurun_out/prof_bwd -f python tools/time_stages.py 1000000 1920 1080 256 1 > gpurun_out/ncu_bwd.log 2>&1
ls -la gpurun_out/
