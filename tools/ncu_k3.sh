#!/bin/bash
# ncu captures for the K3 workload (run under gpurun, one GPU).  Outputs under gpurun_out/.
#   usage: tools/ncu_k3.sh [kernel-regex ...]   (default: launch list + the blend kernels)
mkdir -p gpurun_out
echo "default" | ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_k3.csv python tools/time_stages.py 1000000 1920 1080 256 1 > gpurun_out/ncu_launch.log 2>&1
for k in "${@:-blend_forward chain_backward dfeature alpha_pass}"; do
  for name in $k; do
    echo "default" | ncu --set full --clock-control none --import-source on -k regex:$name -s 2 -c 1 -o gpurun_out/prof_$name -f python tools/time_stages.py 1000000 1920 1080 256 1 > gpurun_out/ncu_$name.log 2>&1
  done
done
ls -la gpurun_out/ | head -30
