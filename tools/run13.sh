cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "default" | timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_k3.csv python tools/time_stages.py 1000000 1920 1080 256 1 > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
timeout 100 ncu --set full --clock-control none --import-source on -k regex:alpha_pass -s 2 -c 1 -o gpurun_out/prof_r02_alpha_pass -f python tools/time_stages.py 1000000 1920 1080 256 1 < /dev/null > gpurun_out/ncu_alpha_pass.log 2>&1
tail -1 gpurun_out/ncu_alpha_pass.log
