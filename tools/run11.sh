cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "SGB_DF_MAP=0 SGB_DF_MAP=1 SGB_DF_MAP=2 SGB_DF_MAP=0,SGB_CHAIN_TMA=0 SGB_DF_MAP=0 SGB_DF_MAP=1" | timeout 200 python tools/time_stages.py 1000000 1920 1080 256 8 2>&1 | grep -v Warning | tee gpurun_out/r02_ab11.txt
