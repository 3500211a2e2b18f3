set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=900 2>&1 | tail -15 > gpurun_out/r02_pytest2.log
cat gpurun_out/r02_pytest2.log
timeout 400 python bench.py --steps 16 --warmup 3 > gpurun_out/r02_bench_k3.json 2> gpurun_out/r02_bench_k3.err
tail -c 400 gpurun_out/r02_bench_k3.err
SGB_CHAIN_V3=1 timeout 200 python bench.py --steps 16 --warmup 3 --no-baselines --quick > gpurun_out/r02_bench_k3_chainv3.json 2> gpurun_out/r02_bench_k3_chainv3.err
timeout 200 python bench.py --config K2 > gpurun_out/r02_bench_k2.json 2> gpurun_out/r02_bench_k2.err
tail -c 400 gpurun_out/r02_bench_k2.err
SGB_K4_VIEWS=8 timeout 400 python bench.py --config K4 --steps 2 > gpurun_out/r02_bench_k4_v8.json 2> gpurun_out/r02_bench_k4_v8.err
tail -c 600 gpurun_out/r02_bench_k4_v8.err
SGB_K5_VIEWS=60 timeout 400 python bench.py --config K5 --steps 2 > gpurun_out/r02_bench_k5_v60.json 2> gpurun_out/r02_bench_k5_v60.err
tail -c 600 gpurun_out/r02_bench_k5_v60.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/r02_bench_ref_k3.json 2> gpurun_out/r02_bench_ref_k3.err
echo "default" | timeout 300 ncu --set full --clock-control none --import-source on -k regex:chain_backward -s 2 -c 1 -o gpurun_out/prof_r02_chain_warp -f python tools/time_stages.py 1000000 1920 1080 256 1 > gpurun_out/ncu_r02_chain.log 2>&1
ls -la gpurun_out | tail -20
