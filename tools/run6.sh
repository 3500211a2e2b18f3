set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
PT="python -m pytest -q --timeout=150 --timeout-method=thread"
timeout 200 $PT tests/test_fusion_gpu.py tests/test_parity_sizes_gpu.py -m gpu -k "fusion or k5" 2>&1 | tail -6 > gpurun_out/r02_pt6_fusion.log; cat gpurun_out/r02_pt6_fusion.log
SGB_BLEND_MMA=1 timeout 300 $PT tests/test_parity_gpu.py tests/test_parity_sizes_gpu.py tests/test_batch_gpu.py -m gpu -k "backward or channel_forward or nonfinite or k3_full or tiles or not_multiple or batch_equals" 2>&1 | tail -25 > gpurun_out/r02_pt6_mma.log; cat gpurun_out/r02_pt6_mma.log
SGB_BLEND_MMA=1 timeout 150 python bench.py --steps 16 --warmup 3 --no-baselines --quick > gpurun_out/r02_b6_k3_mma.json 2> gpurun_out/r02_b6_k3_mma.err
tail -c 300 gpurun_out/r02_b6_k3_mma.err
python -c "
import json;d=json.load(open('gpurun_out/r02_b6_k3_mma.json'));print('mma',d['ms_per_step'],{k:round(x,3) for k,x in d['stage_ms'].items()})"
timeout 200 python bench.py --config K5 --steps 3 > gpurun_out/r02_bench_k5_n1.json 2> gpurun_out/r02_bench_k5_n1.err
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_k5_n1.json'));print('k5',d['ms_per_step'],d['stage_ms'],d['roofline']['frac'])"
echo "SGB_BLEND_MMA=1" | SGB_BLEND_MMA=1 timeout 250 ncu --set full --clock-control none --import-source on -k regex:mma_kernel -s 3 -c 3 -o gpurun_out/prof_r02_mma -f python tools/time_stages.py 1000000 1920 1080 256 1 > gpurun_out/ncu_r02_mma.log 2>&1
tail -3 gpurun_out/ncu_r02_mma.log
ls -la gpurun_out | tail -8
