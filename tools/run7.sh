set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
PT="python -m pytest -q --timeout=150 --timeout-method=thread"
timeout 120 $PT tests/test_fusion_gpu.py -m gpu 2>&1 | tail -4 > gpurun_out/r02_pt7_fusion.log; cat gpurun_out/r02_pt7_fusion.log
SGB_BLEND_MMA=1 timeout 300 $PT tests/test_parity_gpu.py tests/test_parity_sizes_gpu.py -m gpu -k "backward or channel_forward or nonfinite or k3_full or k4_full or not_multiple" 2>&1 | tail -12 > gpurun_out/r02_pt7_mma.log; cat gpurun_out/r02_pt7_mma.log
SGB_BLEND_MMA=1 timeout 150 python bench.py --steps 24 --warmup 3 --no-baselines --quick > gpurun_out/r02_bench_k3_n1_mma.json 2> gpurun_out/r02_bench_k3_n1_mma.err
tail -c 300 gpurun_out/r02_bench_k3_n1_mma.err
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_k3_n1_mma.json'));print('mma',d['ms_per_step'],d['e2e']['ms_per_step'],{k:round(x,3) for k,x in d['stage_ms'].items()})"
timeout 200 python bench.py --config K5 --steps 3 > gpurun_out/r02_bench_k5_n1.json 2> gpurun_out/r02_bench_k5_n1.err
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_k5_n1.json'));print('k5',d['ms_per_step'],d['stage_ms'],d['roofline']['frac'],d['e2e']['ms_per_step'])"
