set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
PT="python -m pytest -q --timeout=150 --timeout-method=thread"
timeout 420 $PT tests/test_parity_gpu.py tests/test_parity_sizes_gpu.py tests/test_batch_gpu.py tests/test_render_api_gpu.py -m gpu -x 2>&1 | tail -12 > gpurun_out/r02_pt8.log; cat gpurun_out/r02_pt8.log
timeout 150 python bench.py --steps 24 --warmup 3 --no-baselines --quick > gpurun_out/r02_bench_k3_n1_try8.json 2> gpurun_out/r02_bench_k3_n1_try8.err
tail -c 300 gpurun_out/r02_bench_k3_n1_try8.err
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_k3_n1_try8.json'));print('k3',d['ms_per_step'],d['e2e']['ms_per_step'],{k:round(x,3) for k,x in d['stage_ms'].items()})"
