cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
PT="python -m pytest -q --timeout=200 --timeout-method=thread"
timeout 800 $PT tests -m gpu 2>&1 | tail -6 > gpurun_out/r02_final_pytest.log; cat gpurun_out/r02_final_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/r02_final_smoke.log
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_k3_n1.json 2> gpurun_out/r02_bench_k3_n1.err; tail -c 400 gpurun_out/r02_bench_k3_n1.err
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_k3_n1.json'));print('k3',d['ms_per_step'],d['e2e']['ms_per_step'],d['roofline'].get('traffic'),d['build'],{k:round(x,3) for k,x in d['stage_ms'].items()})"
timeout 120 python bench.py --config K2 --steps 20 --warmup 3 > gpurun_out/r02_bench_k2_n1.json 2> gpurun_out/r02_bench_k2_n1.err; python -c "
import json;d=json.load(open('gpurun_out/r02_bench_k2_n1.json'));print('k2',d['ms_per_step'],d['e2e']['ms_per_step'])"
timeout 240 python bench.py --config K4 --steps 2 --warmup 1 > gpurun_out/r02_bench_k4_n1.json 2> gpurun_out/r02_bench_k4_n1.err; python -c "
import json;d=json.load(open('gpurun_out/r02_bench_k4_n1.json'));print('k4',d['ms_per_step'],d.get('view_ms'),d['batched_vs_loop'])"
SGB_BLEND_MMA=1 timeout 150 python bench.py --steps 20 --warmup 3 --no-baselines --quick > gpurun_out/r02_bench_k3_n1_mma.json 2> gpurun_out/r02_bench_k3_n1_mma.err
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_k3_n1_mma.json'));print('mma',d['ms_per_step'],d['e2e']['ms_per_step'],{k:round(x,3) for k,x in d['stage_ms'].items()})"
SGB_BLEND_MMA=1 timeout 300 $PT tests/test_parity_gpu.py tests/test_parity_sizes_gpu.py -m gpu -k "backward or channel_forward or nonfinite or k3_full or not_multiple" 2>&1 | tail -3 > gpurun_out/r02_final_pytest_mma.log; cat gpurun_out/r02_final_pytest_mma.log
