set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
PT="python -m pytest -q --timeout=150 --timeout-method=thread"
timeout 200 $PT tests/test_fusion_gpu.py tests/test_parity_sizes_gpu.py -m gpu -k "fusion or k5 or nonfinite" 2>&1 | tail -6 > gpurun_out/r02_pt5_fusion.log; cat gpurun_out/r02_pt5_fusion.log
timeout 200 $PT tests/test_parity_gpu.py -m gpu -k "backward or tiles or not_multiple" 2>&1 | tail -6 > gpurun_out/r02_pt5_bwd.log; cat gpurun_out/r02_pt5_bwd.log
SGB_BLEND_MMA=1 timeout 150 $PT tests/test_parity_gpu.py tests/test_parity_sizes_gpu.py -k "channel_forward_vs_reference or nonfinite or k3_full" 2>&1 | tail -8 > gpurun_out/r02_pt5_mma.log; cat gpurun_out/r02_pt5_mma.log
SGB_BLEND_MMA=1 timeout 150 python bench.py --steps 16 --warmup 3 --no-baselines --quick > gpurun_out/r02_b5_k3_mma.json 2> gpurun_out/r02_b5_k3_mma.err
python -c "
import json;d=json.load(open('gpurun_out/r02_b5_k3_mma.json'));print('mma',d['ms_per_step'],{k:round(x,3) for k,x in d['stage_ms'].items()})"
timeout 400 python bench.py --steps 24 --warmup 3 > gpurun_out/r02_bench_k3_n1.json 2> gpurun_out/r02_bench_k3_n1.err
tail -c 300 gpurun_out/r02_bench_k3_n1.err
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_k3_n1.json'));print('k3',d['ms_per_step'],d['e2e']['ms_per_step'],{k:round(x,3) for k,x in d['stage_ms'].items()});print(d['reference_cuda']);print(d['cpu_baseline']);print(d['cpu_preprocess_torch'])"
timeout 200 python bench.py --config K2 > gpurun_out/r02_bench_k2_n1.json 2> gpurun_out/r02_bench_k2_n1.err
timeout 400 python bench.py --config K4 --steps 3 > gpurun_out/r02_bench_k4_n1.json 2> gpurun_out/r02_bench_k4_n1.err
tail -c 300 gpurun_out/r02_bench_k4_n1.err
timeout 300 python bench.py --config K5 --steps 3 > gpurun_out/r02_bench_k5_n1.json 2> gpurun_out/r02_bench_k5_n1.err
tail -c 600 gpurun_out/r02_bench_k5_n1.err
python - <<'PY'
import json
for f in ("k2_n1","k4_n1","k5_n1"):
    try:
        d=json.load(open(f"gpurun_out/r02_bench_{f}.json"))
        print(f, d["ms_per_step"], d["e2e"]["ms_per_step"], {k:round(x,3) for k,x in d["stage_ms"].items()}, d.get("batched_vs_loop"), d["config"].get("N_vis_per_view"), d["roofline"]["frac"], d.get("cpu_baseline"))
    except Exception as e:
        print(f, "ERR", e)
PY
