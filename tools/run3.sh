set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
./tools/tc_probe.bin > gpurun_out/r02_tc_probe.txt 2>&1
cat gpurun_out/r02_tc_probe.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 2>&1 | tail -40 > gpurun_out/r02_pytest3.log
cat gpurun_out/r02_pytest3.log
timeout 200 python bench.py --steps 16 --warmup 3 --no-baselines --quick > gpurun_out/r02_bench_k3_b.json 2> gpurun_out/r02_bench_k3_b.err
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_k3_b.json'));print(d['ms_per_step'],d['stage_ms'])"
SGB_K5_VIEWS=24 timeout 500 python bench.py --config K5 --steps 2 --no-baselines > gpurun_out/r02_bench_k5_v24.json 2> gpurun_out/r02_bench_k5_v24.err
tail -c 1500 gpurun_out/r02_bench_k5_v24.err
SGB_BLEND_MMA=1 timeout 150 python -m pytest tests/test_parity_gpu.py -q -k "channel_forward_vs_reference" --timeout=120 2>&1 | tail -25 > gpurun_out/r02_pytest_mma.log
cat gpurun_out/r02_pytest_mma.log
ls -la gpurun_out | tail -12
