# 2-GPU box: multi-GPU parity tests (NCCL) + bench parity check at N = 2.   usage: gpurun --gpus 2 -- bash tools/run_mgpu2.sh
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 400 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout=300 --timeout-method=thread 2>&1 | tail -15 > gpurun_out/r02_pt_mgpu2.log
cat gpurun_out/r02_pt_mgpu2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 16 --warmup 3 --quick > gpurun_out/r02_bench_k3_n2.json 2> gpurun_out/r02_bench_k3_n2.err
tail -c 600 gpurun_out/r02_bench_k3_n2.err
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_k3_n2.json'));print(d['ms_per_step'],d.get('multi_gpu_parity'),d.get('exchange'),d.get('per_rank_ms_per_step'))"
