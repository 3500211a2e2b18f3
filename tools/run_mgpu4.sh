# 4-GPU box: K3 weak-scaling bench at N = 4.   usage: gpurun --gpus 4 -- bash tools/run_mgpu4.sh
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 12 --warmup 3 --quick --no-baselines > gpurun_out/r02_bench_k3_n4.json 2> gpurun_out/r02_bench_k3_n4.err
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_k3_n4.json'));print(d['ms_per_step'],d.get('multi_gpu_parity',{}).get('ok'),d.get('exchange'),d.get('per_rank_ms_per_step'))"
