cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python bench.py --steps 6 --warmup 3 --quick --no-baselines > gpurun_out/chk_n1.out 2> gpurun_out/chk_n1.err; echo "lines: $(wc -l < gpurun_out/chk_n1.out)"; head -c 150 gpurun_out/chk_n1.out; echo
NCCL_DEBUG=VERSION timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 1 --steps 6 --warmup 3 --quick --no-baselines > gpurun_out/chk_t1.out 2> gpurun_out/chk_t1.err; echo "lines: $(wc -l < gpurun_out/chk_t1.out)"; head -c 150 gpurun_out/chk_t1.out; echo; grep -c "NCCL version" gpurun_out/chk_t1.err
timeout 200 python bench.py --config K5 --steps 3 > gpurun_out/r02_bench_k5_n1.json 2> gpurun_out/r02_bench_k5_n1.err; python -c "
import json;d=json.load(open('gpurun_out/r02_bench_k5_n1.json'));print('k5',d['ms_per_step'],d['stage_ms'],d['roofline']['frac'],d['roofline']['traffic'],d['e2e']['ms_per_step'])"
