set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
./tools/microbench.bin > gpurun_out/r02_microbench.txt 2>&1
cat gpurun_out/r02_microbench.txt
timeout 1500 python -m pytest tests -m gpu -x -q --timeout=900 2>&1 | tail -40 > gpurun_out/r02_pytest1.log
cat gpurun_out/r02_pytest1.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err
tail -c 3000 gpurun_out/r02_bench_a.json
tail -5 gpurun_out/r02_bench_a.err
