# 8-GPU box: K4 / K5 at N = 8 (each time-boxed; K3 at N = 1..8 is the driver's own scaling run).
#   usage: gpurun --gpus 8 -- bash tools/run_mgpu8.sh
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29522 bench.py --gpus 8 --config K4 --steps 3 --warmup 3 > gpurun_out/r02_bench_k4_n8.json 2> gpurun_out/r02_bench_k4_n8.err
tail -c 400 gpurun_out/r02_bench_k4_n8.err
timeout 150 $TR --master-port 29523 bench.py --gpus 8 --config K5 --steps 3 --warmup 3 > gpurun_out/r02_bench_k5_n8.json 2> gpurun_out/r02_bench_k5_n8.err
tail -c 400 gpurun_out/r02_bench_k5_n8.err
timeout 100 $TR --master-port 29521 bench.py --gpus 8 --steps 12 --warmup 3 --quick --no-baselines > gpurun_out/r02_bench_k3_n8.json 2> gpurun_out/r02_bench_k3_n8.err
tail -c 400 gpurun_out/r02_bench_k3_n8.err
python - <<'PY'
import json
for f in ("k4_n8","k5_n8","k3_n8"):
    try:
        s=open(f"gpurun_out/r02_bench_{f}.json").read()
        d=json.loads([l for l in s.splitlines() if l.strip().startswith("{")][0])
        print(f, d["ms_per_step"], d["value"], d.get("exchange"), (d.get("multi_gpu_parity") or {}).get("ok"), d.get("per_rank_ms_per_step"))
    except Exception as e:
        print(f, "ERR", e)
PY
