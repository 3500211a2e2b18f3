"""torchrun target: where the N-GPU step time goes (K3).  Prints, on rank 0:
  all-reduce of the (P,C) gradient alone; step without exchange; step with the exchange after the backward;
  step with the exchange overlapped (OverlappedFeatureGradReduce); library stage times in the last mode."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.distributed as dist
from semantic_gaussians_b200 import _lib
from semantic_gaussians_b200.distributed import OverlappedFeatureGradReduce, nccl_overlap_options
from semantic_gaussians_b200.scene_synth import make_scene, orbit_cameras
from util import dev_cam, dev_scene, run_ours

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev, pg_options=nccl_overlap_options())
P, C, W, H = 1000000, 256, 1920, 1080
scene = make_scene(P, 0, channels=C); cams = orbit_cameras(8, W, H)
sc = dev_scene(scene, dev, requires_grad=True); bg = torch.zeros(C, device=dev)
dL = torch.randn((C, H, W), device=dev) / (H * W)
ctx = _lib.ctx_for(lr, torch.cuda.current_stream(dev).cuda_stream)
ov = OverlappedFeatureGradReduce(dev)
dcams = [dev_cam(c, dev) for c in cams]


def timed(fn, reps=10, warm=3):
    for i in range(warm): fn(i)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): fn(i)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


buf = torch.zeros((P, C), device=dev)
t_ar = timed(lambda i: dist.all_reduce(buf))


def step(mode):
    def f(i):
        o = run_ours("chn", sc, dcams[(i * world + rank) % 8], bg, use_features=True)
        o["color"].backward(dL)
        g = sc["features"].grad
        if mode == "after": dist.all_reduce(g)
        elif mode == "overlap":
            ov.start(g); ov.finish()
        for v in sc.values():
            if v is not None: v.grad = None
    return f


t_none, t_after = timed(step("none")), timed(step("after"))
_lib.profile_enable(ctx, True)
t_ov = timed(step("overlap"))
st = _lib.profile_read(ctx)
if rank == 0:
    print(f"world {world}: all-reduce 1 GB alone {t_ar:.3f} ms ({P * C * 4 / 1e9 / t_ar * 1e3:.0f} GB/s algbw) | step no-exchange {t_none:.3f} | "
          f"exchange after backward {t_after:.3f} | overlapped {t_ov:.3f} ms")
    print("stages (overlapped):", " ".join(f"{k}={v[0] / max(v[1], 1):.3f}" for k, v in st.items() if v[1]))
dist.destroy_process_group()
