"""Stage timings of K2 (1M Gaussians, SH RGB + depth, 1080p, forward only, and forward+backward)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from semantic_gaussians_b200 import _lib
from semantic_gaussians_b200.scene_synth import make_scene, orbit_cameras
from util import dev_cam, dev_scene, run_ours
dev = torch.device("cuda:0")
scene = make_scene(1000000, 0, sh=True); cams = orbit_cameras(8, 1920, 1080)
ctx = _lib.ctx_for(0, torch.cuda.current_stream(dev).cuda_stream)
for bwd in (False, True):
    sc = dev_scene(scene, dev, requires_grad=bwd); bg = torch.zeros(3, device=dev); dL = torch.randn(3, 1080, 1920, device=dev)
    def step(i):
        o = run_ours("rgbd", sc, dev_cam(cams[i % 8], dev), bg, use_features=False)
        if bwd:
            o["color"].backward(dL)
            for v in sc.values():
                if v is not None: v.grad = None
    for i in range(3): step(i)
    torch.cuda.synchronize(); _lib.profile_enable(ctx, True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10): step(i)
    e1.record(); torch.cuda.synchronize()
    st = _lib.profile_read(ctx); _lib.profile_enable(ctx, False)
    print("bwd" if bwd else "fwd", f"total {e0.elapsed_time(e1)/10:.3f} ms |", " ".join(f"{k}={v[0]/max(v[1],1):.3f}" for k, v in st.items() if v[1]), flush=True)
