"""K5 fusion stage timings on one GPU (2 M Gaussians, 512-ch fp16 640x480 maps), GPU part only."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from semantic_gaussians_b200 import _lib
from semantic_gaussians_b200.fusion import PointCloudToImageMapper
from semantic_gaussians_b200.scene_synth import make_scene, room_cameras
dev = torch.device("cuda:0")
P, C, w, h, nviews = 2_000_000, 512, 640, 480, 6
scene = make_scene(P, 0, kind="room"); cams = room_cameras(nviews, w, h)
fm = torch.from_numpy(np.random.default_rng(0).standard_normal((C, h, w)).astype(np.float16)).to(dev)
xyz = torch.as_tensor(scene.xyz, device=dev); fs = torch.zeros((P, C), device=dev); cnt = torch.zeros(P, device=dev)
depth = None   # the K5 bench rule (`depth: none`, fusion_utils.py:70-72): ~480 k visible Gaussians per view
ctx = _lib.ctx_for(0, torch.cuda.current_stream(dev).cuda_stream)
mappers = [PointCloudToImageMapper([w, h], 0.25, 10, c.intrinsics(), device=dev) for c in cams]
for i in range(nviews): mappers[i].accumulate(cams[i].world_view_transform, xyz, fm, fs, cnt, depth)
torch.cuda.synchronize(); _lib.profile_enable(ctx, True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); e0.record()
for rep in range(3):
    for i in range(nviews): mappers[i].accumulate(cams[i].world_view_transform, xyz, fm, fs, cnt, depth)
e1.record(); torch.cuda.synchronize(); st = _lib.profile_read(ctx)
print(f"K5 fusion {e0.elapsed_time(e1) / (3 * nviews):.3f} ms/view |", " ".join(f"{a}={v[0] / max(v[1], 1):.3f}" for a, v in st.items() if v[1]))
