"""Runs sgb_semantic_head a few times on a synthetic (C,H,W) image (ncu target). usage: head_only.py [C K H W reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from semantic_gaussians_b200.semantic import semantic_head
C, K, H, W, reps = (int(x) for x in (sys.argv[1:6] + ["256", "21", "1080", "1920", "5"][len(sys.argv) - 1:]))
dev = torch.device("cuda:0")
img = torch.randn((C, H, W), device=dev)
text = torch.nn.functional.normalize(torch.randn(K, C, device=dev), dim=1)
for i in range(reps):
    semantic_head(img, text)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(reps):
    semantic_head(img, text)
e1.record(); torch.cuda.synchronize()
print(f"head C={C} K={K} {W}x{H}: {e0.elapsed_time(e1) / reps:.3f} ms")
