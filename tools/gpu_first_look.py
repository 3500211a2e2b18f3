"""Diagnostic run on the GPU box: bitwise comparison of every stage against the compiled
reference, gradient comparison, and first timings.  Prints a report (not a pass/fail test)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

from oracle import ref as refmod
from semantic_gaussians_b200.scene_synth import make_scene, orbit_cameras
from util import dev_cam, dev_scene, frac_bad, ours_state, rel_err, run_ours

dev = torch.device("cuda:0")
print(torch.cuda.get_device_name(0), "cpu cores", os.cpu_count(), flush=True)


def bits_equal(a, b):
    a, b = a.contiguous().view(torch.int32).reshape(-1), b.contiguous().view(torch.int32).reshape(-1)
    return int((a != b).sum()), a.numel()


def compare_forward(P, W, H, C, use_features, want_depth, seed=0, view=0, nviews=4):
    scene = make_scene(P, seed, sh=not use_features, channels=C if use_features else 0)
    cam = orbit_cameras(nviews, W, H)[view]
    sc, cm = dev_scene(scene, dev), dev_cam(cam, dev)
    st = ours_state(sc, cm, C, use_features=use_features, want_depth=want_depth)
    r = refmod.RefRasterizer("rgbd" if want_depth else "chn")
    bg = torch.zeros(C, device=dev)
    out = r.forward(bg=bg, means3D=sc["means3D"], opacities=sc["opacities"], viewmatrix=cm["viewmatrix"],
                    projmatrix=cm["projmatrix"], campos=cm["campos"], tanfovx=cm["tanfovx"], tanfovy=cm["tanfovy"],
                    W=W, H=H, shs=None if use_features else sc["shs"],
                    colors_precomp=sc["features"] if use_features else None, scales=sc["scales"],
                    rotations=sc["rotations"], num_channels=C)
    vis = out["radii"] > 0
    print(f"--- fwd P={P} {W}x{H} C={C} feat={use_features} depth={want_depth} view={view}: "
          f"P_vis={int(vis.sum())} R_ref={out['R']} R_ours={st['R']}")
    print("  radii mismatches:", int((out["radii"] != st["radii"]).sum()))
    for name in ("depths", "means2D", "conic_opacity", "cov3D", "tiles_touched") + (() if use_features else ("rgb", "clamped")):
        a, b = st[name], r.field(name)
        if a.dtype == torch.uint8:
            bad = int((a[vis] != b[vis]).sum()); n = int(vis.sum()) * 3
        else:
            bad, n = bits_equal(a[vis], b[vis])
        print(f"  {name}: bit mismatches {bad}/{n}")
    if out["R"] == st["R"]:
        print("  point_list mismatches:", int((st["point_list"] != r.field("point_list")).sum()))
        print("  ranges mismatches:", int((st["ranges"] != r.field("ranges")).sum()))
    print("  n_contrib mismatches:", int((st["n_contrib"] != r.field("n_contrib")).sum()), "/", W * H)
    print("  final_T bit mismatches:", bits_equal(st["final_T"], r.field("accum_alpha")))
    print("  color bit mismatches:", bits_equal(st["color"], out["color"]), "rel err", rel_err(st["color"], out["color"]))
    if want_depth:
        print("  depth bit mismatches:", bits_equal(st["depth"], out["depth"]))
    nc = r.field("n_contrib").float()
    print(f"  stats: n_contrib mean {nc.mean():.1f} max {nc.max():.0f}; R/tile {out['R'] / (((W+15)//16)*((H+15)//16)):.0f}")
    return scene, cam


def compare_backward(P, W, H, C, use_features, refname, seed=0):
    scene = make_scene(P, seed, sh=not use_features, channels=C if use_features else 0)
    cam = orbit_cameras(4, W, H)[1]
    sc, cm = dev_scene(scene, dev, requires_grad=True), dev_cam(cam, dev)
    bg = torch.zeros(C, device=dev)
    o = run_ours("chn", sc, cm, bg, use_features=use_features)
    dL = torch.as_tensor(np.random.default_rng(5).standard_normal((C, H, W)).astype(np.float32), device=dev)
    (o["color"] * dL).sum().backward()
    r = refmod.RefRasterizer(refname)
    sd = {k: (v.detach() if v is not None else None) for k, v in sc.items()}
    r.forward(bg=bg, means3D=sd["means3D"], opacities=sd["opacities"], viewmatrix=cm["viewmatrix"],
              projmatrix=cm["projmatrix"], campos=cm["campos"], tanfovx=cm["tanfovx"], tanfovy=cm["tanfovy"], W=W, H=H,
              shs=None if use_features else sd["shs"], colors_precomp=sd["features"] if use_features else None,
              scales=sd["scales"], rotations=sd["rotations"], num_channels=C)
    g = r.backward(dL)
    print(f"--- bwd P={P} {W}x{H} C={C} feat={use_features} ref={refname}")
    pairs = [("dL_dmeans2D", o["means2D"].grad), ("dL_dopacity", sc["opacities"].grad.view(-1)),
             ("dL_dmeans3D", sc["means3D"].grad), ("dL_dscales", sc["scales"].grad), ("dL_drotations", sc["rotations"].grad)]
    pairs.append(("dL_dcolors", sc["features"].grad) if use_features else ("dL_dsh", sc["shs"].grad))
    for name, got in pairs:
        print(f"  {name}: rel err {rel_err(got, g[name]):.3e}  frac>1e-4 {frac_bad(got, g[name]):.3e}")


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def timing(P, W, H, C, use_features, want_depth, backward, refname=None, nrep=10):
    scene = make_scene(P, 0, sh=not use_features, channels=C if use_features else 0)
    cam = orbit_cameras(4, W, H)[1]
    sc, cm = dev_scene(scene, dev, requires_grad=backward), dev_cam(cam, dev)
    bg = torch.zeros(C, device=dev)
    dL = torch.randn((C, H, W), device=dev)

    def ours():
        o = run_ours("rgbd" if want_depth else "chn", sc, cm, bg, use_features=use_features)
        if backward:
            o["color"].backward(dL)
            for v in sc.values():
                if v is not None:
                    v.grad = None
    t = timeit(ours, nrep)
    print(f"--- time P={P} {W}x{H} C={C} depth={want_depth} bwd={backward}: ours {t:.3f} ms", flush=True)
    if refname:
        r = refmod.RefRasterizer(refname)
        sd = {k: (v.detach() if v is not None else None) for k, v in sc.items()}

        def reff():
            r.forward(bg=bg, means3D=sd["means3D"], opacities=sd["opacities"], viewmatrix=cm["viewmatrix"],
                      projmatrix=cm["projmatrix"], campos=cm["campos"], tanfovx=cm["tanfovx"], tanfovy=cm["tanfovy"],
                      W=W, H=H, shs=None if use_features else sd["shs"],
                      colors_precomp=sd["features"] if use_features else None, scales=sd["scales"],
                      rotations=sd["rotations"], num_channels=C)
            if backward:
                r.backward(dL)
        tr = timeit(reff, max(2, nrep // 3), 1)
        print(f"    reference CUDA ({refname}) {tr:.3f} ms  -> speed-up {tr / t:.2f}x", flush=True)


if __name__ == "__main__":
    what = sys.argv[1:] or ["fwd", "bwd", "time"]
    if "fwd" in what:
        compare_forward(10000, 256, 256, 3, False, True)
        compare_forward(200000, 640, 480, 3, False, True, view=2)
        compare_forward(1000000, 1920, 1080, 3, False, True, view=1)
        compare_forward(100000, 640, 480, 32, True, False)
        compare_forward(100000, 640, 480, 100, True, False, view=3)
    if "bwd" in what:
        compare_backward(10000, 256, 256, 3, False, "chn")
        compare_backward(50000, 320, 240, 3, True, "chn")
        compare_backward(50000, 320, 240, 100, True, "chn_c100")
        compare_backward(100000, 640, 480, 256, True, "chn_c256")
    if "time" in what:
        timing(1000000, 1920, 1080, 3, False, True, False, "rgbd")
        timing(1000000, 1920, 1080, 256, True, False, False, "chn", nrep=5)
        timing(1000000, 1920, 1080, 256, True, False, True, None, nrep=5)
