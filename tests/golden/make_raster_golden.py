"""Generates tests/golden/raster_golden_k1.npz on the GPU box by running the UNMODIFIED compiled
reference rasterizer (oracle/_ref/libref_{rgbd,chn,chn_c100}.so, built by oracle/build.py) on the
seeded K1-size scene (10k Gaussians, 256x256).  Inputs are regenerated from the seed by the tests;
the file stores the reference's outputs and opaque-state fields.

    gpurun -- 'python tests/golden/make_raster_golden.py'   # writes gpurun_out/raster_golden_k1.npz
then copy gpurun_out/raster_golden_k1.npz to tests/golden/."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

K1 = dict(P=10000, W=256, H=256, seed=0)
KF = dict(P=3000, W=64, H=48, C=100, seed=11)        # feature raster (chn / chn_c100)


def golden_inputs(kind):
    from semantic_gaussians_b200.scene_synth import make_scene, orbit_cameras
    if kind == "k1":
        scene = make_scene(K1["P"], K1["seed"], sh=True)
        cam = orbit_cameras(1, K1["W"], K1["H"])[0]
        rng = np.random.default_rng(5)
        dL = rng.standard_normal((3, K1["H"], K1["W"])).astype(np.float32)
        bg = np.array([0.1, 0.2, 0.3], np.float32)
    else:
        scene = make_scene(KF["P"], KF["seed"], channels=KF["C"], scale_mean=0.04)
        cam = orbit_cameras(3, KF["W"], KF["H"])[1]
        rng = np.random.default_rng(6)
        dL = rng.standard_normal((KF["C"], KF["H"], KF["W"])).astype(np.float32)
        bg = (0.01 * np.arange(KF["C"])).astype(np.float32)
    return scene, cam, dL, bg


def main():
    from oracle import ref as refmod
    from util import dev_cam, dev_scene
    dev = torch.device("cuda:0")
    out = {}
    # ---- K1: RGB (SH degree 3) + median depth through the rgbd library, fwd + bwd
    scene, cam, dL, bg = golden_inputs("k1")
    sc, cm = dev_scene(scene, dev), dev_cam(cam, dev)
    r = refmod.RefRasterizer("rgbd")
    f = r.forward(bg=torch.as_tensor(bg, device=dev), means3D=sc["means3D"], opacities=sc["opacities"],
                  viewmatrix=cm["viewmatrix"], projmatrix=cm["projmatrix"], campos=cm["campos"],
                  tanfovx=cm["tanfovx"], tanfovy=cm["tanfovy"], W=cm["W"], H=cm["H"], shs=sc["shs"],
                  scales=sc["scales"], rotations=sc["rotations"], num_channels=3)
    out["k1_R"] = np.int64(f["R"])
    out["k1_color"] = f["color"].cpu().numpy()
    out["k1_depth"] = f["depth"].cpu().numpy()
    out["k1_radii"] = f["radii"].cpu().numpy()
    for name in ("depths", "means2D", "conic_opacity", "cov3D", "rgb", "clamped", "tiles_touched", "point_list",
                 "ranges", "n_contrib", "accum_alpha"):
        out["k1_" + name] = r.field(name).cpu().numpy()
    g = r.backward(torch.as_tensor(dL, device=dev))
    for k, v in g.items():
        out["k1_" + k] = v.cpu().numpy()
    # ---- KF: 100-channel feature raster; forward by the stock chn library, backward by the
    #      NUM_CHANNELS=100 rebuild (the stock backward is 3-channel only)
    scene, cam, dL, bg = golden_inputs("kf")
    sc, cm = dev_scene(scene, dev), dev_cam(cam, dev)
    kw = dict(bg=torch.as_tensor(bg, device=dev), means3D=sc["means3D"], opacities=sc["opacities"],
              viewmatrix=cm["viewmatrix"], projmatrix=cm["projmatrix"], campos=cm["campos"], tanfovx=cm["tanfovx"],
              tanfovy=cm["tanfovy"], W=cm["W"], H=cm["H"], colors_precomp=sc["features"], scales=sc["scales"],
              rotations=sc["rotations"], num_channels=KF["C"])
    r0 = refmod.RefRasterizer("chn")
    f0 = r0.forward(**kw)
    out["kf_R"] = np.int64(f0["R"])
    out["kf_color"] = f0["color"].cpu().numpy()
    out["kf_radii"] = f0["radii"].cpu().numpy()
    out["kf_n_contrib"] = r0.field("n_contrib").cpu().numpy()
    out["kf_accum_alpha"] = r0.field("accum_alpha").cpu().numpy()
    out["kf_point_list"] = r0.field("point_list").cpu().numpy()
    r1 = refmod.RefRasterizer("chn_c100")
    r1.forward(**kw)
    g = r1.backward(torch.as_tensor(dL, device=dev))
    for k, v in g.items():
        if k != "dL_dsh":
            out["kf_" + k] = v.cpu().numpy()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "raster_golden_k1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
