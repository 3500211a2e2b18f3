"""The product's per-Gaussian gradient algebra (semantic-gaussians_b200/csrc/geom_grad.cuh: matrix-calculus form of
backward.cu:141-391) compiled for the HOST by g++ (tests/host/geom_grad_host.cpp) and compared with the oracle's
restatement on seeded scenes.  The GPU kernel calls the very same functions; its parity against the compiled
reference is in tests/test_parity_gpu.py."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle as orc

from semantic_gaussians_b200.scene_synth import make_scene, orbit_cameras

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("geomgrad") / "libgeomgrad_host.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++",
                           "-I", os.path.join(ROOT, "semantic-gaussians_b200", "csrc"),
                           os.path.join(HERE, "host", "geom_grad_host.cpp"), "-o", out])
    return C.CDLL(out)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _run(fn, scene, cam, pre, W, H, deg, g2d, gconic, gcol, use_sh, use_factors, scale_modifier=1.0):
    P = scene.P
    f32 = lambda a: np.ascontiguousarray(a, np.float32)  # noqa: E731
    shs = f32(scene.shs) if use_sh else None
    M = shs.shape[1] if use_sh else 0
    out = dict(mean=np.zeros((P, 3), np.float32), cov=np.zeros((P, 6), np.float32),
               sh=np.zeros((P, max(M, 1), 3), np.float32), scale=np.zeros((P, 3), np.float32),
               rot=np.zeros((P, 4), np.float32))
    cd = orc.cam_dict(cam)
    view, proj, cpos = f32(cd["viewmatrix"]).reshape(-1), f32(cd["projmatrix"]).reshape(-1), f32(cd["campos"]).reshape(-1)
    fx = np.float32(W) / (np.float32(2.0) * np.float32(cd["tanfovx"]))
    fy = np.float32(H) / (np.float32(2.0) * np.float32(cd["tanfovy"]))
    scales = f32(scene.scales) if use_factors else None
    rots = f32(scene.rotations) if use_factors else None
    xyz = f32(scene.xyz)
    fn(C.c_int(P), C.c_int(deg), C.c_int(M), _p(xyz), _p(pre["radii"]), _p(shs), _p(pre["clamped"]), _p(scales), _p(rots),
       C.c_float(scale_modifier), _p(pre["cov3D"]), _p(view), _p(proj), C.c_float(fx), C.c_float(fy),
       C.c_float(cd["tanfovx"]), C.c_float(cd["tanfovy"]), _p(cpos), _p(g2d), _p(gconic), _p(out["mean"]), _p(gcol),
       _p(out["cov"]), _p(out["sh"]), _p(out["scale"]), _p(out["rot"]))
    return out


def _close(a, b, tol=2e-5):
    scale = max(float(np.abs(b).max()), 1e-20)
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max()) / scale <= tol


@pytest.mark.parametrize("deg,use_sh,use_factors,mod,seed", [(3, True, True, 1.0, 1), (2, True, True, 0.7, 2),
                                                              (1, True, False, 1.0, 3), (0, True, True, 1.3, 4),
                                                              (0, False, True, 1.0, 5)])
def test_matches_oracle(host_lib, deg, use_sh, use_factors, mod, seed):
    W, H = 160, 112
    scene = make_scene(4000, seed=seed, sh=True, scale_mean=0.08)
    cam = orbit_cameras(4, W, H)[seed % 4]
    f = orc.forward(orc.scene_dict(scene), orc.cam_dict(cam), W, H, np.zeros(3, np.float32), sh_degree=deg,
                    scale_modifier=mod)
    pre = f["pre"]
    rng = np.random.default_rng(seed)
    P = scene.P
    g2d = rng.standard_normal((P, 3)).astype(np.float32)
    gconic = rng.standard_normal((P, 4)).astype(np.float32)
    gcol = rng.standard_normal((P, 3)).astype(np.float32)
    host_lib.host_geom_backward.restype = None
    orc.lib().orc_geom_backward.restype = None
    a = _run(host_lib.host_geom_backward, scene, cam, pre, W, H, deg, g2d, gconic, gcol, use_sh, use_factors, mod)
    b = _run(orc.lib().orc_geom_backward, scene, cam, pre, W, H, deg, g2d, gconic, gcol, use_sh, use_factors, mod)
    vis = pre["radii"] > 0
    assert vis.sum() > 500
    for k in ("mean", "cov", "sh", "scale", "rot"):
        assert np.all(a[k][~vis] == 0), k                 # culled Gaussians keep the zeros
        assert np.all(np.isfinite(a[k])), k
        # per-Gaussian relative agreement too (not only against the global maximum)
        assert _close(a[k], b[k]), (k, float(np.abs(a[k] - b[k]).max()), float(np.abs(b[k]).max()))
        num = np.abs(a[k].reshape(P, -1) - b[k].reshape(P, -1)).max(1)
        den = np.abs(b[k].reshape(P, -1)).max(1) + 1e-12
        assert np.quantile((num / den)[vis], 0.999) < 1e-3, k
    if not use_sh:
        assert not a["sh"].any()
    if not use_factors:
        assert not a["scale"].any() and not a["rot"].any()


def test_frustum_clamp_masks_direct_terms(host_lib):
    """A Gaussian far outside the frustum sideways (|tx/tz| > 1.3 tan): the clamp branch (backward.cu:170-176, 252-253)."""
    W, H = 160, 112
    scene = make_scene(3000, seed=9, sh=True, scale_mean=0.08)
    cam = orbit_cameras(4, W, H)[0]
    cd = orc.cam_dict(cam)
    view = np.asarray(cd["viewmatrix"], np.float32).reshape(4, 4)      # row-vector convention: t = p @ view[:3,:3] + view[3,:3]
    # push a tenth of the points sideways in VIEW space beyond the clamp but keep them in front of the camera
    t = scene.xyz.astype(np.float64) @ view[:3, :3].astype(np.float64) + view[3, :3]
    sel = np.arange(scene.P) % 10 == 0
    t[sel, 0] = np.sign(t[sel, 0] + 1e-9) * 1.45 * cd["tanfovx"] * np.abs(t[sel, 2])
    scene.scales[sel] *= 8.0                                           # big enough to still reach the image
    scene.xyz[:] = ((t - view[3, :3]) @ np.linalg.inv(view[:3, :3].astype(np.float64))).astype(np.float32)
    f = orc.forward(orc.scene_dict(scene), cd, W, H, np.zeros(3, np.float32), sh_degree=3)
    pre = f["pre"]
    rng = np.random.default_rng(0)
    P = scene.P
    g2d = rng.standard_normal((P, 3)).astype(np.float32)
    gconic = rng.standard_normal((P, 4)).astype(np.float32)
    gcol = rng.standard_normal((P, 3)).astype(np.float32)
    host_lib.host_geom_backward.restype = None
    orc.lib().orc_geom_backward.restype = None
    a = _run(host_lib.host_geom_backward, scene, cam, pre, W, H, 3, g2d, gconic, gcol, True, True)
    b = _run(orc.lib().orc_geom_backward, scene, cam, pre, W, H, 3, g2d, gconic, gcol, True, True)
    hit = sel & (pre["radii"] > 0)
    assert hit.sum() > 5, "no clamped Gaussian survived the cull: the case is not exercised"
    for k in ("mean", "cov", "scale", "rot"):
        assert _close(a[k][hit], b[k][hit], 5e-5), k
        assert _close(a[k], b[k]), k
