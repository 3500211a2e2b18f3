"""The CPU oracle (oracle/raster_oracle.c): self-consistency, independent finite-difference check of
its backward, and the pin against outputs of the compiled reference (tests/golden/raster_golden_k1.npz,
generated on a B200 by tests/golden/make_raster_golden.py)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from oracle import oracle as orc  # noqa: E402

from semantic_gaussians_b200.scene_synth import make_scene, orbit_cameras  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "raster_golden_k1.npz")


@pytest.fixture(scope="module")
def small():
    scene = make_scene(3000, seed=5, sh=True, scale_mean=0.05)
    cam = orbit_cameras(3, 96, 80)[1]
    f = orc.forward(orc.scene_dict(scene), orc.cam_dict(cam), 96, 80, np.array([0.2, 0.1, 0.0], np.float32),
                    want_depth=True)
    return scene, cam, f


def test_binning_invariants(small):
    scene, cam, f = small
    pre, b = f["pre"], f["bin"]
    assert b["R"] == int(pre["tiles_touched"].sum()) > 0
    keys = b["keys"]
    assert np.all(keys[1:] >= keys[:-1])                                   # sorted by (tile, depth bits)
    same = keys[1:] == keys[:-1]
    assert np.all(b["point_list"][1:][same] > b["point_list"][:-1][same])  # ties keep ascending id (stable)
    tiles = (keys >> np.uint64(32)).astype(np.int64)
    rng = b["ranges"].astype(np.int64)
    for t in np.unique(tiles):
        idx = np.nonzero(tiles == t)[0]
        assert rng[t, 0] == idx[0] and rng[t, 1] == idx[-1] + 1
    assert np.all(rng[np.setdiff1d(np.arange(rng.shape[0]), np.unique(tiles))] == 0)
    vis = pre["radii"] > 0
    assert np.all(pre["depths"][vis] > 0.2)                                # near cull (auxiliary.h:154)
    w = pre["rect"][:, 2] - pre["rect"][:, 0]
    h = pre["rect"][:, 3] - pre["rect"][:, 1]
    assert np.array_equal((w * h)[vis].astype(np.uint32), pre["tiles_touched"][vis])


def test_forward_invariants(small):
    scene, cam, f = small
    T = f["final_T"]
    assert np.all((T >= 1e-4 * 0.0) & (T <= 1.0))
    rng = f["bin"]["ranges"].astype(np.int64)
    assert f["n_contrib"].max() <= (rng[:, 1] - rng[:, 0]).max()
    d = f["depth"][0]
    assert np.all((d == 15.0) | ((d > 0.2) & (d < 10)))                    # default 15.0 or a real depth
    # colour = sum + T * bg: with all-zero features the image is exactly T * bg
    z = orc.render_forward(f["pre"], f["bin"], np.zeros((scene.P, 3), np.float32), np.array([0.5, 1, 2], np.float32), 96, 80)
    assert np.array_equal(z["color"][1].ravel(), z["final_T"])


def test_backward_is_the_derivative_of_forward():
    """Feature path (exactly linear in the features) and opacity (smooth) by central differences."""
    C, W, H = 5, 48, 40
    scene = make_scene(300, seed=8, channels=C, scale_mean=0.08)
    cam = orbit_cameras(2, W, H)[0]
    sd, cd = orc.scene_dict(scene), orc.cam_dict(cam)
    bg = np.linspace(0.1, 0.5, C).astype(np.float32)
    dL = np.random.default_rng(0).standard_normal((C, H, W)).astype(np.float32)
    f = orc.forward(sd, cd, W, H, bg, features=scene.features)
    g = orc.backward(f, sd, cd, W, H, bg, dL, features=scene.features)
    loss = lambda fw: float((fw["color"].astype(np.float64) * dL).sum())
    rng = np.random.default_rng(1)
    vis = np.nonzero(np.abs(g["dL_dcolors"]).sum(1) > 0)[0]
    for gid in rng.choice(vis, 6, replace=False):
        ch = int(rng.integers(0, C))
        feats = scene.features.copy()
        feats[gid, ch] += 0.5
        lp = loss(orc.forward(sd, cd, W, H, bg, features=feats))
        feats[gid, ch] -= 1.0
        lm = loss(orc.forward(sd, cd, W, H, bg, features=feats))
        assert abs((lp - lm) - g["dL_dcolors"][gid, ch]) <= 2e-3 * max(1.0, abs(g["dL_dcolors"][gid, ch]))
    big = vis[np.argsort(-np.abs(g["dL_dopacity"][vis]))[:5]]
    for gid in big:
        eps = 2e-3
        op = scene.opacity.copy()
        op[gid] += eps
        lp = loss(orc.forward(dict(sd, opacity=op), cd, W, H, bg, features=scene.features))
        op[gid] -= 2 * eps
        lm = loss(orc.forward(dict(sd, opacity=op), cd, W, H, bg, features=scene.features))
        fd = (lp - lm) / (2 * eps)
        assert abs(fd - g["dL_dopacity"][gid]) <= 0.05 * abs(g["dL_dopacity"][gid]) + 1e-3


def test_mean3d_gradient_by_finite_differences():
    C, W, H = 3, 40, 32
    scene = make_scene(150, seed=2, channels=C, scale_mean=0.1)
    cam = orbit_cameras(2, W, H)[1]
    sd, cd = orc.scene_dict(scene), orc.cam_dict(cam)
    bg = np.zeros(C, np.float32)
    dL = np.random.default_rng(4).standard_normal((C, H, W)).astype(np.float32)
    f = orc.forward(sd, cd, W, H, bg, features=scene.features)
    g = orc.backward(f, sd, cd, W, H, bg, dL, features=scene.features)
    loss = lambda fw: float((fw["color"].astype(np.float64) * dL).sum())
    order = np.argsort(-np.abs(g["dL_dmeans3D"]).sum(1))[:4]
    ok = 0
    for gid in order:
        for ax in range(3):
            eps = 2e-3
            xyz = scene.xyz.copy()
            xyz[gid, ax] += eps
            lp = loss(orc.forward(dict(sd, xyz=xyz), cd, W, H, bg, features=scene.features))
            xyz[gid, ax] -= 2 * eps
            lm = loss(orc.forward(dict(sd, xyz=xyz), cd, W, H, bg, features=scene.features))
            fd = (lp - lm) / (2 * eps)
            an = g["dL_dmeans3D"][gid, ax]
            ok += abs(fd - an) <= 0.1 * abs(an) + 0.02 * np.abs(g["dL_dmeans3D"][gid]).max()
    assert ok >= 10        # alpha thresholds make the forward piecewise smooth: allow 2 of 12 to miss


@pytest.mark.skipif(not os.path.exists(GOLD), reason="golden fixture not generated yet (needs one GPU run)")
def test_oracle_pinned_to_reference_golden_k1():
    """Float state within a few ulp of the compiled reference; integer stage equal except where the
    oracle's own pre-rounding value is within 1e-4 of a decision boundary; pixels / gradients within
    tolerance away from threshold-fragile pixels."""
    from make_raster_golden import K1, golden_inputs
    gold = np.load(GOLD)
    scene, cam, dL, bg = golden_inputs("k1")
    W, H = K1["W"], K1["H"]
    f = orc.forward(orc.scene_dict(scene), orc.cam_dict(cam), W, H, bg, want_depth=True)
    pre = f["pre"]
    vis = gold["k1_radii"] > 0
    mism = pre["radii"] != gold["k1_radii"]
    rr = pre["raw_radius"][mism]
    assert mism.sum() <= 5 and np.all(np.abs(rr - np.round(rr)) < 1e-4 * np.maximum(rr, 1))
    both = vis & (pre["radii"] > 0)
    np.testing.assert_allclose(pre["means2D"][both], gold["k1_means2D"][both], rtol=2e-6, atol=2e-4)
    np.testing.assert_allclose(pre["conic_opacity"][both], gold["k1_conic_opacity"][both], rtol=2e-4, atol=1e-7)
    ulp = np.abs(pre["depths"][both].view(np.int32).astype(np.int64) - gold["k1_depths"][both].view(np.int32))
    assert ulp.max() <= 4                       # view-space depth: a 4-term dot product, contraction differences only
    np.testing.assert_allclose(pre["rgb"][both], gold["k1_rgb"][both], rtol=1e-5, atol=1e-6)
    assert (pre["tiles_touched"] != gold["k1_tiles_touched"].view(np.uint32)).sum() <= 5
    if f["bin"]["R"] == int(gold["k1_R"]):
        assert (f["bin"]["point_list"] != gold["k1_point_list"].view(np.uint32)).mean() < 1e-3
    ok = ~f["fragile"]
    assert ok.mean() > 0.995
    nc = gold["k1_n_contrib"].view(np.uint32).reshape(H, W)
    assert (f["n_contrib"].reshape(H, W)[ok] != nc[ok]).mean() < 2e-3
    err = np.abs(f["color"] - gold["k1_color"])[:, ok]
    assert err.max() <= 1e-4 * np.abs(gold["k1_color"]).max() + 1e-6
    dep_ok = ok & (np.abs(f["depth"][0] - gold["k1_depth"][0]) < 1e-5)
    assert dep_ok.sum() >= 0.995 * ok.sum()
    g = orc.backward(f, orc.scene_dict(scene), orc.cam_dict(cam), W, H, bg, dL)
    for name in ("dL_dmeans3D", "dL_dopacity", "dL_dsh", "dL_dscales", "dL_drotations"):
        want = gold["k1_" + name].astype(np.float64).reshape(g[name].shape)
        tol = 1e-3 * np.abs(want) + 1e-3 * np.abs(want).max()
        assert (np.abs(g[name] - want) > tol).mean() < 5e-3, name
