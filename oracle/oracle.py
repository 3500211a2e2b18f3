"""TEST INFRASTRUCTURE — numpy/ctypes front end of the CPU restatement (oracle/raster_oracle.c).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this module."""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            import importlib.util
            spec = importlib.util.spec_from_file_location("oracle_build", os.path.join(HERE, "build.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mod.build_oracle()
        _lib = C.CDLL(path)
        _lib.orc_count_instances.restype = C.c_longlong
        _lib.orc_bin.restype = C.c_longlong
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> int:
    """OpenMP threads of the C restatement (overrides an inherited OMP_NUM_THREADS)."""
    lib().orc_set_num_threads(C.c_int(int(n)))
    return num_threads()


def preprocess(means3D, scales, rotations, opacities, viewmatrix, projmatrix, campos, W, H, tanfovx, tanfovy,
               shs=None, colors_precomp=None, cov3D_precomp=None, scale_modifier=1.0, sh_degree=3):
    """forward.cu:155-256.  Returns a dict of per-Gaussian arrays."""
    means3D = _f32(means3D)
    P = means3D.shape[0]
    scales, rotations, opacities = _f32(scales), _f32(rotations), _f32(np.asarray(opacities).reshape(-1))
    shs, colors_precomp, cov3D_precomp = _f32(shs), _f32(colors_precomp), _f32(cov3D_precomp)
    view, proj, cam = _f32(viewmatrix).reshape(-1), _f32(projmatrix).reshape(-1), _f32(campos).reshape(-1)
    M = shs.shape[1] if shs is not None else 0
    out = dict(radii=np.zeros(P, np.int32), means2D=np.zeros((P, 2), np.float32), depths=np.zeros(P, np.float32),
               cov3D=np.zeros((P, 6), np.float32), conic_opacity=np.zeros((P, 4), np.float32),
               rgb=np.zeros((P, 3), np.float32), clamped=np.zeros((P, 3), np.uint8),
               tiles_touched=np.zeros(P, np.uint32), rect=np.zeros((P, 4), np.int32),
               raw_radius=np.zeros(P, np.float32))
    lib().orc_preprocess(C.c_int(P), C.c_int(sh_degree), C.c_int(M), _p(means3D), _p(scales),
                         C.c_float(scale_modifier), _p(rotations), _p(opacities), _p(shs), _p(cov3D_precomp),
                         _p(colors_precomp), _p(view), _p(proj), _p(cam), C.c_int(W), C.c_int(H),
                         C.c_float(tanfovx), C.c_float(tanfovy), _p(out["radii"]), _p(out["means2D"]),
                         _p(out["depths"]), _p(out["cov3D"]), _p(out["conic_opacity"]), _p(out["rgb"]),
                         _p(out["clamped"]), _p(out["tiles_touched"]), _p(out["rect"]), _p(out["raw_radius"]))
    if cov3D_precomp is not None:
        out["cov3D"] = cov3D_precomp
    return out


def mark_visible(means3D, viewmatrix):
    means3D = _f32(means3D)
    P = means3D.shape[0]
    present = np.zeros(P, np.uint8)
    lib().orc_mark_visible(C.c_int(P), _p(means3D), _p(_f32(viewmatrix).reshape(-1)), _p(present))
    return present.astype(bool)


def bin_instances(pre, W, H, tile_rows=None):
    """rasterizer_impl.cu:70-138, 277-321 → point_list (R,), keys (R,) u64, ranges (tiles,2).
    tile_rows=(r0, r1) restricts emission to those tile rows (bench.py's bounded CPU sample)."""
    P = pre["radii"].shape[0]
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    r0, r1 = tile_rows if tile_rows is not None else (0, 0)
    if tile_rows is None:
        cap = int(lib().orc_count_instances(C.c_int(P), _p(pre["tiles_touched"])))
    else:  # upper bound: every visible Gaussian's rect width times the band height
        rect = pre["rect"]
        y0 = np.maximum(rect[:, 1], r0)
        y1 = np.minimum(rect[:, 3], r1)
        cap = int((np.maximum(y1 - y0, 0).astype(np.int64) * (rect[:, 2] - rect[:, 0])).sum())
    point_list = np.zeros(max(cap, 1), np.uint32)
    keys = np.zeros(max(cap, 1), np.uint64)
    ranges = np.zeros((tiles, 2), np.uint32)
    R = int(lib().orc_bin(C.c_int(P), C.c_int(W), C.c_int(H), _p(pre["radii"]), _p(pre["means2D"]), _p(pre["depths"]),
                          _p(pre["tiles_touched"]), _p(point_list), _p(keys), _p(ranges), C.c_int(r0), C.c_int(r1)))
    return dict(R=R, point_list=point_list[:R], keys=keys[:R], ranges=ranges)


def render_forward(pre, binning, features, bg, W, H, want_depth=False, rows=(0, 0)):
    features = _f32(features)
    Cn = features.shape[1]
    bg = _f32(bg).reshape(-1)
    out_color = np.zeros((Cn, H, W), np.float32)
    final_T = np.zeros(H * W, np.float32)
    n_contrib = np.zeros(H * W, np.uint32)
    out_depth = np.zeros((1, H, W), np.float32) if want_depth else None
    fragile = np.zeros(H * W, np.uint8)
    pl = binning["point_list"] if binning["R"] > 0 else np.zeros(1, np.uint32)
    lib().orc_render_forward(C.c_int(W), C.c_int(H), C.c_int(Cn), _p(binning["ranges"]), _p(pl), _p(pre["means2D"]),
                             _p(features), _p(pre["conic_opacity"]), _p(pre["depths"]), _p(bg), _p(out_color),
                             _p(final_T), _p(n_contrib), _p(out_depth), _p(fragile), C.c_int(rows[0]), C.c_int(rows[1]))
    return dict(color=out_color, final_T=final_T, n_contrib=n_contrib, depth=out_depth,
                fragile=fragile.reshape(H, W).astype(bool))


def forward(scene_arrays: dict, cam: dict, W: int, H: int, bg, features=None, want_depth=False, sh_degree=3,
            scale_modifier=1.0):
    """Whole forward of one view.  scene_arrays: xyz, scales, rotations, opacity [, shs] [, cov3D_precomp];
    cam: viewmatrix, projmatrix, campos, tanfovx, tanfovy; features (P,C) = colors_precomp or None (SH)."""
    pre = preprocess(scene_arrays["xyz"], scene_arrays.get("scales"), scene_arrays.get("rotations"),
                     scene_arrays["opacity"], cam["viewmatrix"], cam["projmatrix"], cam["campos"], W, H,
                     cam["tanfovx"], cam["tanfovy"], shs=scene_arrays.get("shs") if features is None else None,
                     colors_precomp=features, cov3D_precomp=scene_arrays.get("cov3D_precomp"),
                     scale_modifier=scale_modifier, sh_degree=sh_degree)
    b = bin_instances(pre, W, H)
    colors = features if features is not None else pre["rgb"]
    r = render_forward(pre, b, colors, bg, W, H, want_depth)
    return dict(pre=pre, bin=b, colors=_f32(colors), **r)


def backward(fwd: dict, scene_arrays: dict, cam: dict, W: int, H: int, bg, dL_dpix, features=None, sh_degree=3,
             scale_modifier=1.0, rows=(0, 0)):
    """backward.cu: blend backward + per-Gaussian backward.  Returns the reference's gradient set
    (float64 for the four accumulated-by-atomics quantities, float32 for the per-Gaussian chain,
    which consumes them rounded to float32 like the reference's buffers)."""
    pre, b = fwd["pre"], fwd["bin"]
    P = pre["radii"].shape[0]
    colors = fwd["colors"]
    Cn = colors.shape[1]
    dL_dpix = _f32(dL_dpix)
    bg = _f32(bg).reshape(-1)
    g_mean2D = np.zeros((P, 3), np.float64)
    g_conic = np.zeros((P, 4), np.float64)
    g_opac = np.zeros(P, np.float64)
    g_colors = np.zeros((P, Cn), np.float64)
    pl = b["point_list"] if b["R"] > 0 else np.zeros(1, np.uint32)
    lib().orc_render_backward(C.c_int(W), C.c_int(H), C.c_int(Cn), _p(b["ranges"]), _p(pl), _p(bg),
                              _p(pre["means2D"]), _p(pre["conic_opacity"]), _p(colors), _p(fwd["final_T"]),
                              _p(fwd["n_contrib"]), _p(dL_dpix), _p(g_mean2D), _p(g_conic), _p(g_opac), _p(g_colors),
                              C.c_int(rows[0]), C.c_int(rows[1]))
    shs = _f32(scene_arrays.get("shs")) if features is None else None
    M = shs.shape[1] if shs is not None else 0
    scales, rots = _f32(scene_arrays.get("scales")), _f32(scene_arrays.get("rotations"))
    means3D = _f32(scene_arrays["xyz"])
    d_mean3D = np.zeros((P, 3), np.float32)
    d_cov = np.zeros((P, 6), np.float32)
    d_sh = np.zeros((P, max(M, 1), 3), np.float32)
    d_scale = np.zeros((P, 3), np.float32)
    d_rot = np.zeros((P, 4), np.float32)
    m2d32, conic32 = g_mean2D.astype(np.float32), g_conic.astype(np.float32)
    col32 = g_colors.astype(np.float32) if Cn == 3 else np.zeros((P, 3), np.float32)
    view, proj, cam_pos = (_f32(cam["viewmatrix"]).reshape(-1), _f32(cam["projmatrix"]).reshape(-1),
                           _f32(cam["campos"]).reshape(-1))
    focal_x = np.float32(W) / (np.float32(2.0) * np.float32(cam["tanfovx"]))
    focal_y = np.float32(H) / (np.float32(2.0) * np.float32(cam["tanfovy"]))
    lib().orc_geom_backward(C.c_int(P), C.c_int(sh_degree), C.c_int(M), _p(means3D), _p(pre["radii"]), _p(shs),
                            _p(pre["clamped"]), _p(scales), _p(rots), C.c_float(scale_modifier), _p(pre["cov3D"]),
                            _p(view), _p(proj), C.c_float(focal_x), C.c_float(focal_y), C.c_float(cam["tanfovx"]),
                            C.c_float(cam["tanfovy"]), _p(cam_pos), _p(m2d32), _p(conic32), _p(d_mean3D), _p(col32),
                            _p(d_cov), _p(d_sh), _p(d_scale), _p(d_rot))
    return dict(dL_dmeans2D=g_mean2D, dL_dconic=g_conic, dL_dopacity=g_opac, dL_dcolors=g_colors,
                dL_dmeans3D=d_mean3D, dL_dcov3D=d_cov, dL_dsh=d_sh[:, :M], dL_dscales=d_scale, dL_drotations=d_rot)


def cam_dict(cam) -> dict:
    """SynthCamera → the dict the oracle takes."""
    return dict(viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform, campos=cam.camera_center,
                tanfovx=math.tan(cam.FoVx * 0.5), tanfovy=math.tan(cam.FoVy * 0.5))


def scene_dict(scene) -> dict:
    d = dict(xyz=scene.xyz, scales=scene.scales, rotations=scene.rotations, opacity=scene.opacity)
    if scene.shs is not None:
        d["shs"] = scene.shs
    return d
