"""TEST INFRASTRUCTURE — numpy restatement of the reference's fusion projection
(dataset/fusion_utils.py:17-78 PointCloudToImageMapper) and accumulate loop (fusion.py:127-148).
Written independently of the reference file so it can run where /root/reference is absent;
tests/test_fusion_cpu.py pins it against the imported reference class in this container."""
from __future__ import annotations

import numpy as np


def rescale_intrinsics(intrinsics, image_dim):
    """fusion_utils.py:22-28."""
    K = np.array(intrinsics, dtype=np.float64).copy()
    sx = image_dim[0] / (K[0, 2] * 2)
    sy = image_dim[1] / (K[1, 2] * 2)
    K[0, 0] *= sx
    K[1, 1] *= sy
    K[0, 2] = image_dim[0] / 2
    K[1, 2] = image_dim[1] / 2
    return K


def compute_mapping(world_to_camera, coords, image_dim, K, vis_thres, cut_bound, depth=None):
    """fusion_utils.py:30-78 → (N,3) int64 [v, u, mask]."""
    N = coords.shape[0]
    mapping = np.zeros((3, N), dtype=np.int64)
    coords_new = np.concatenate([coords, np.ones([N, 1])], axis=1).T          # float64 (:41)
    p = np.matmul(np.asarray(world_to_camera).T, coords_new)                   # (:43)
    with np.errstate(divide="ignore", invalid="ignore"):
        p[0] = (p[0] * K[0][0]) / p[2] + K[0][2]
        p[1] = (p[1] * K[1][1]) / p[2] + K[1][2]
        pi = np.round(p).astype(np.int64)                                       # round half to even (:48)
    inside = ((pi[0] >= cut_bound) * (pi[1] >= cut_bound) * (pi[0] < image_dim[0] - cut_bound)
              * (pi[1] < image_dim[1] - cut_bound))
    if isinstance(depth, str):                                                  # "surface" (:57-61)
        zb = np.ones((image_dim[1], image_dim[0])) * 999999
        ok = (p[2] > 0.2) & inside
        np.minimum.at(zb, (pi[1][ok], pi[0][ok]), p[2][ok])
        depth = zb
    if depth is not None:                                                       # (:63-69)
        depth = np.asarray(depth)
        dcur = depth[pi[1][inside], pi[0][inside]]
        occ = np.abs(depth[pi[1][inside], pi[0][inside]] - p[2][inside]) <= vis_thres * dcur
        inside[inside == True] = occ                                            # noqa: E712
    else:
        inside = (p[2] > 0) * inside                                            # (:70-72)
    mapping[0][inside] = pi[1][inside]
    mapping[1][inside] = pi[0][inside]
    mapping[2][inside] = 1
    return mapping.T


def accumulate(features, mapping, feat_sum, count):
    """fusion.py:136-144 for one view: features (C,h,w) fp16/fp32; feat_sum (P,C) f32; count (P,) f32."""
    mask = mapping[:, 2] != 0
    if mask.sum() == 0:
        return
    fm = features[:, mapping[:, 0], mapping[:, 1]].T          # gathers pixel (0,0) for masked-out points
    count[mask] += 1
    feat_sum[mask] += fm[mask].astype(np.float32)


def normalize(feat_sum, count):
    """fusion.py:146-147."""
    count[count == 0] = 1e-5
    feat_sum /= count[:, None]
