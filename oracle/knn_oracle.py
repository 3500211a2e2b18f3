"""TEST INFRASTRUCTURE ONLY — numpy restatement of the reference's distCUDA2
(submodules/simple-knn/simple_knn.cu:136-183): mean of the three smallest squared distances to the other
points.  Brute force, float32 arithmetic in the order of simple_knn.cu:140-141 (x, then y, then z; numpy has no
fused multiply-add, so the GPU results may differ in the last bit).  Pinned against the compiled reference
(oracle/_ref/libref_knn.so) in tests/test_knn_gpu.py."""
import numpy as np


def mean_dist2_3nn(points: np.ndarray) -> np.ndarray:
    p = np.ascontiguousarray(points, dtype=np.float32)
    n = p.shape[0]
    out = np.empty(n, np.float32)
    fmax = np.float32(np.finfo(np.float32).max)
    for i in range(n):
        d = p - p[i]
        dist = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        dist = np.delete(dist, i)
        best = np.sort(dist)[:3]
        if best.shape[0] < 3:
            best = np.concatenate([best, np.full(3 - best.shape[0], fmax, np.float32)])
        with np.errstate(over="ignore"):
            out[i] = ((best[0] + best[1]) + best[2]) / np.float32(3.0)
    return out
