"""CPU: the brute-force distCUDA2 oracle against an independent exact 3-NN (scipy cKDTree, float64)."""
import os
import sys

import numpy as np
from scipy.spatial import cKDTree

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.knn_oracle import mean_dist2_3nn


def test_oracle_matches_kdtree():
    rng = np.random.default_rng(0)
    for n in (4, 50, 700):
        p = rng.uniform(-2, 2, (n, 3)).astype(np.float32)
        d, _ = cKDTree(p.astype(np.float64)).query(p.astype(np.float64), k=4)
        assert np.allclose(mean_dist2_3nn(p), (d[:, 1:] ** 2).mean(axis=1), rtol=1e-5)


def test_oracle_duplicates_and_tiny_clouds():
    p = np.array([[0, 0, 0], [0, 0, 0], [1, 0, 0], [0, 2, 0], [0, 0, 3]], np.float32)
    o = mean_dist2_3nn(p)
    assert o[0] == np.float32((0 + 1 + 4) / 3) and o[1] == o[0]       # the duplicate counts with distance 0
    assert np.isinf(mean_dist2_3nn(np.zeros((1, 3), np.float32))).all()
    assert (mean_dist2_3nn(p[2:]) > 1e37).all()                        # 2 neighbours + one FLT_MAX term
