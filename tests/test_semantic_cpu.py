"""CPU: the semantic-head oracle against the reference's own torch expressions (eval_segmentation.py:155-157,
:132 — plain torch ops, restated here verbatim as the pin since the script itself needs CLIP/OpenSeg)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import semantic_oracle as so


def _case(C, K, H, W, seed):
    rng = np.random.default_rng(seed)
    r = rng.standard_normal((C, H, W)).astype(np.float32) * rng.uniform(0.1, 3.0, (1, H, W)).astype(np.float32)
    t = rng.standard_normal((K, C)).astype(np.float32)
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    return r, t


def test_head_oracle_matches_reference_expressions():
    for C, K, H, W in ((16, 5, 7, 9), (64, 21, 12, 10), (33, 40, 5, 6)):
        r, t = _case(C, K, H, W, C + K)
        rendering, text_features = torch.from_numpy(r), torch.from_numpy(t)
        rendering = rendering / (rendering.norm(dim=0, keepdim=True) + 1e-8)
        sim = torch.einsum("cq,qhw->chw", text_features, rendering)
        label = sim[1:].argmax(dim=0)
        osim, olabel = so.semantic_head(r, t, 1)
        assert np.allclose(osim, sim.numpy(), rtol=1e-5, atol=2e-6)
        clear = so.label_margin(osim) > 1e-5
        assert clear.mean() > 0.9 and np.array_equal(olabel[clear], label.numpy()[clear])


def test_head_oracle_zero_pixel_and_first_class():
    r, t = _case(8, 4, 3, 3, 1)
    r[:, 1, 1] = 0.0
    sim, label = so.semantic_head(r, t, 1)
    assert np.all(sim[:, 1, 1] == 0.0) and label[1, 1] == 0        # all-zero pixel: sim 0, first class wins
    sim0, label0 = so.semantic_head(r, t, 0)
    assert np.array_equal(sim0, sim) and label0.max() <= 3


def test_logits_oracle_matches_reference_expression():
    rng = np.random.default_rng(5)
    f = rng.standard_normal((50, 24)).astype(np.float32)
    t = rng.standard_normal((7, 24)).astype(np.float32)
    sim = torch.einsum("cq,dq->dc", torch.from_numpy(t), torch.from_numpy(f))
    assert np.allclose(so.feature_logits(f, t), sim.numpy(), rtol=1e-5, atol=1e-5)
