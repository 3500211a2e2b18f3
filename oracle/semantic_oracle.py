"""TEST INFRASTRUCTURE ONLY — numpy restatement of the reference's open-vocabulary semantic head, used as the
checker for semantic-gaussians_b200/csrc/semantic.cu (never imported by the product path).

Follows eval_segmentation.py:153-157 (per-pixel normalise, einsum with the text features, arg-max over
classes 1..K-1) and eval_segmentation.py:132 / view_viser.py:185 (per-Gaussian similarities).  Pinned
against the reference's own torch expressions in tests/test_semantic_cpu.py (the lines are plain torch
ops that run on the CPU)."""
import numpy as np


def semantic_head(rendering: np.ndarray, text: np.ndarray, first_class: int = 1):
    """rendering (C,H,W) f32, text (K,C) f32 -> sim (K,H,W) f32, label (H,W) int64."""
    r = rendering.astype(np.float32)
    norm = np.sqrt((r.astype(np.float64) ** 2).sum(axis=0)).astype(np.float32)          # :155 norm(dim=0)
    r = r / (norm[None] + np.float32(1e-8))
    sim = np.einsum("cq,qhw->chw", text.astype(np.float64), r.astype(np.float64)).astype(np.float32)  # :156
    label = sim[first_class:].argmax(axis=0).astype(np.int64)                            # :157
    return sim, label


def feature_logits(features: np.ndarray, text: np.ndarray):
    """features (P,C), text (K,C) -> (P,K): einsum("cq,dq->dc")  (eval_segmentation.py:132)."""
    return np.einsum("cq,dq->dc", text.astype(np.float64), features.astype(np.float64)).astype(np.float32)


def label_margin(sim: np.ndarray, first_class: int = 1):
    """gap between the best and the second-best class per pixel (labels are only comparable where it is
    larger than the fp32 noise of the two implementations)."""
    s = np.sort(sim[first_class:], axis=0)
    return s[-1] - s[-2] if s.shape[0] > 1 else np.full(sim.shape[1:], np.inf, np.float32)


def distill_loss_and_grad(rendering: np.ndarray, class_emb: np.ndarray, labels: np.ndarray):
    """loss = -mean(rendering * class_emb[labels].transpose(2,0,1)); grad = d loss / d rendering  (float64 sums)."""
    C = rendering.shape[0]
    tgt = class_emb[labels.astype(np.int64)].transpose(2, 0, 1).astype(np.float64)          # (C,H,W)
    n = float(C * labels.size)
    loss = -(rendering.astype(np.float64) * tgt).sum() / n
    grad = (-(tgt / n)).astype(np.float32)
    return loss, grad
