"""Open-vocabulary semantic head on the GPU: what every ``render_chn`` caller of
the reference does with the rendered feature image (eval_segmentation.py:153-157, 253-257, 394-398;
view_viser.py:312-315) and with the per-Gaussian features (eval_segmentation.py:132, view_viser.py:185).

``semantic_head``      one pass over the (C,H,W) image: L2-normalise per pixel, similarities with the K
                       text embeddings, arg-max label map.
``feature_logits``     per-Gaussian similarities ``einsum("cq,dq->dc", text, features)``.
``render_semantic_labels``  label map WITHOUT the (C,H,W) feature image: alpha blending is linear in the
                       blended attribute, so rendering the K per-Gaussian similarities gives the same
                       un-normalised per-pixel similarities; the positive per-pixel normalisation does
                       not change the arg-max."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib


def _check(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (the semantic head has no CPU path)")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _stream_ctx(t: torch.Tensor):
    # callers hold `with torch.cuda.device(t.device)`: the kernels launch on a stream of t's device
    stream = torch.cuda.current_stream(t.device).cuda_stream
    return stream, _lib.ctx_for(t.device.index, stream)


def semantic_head(rendering: torch.Tensor, text_features: torch.Tensor, first_class: int = 1,
                  return_sim: bool = True, return_label: bool = True
                  ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """rendering (C,H,W), text_features (K,C)  ->  (sim (K,H,W) float32, label (H,W) int64) with

        r     = rendering / (rendering.norm(dim=0, keepdim=True) + 1e-8)
        sim   = torch.einsum("cq,qhw->chw", text_features, r)
        label = sim[first_class:].argmax(dim=0)          # the reference then adds 1 (class 0 = "other")
    """
    r = _check(rendering, "rendering")
    t = _check(text_features, "text_features").to(r.device)
    if r.ndim != 3 or t.ndim != 2 or t.shape[1] != r.shape[0]:
        raise ValueError("rendering must be (C,H,W) and text_features (K,C)")
    C_, H, W = r.shape
    K = t.shape[0]
    if not (0 <= first_class < K):
        raise ValueError("first_class out of range")
    sim = torch.empty((K, H, W), dtype=torch.float32, device=r.device) if return_sim else None
    label = torch.empty((H, W), dtype=torch.int64, device=r.device) if return_label else None
    with torch.cuda.device(r.device):
        stream, ctx = _stream_ctx(r)
        _lib.check(_lib.load().sgb_semantic_head(ctx, C_, K, H * W, r.data_ptr(), t.data_ptr(), first_class,
                                                sim.data_ptr() if sim is not None else None,
                                                label.data_ptr() if label is not None else None, stream),
                   "sgb_semantic_head")
    return sim, label


def feature_logits(features: torch.Tensor, text_features: torch.Tensor, pad_to: int = 1) -> torch.Tensor:
    """features (P,C), text_features (K,C) -> (P, Kpad) similarities, ``einsum("cq,dq->dc", text, features)``
    in columns [0,K), zeros in the padding columns (Kpad = K rounded up to a multiple of ``pad_to``)."""
    f = _check(features, "features")
    t = _check(text_features, "text_features").to(f.device)
    if f.ndim != 2 or t.ndim != 2 or t.shape[1] != f.shape[1]:
        raise ValueError("features must be (P,C) and text_features (K,C)")
    P, C_ = f.shape
    K = t.shape[0]
    Kpad = ((K + pad_to - 1) // pad_to) * pad_to
    out = torch.empty((P, Kpad), dtype=torch.float32, device=f.device)
    with torch.cuda.device(f.device):
        stream, _ = _stream_ctx(f)
        _lib.check(_lib.load().sgb_feature_logits(P, C_, K, Kpad, f.data_ptr(), t.data_ptr(), out.data_ptr(), stream),
                   "sgb_feature_logits")
    return out


def label_argmax(planes: torch.Tensor, num_classes: Optional[int] = None, first_class: int = 1) -> torch.Tensor:
    """planes (K',H,W) -> (H,W) int64 = planes[first_class:num_classes].argmax(dim=0)
    (``rendering[1:].argmax(dim=0)``, eval_segmentation.py:144)."""
    p = _check(planes, "planes")
    if p.ndim != 3:
        raise ValueError("planes must be (K,H,W)")
    K = p.shape[0] if num_classes is None else int(num_classes)
    if not (0 < K <= p.shape[0]) or not (0 <= first_class < K):
        raise ValueError("bad num_classes / first_class")
    H, W = p.shape[1:]
    label = torch.empty((H, W), dtype=torch.int64, device=p.device)
    with torch.cuda.device(p.device):
        stream, _ = _stream_ctx(p)
        _lib.check(_lib.load().sgb_label_argmax(K, first_class, H * W, p.data_ptr(), label.data_ptr(), stream),
                   "sgb_label_argmax")
    return label


def distill_loss_and_grad(rendering: torch.Tensor, class_emb: torch.Tensor, labels: torch.Tensor
                          ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Open-vocabulary distillation loss of a rendered (C,H,W) feature image against per-pixel class embeddings,
    and its gradient, in one pass over the image:

        loss = -(rendering * class_emb[labels].permute(2, 0, 1)).mean()          # labels (H,W) int32/int64
        grad = d loss / d rendering                                               # (C,H,W)

    Pixels whose label is outside [0, K) (ScanNet-style -1 / 255) are ignored: zero gradient, no loss term, and the
    mean runs over the valid pixels only.

    Returns (loss: 0-d float64 CUDA tensor, grad: (C,H,W) float32).  Use as ``rendering.backward(grad)``."""
    r = _check(rendering.detach(), "rendering")
    e = _check(class_emb, "class_emb").to(r.device)
    if r.ndim != 3 or e.ndim != 2 or e.shape[1] != r.shape[0]:
        raise ValueError("rendering must be (C,H,W) and class_emb (K,C)")
    if not labels.is_cuda or labels.dtype not in (torch.int32, torch.int64) or labels.numel() != r.shape[1] * r.shape[2]:
        raise ValueError("labels must be a CUDA int32/int64 tensor with H*W entries")
    lab = labels.contiguous()
    grad = torch.empty_like(r)
    loss2 = torch.zeros(2, dtype=torch.float64, device=r.device)      # [loss, number of non-ignored pixels]
    with torch.cuda.device(r.device):
        stream, _ = _stream_ctx(r)
        _lib.check(_lib.load().sgb_distill_loss(r.shape[0], e.shape[0], r.shape[1] * r.shape[2], r.data_ptr(),
                                               e.data_ptr(), lab.data_ptr(), int(lab.dtype == torch.int64),
                                               grad.data_ptr(), loss2.data_ptr(), stream), "sgb_distill_loss")
    return loss2[0], grad


def render_semantic_labels(viewpoint_camera, pc, pipe, bg_color: torch.Tensor, text_features: torch.Tensor,
                           features: Optional[torch.Tensor] = None, first_class: int = 1, scaling_modifier=1.0,
                           override_shape=None, foreground=None, world_rotate=None,
                           logits: Optional[torch.Tensor] = None) -> dict:
    """Label map of one view straight from per-Gaussian features, never writing the (C,H,W) image.

    Equivalent (up to fp32 re-association of the per-pixel sums) to the reference sequence
    ``render_chn(..., num_channels=C, override_color=features)`` -> normalise -> einsum -> ``sim[1:].argmax``:
    sum_c text[k][c] * (sum_j w_j f_j[c] + T bg[c]) = sum_j w_j (text[k].f_j) + T (text[k].bg).
    ``logits``: a precomputed ``feature_logits(features, text_features, pad_to=4)`` — it depends only on the
    scene and the label set, so an evaluation loop computes it once and passes it for every view.
    Returns {"label": (H,W) int64, "logits": (K,H,W) un-normalised similarities, "radii", "visibility_filter"}."""
    from .renderer import render_chn
    t = _check(text_features, "text_features")
    K = t.shape[0]
    if logits is not None:
        g = _check(logits, "logits")
        if g.ndim != 2 or g.shape[1] < K or g.shape[1] % 4:
            raise ValueError("logits must be feature_logits(features, text_features, pad_to=4)")
    else:
        if features is None:
            features = pc._features_semantic
        g = feature_logits(features, t, pad_to=4)                  # (P, Kpad): 16-byte rows for the blend kernels
    bg = _check(bg_color, "bg_color").reshape(-1)
    bgk = torch.zeros(g.shape[1], dtype=torch.float32, device=g.device)
    bgk[:K] = t.to(g.device) @ bg.to(g.device)
    with torch.no_grad():
        out = render_chn(viewpoint_camera, pc, pipe, bgk, scaling_modifier=scaling_modifier, num_channels=g.shape[1],
                         override_color=g, override_shape=override_shape, foreground=foreground,
                         world_rotate=world_rotate)
    planes = out["render"]
    return {"label": label_argmax(planes, K, first_class), "logits": planes[:K], "radii": out["radii"],
            "visibility_filter": out["visibility_filter"]}
