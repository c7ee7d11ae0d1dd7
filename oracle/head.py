"""Geocell head restated from reference models/super_guessr.py:386-483 (inference branch + CE loss)."""
from __future__ import annotations

from collections import namedtuple

import torch
import torch.nn.functional as F

from .geo import haversine_matrix, smooth_labels

HeadOut = namedtuple("HeadOut", "pooled logits probs pred_cell pred_LLH topk_val topk_idx loss")


@torch.no_grad()
def head_forward(embedding: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, centroids: torch.Tensor,
                 num_candidates: int, panorama: bool, labels: torch.Tensor | None = None,
                 labels_clf: torch.Tensor | None = None, should_smooth_labels: bool = False) -> HeadOut:
    """embedding: [B, 4, D] if panorama else [B, D] (fp32); centroids f64 [C, 2] (lng, lat).

    super_guessr.py:437 view mean; :447 cell_layer; :448 softmax; :454 argmax; :455 index_select on the f64
    centroid table; :459 topk over PROBABILITIES; :469-474 (optional) haversine-smoothed soft-target CE."""
    if panorama:
        out = embedding.mean(dim=1)                                   # :437
    elif embedding.dim() == 3 and embedding.size(1) == 4:
        out = embedding[:, 0]                                         # :440-441
    else:
        out = embedding
    logits = F.linear(out, weight, bias)                              # :447
    probs = torch.softmax(logits, dim=-1)                             # :448
    pred = torch.argmax(probs, dim=-1)                                # :454
    pred_llh = torch.index_select(centroids, 0, pred)                 # :455 (float64)
    topk = torch.topk(probs, num_candidates, dim=-1)                  # :459
    loss = None
    if labels_clf is not None:
        label_probs = labels_clf
        if should_smooth_labels:
            label_probs = smooth_labels(haversine_matrix(labels, centroids.t()))   # :469-471
        loss = F.cross_entropy(logits, label_probs)                   # :474
    return HeadOut(out, logits, probs, pred, pred_llh, topk.values, topk.indices, loss)
