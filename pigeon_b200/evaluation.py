"""The inference hot loop body: geocell prediction, (multi-GPU) gather, prototype refinement.

Mirrors the per-batch body of reference training/train_eval_loop.py:77-103 (`outputs = model(**data)` then
`refiner(outputs.embedding, initial_preds=outputs.preds_LLH, candidate_cells=top5.indices,
candidate_probs=top5.values)`), with the reference's multi-GPU exchange (preprocessing/embed.py:36-37) placed
before the retrieval step so that every rank refines the whole gathered batch against its bank replica.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import dist as pdist


@torch.no_grad()
def predict_batch(model, refiner, batch: Dict[str, torch.Tensor], gather: bool = True):
    """batch: keyword arguments of `SuperGuessr.forward` (dataset column names, as in `model(**data)`).
    Returns (preds_LLH float32 [B_total, 2], preds_geocell int64 [B_total], outputs) — B_total is the gathered
    batch when running distributed with `gather`, else the local one."""
    outputs = model(**batch)
    if isinstance(outputs, tuple) and not hasattr(outputs, "preds_LLH"):     # serving tuple
        pred_LLH, topk, embedding = outputs[0], outputs[1], outputs[-1]
    else:
        pred_LLH, topk, embedding = outputs.preds_LLH, outputs.top5_geocells, outputs.embedding
    if refiner is None:
        return pred_LLH, (topk.indices[:, 0] if topk is not None else None), outputs
    # proto_refiner.py:139-140 averages the views first; the head kernel already produced that mean
    # (`model.last_pooled`), so the gather moves (B, D) instead of (B, 4, D).  Without it the refiner kernel pools.
    emb = getattr(model, "last_pooled", None)
    if emb is None or emb.shape[0] != embedding.shape[0]:
        emb = embedding
    pack = dict(emb=emb.float().contiguous(), idx=topk.indices, val=topk.values, init=pred_LLH)
    if gather and pdist.is_distributed():
        pack = pdist.all_gather_rows(pack)
    _, ll, cell = refiner(pack["emb"], initial_preds=pack["init"], candidate_cells=pack["idx"],
                          candidate_probs=pack["val"])
    return ll, cell, outputs
