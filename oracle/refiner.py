"""ProtoRefiner.forward restated on the CPU from reference models/proto_refiner.py (line numbers below).

The bank is the same CSR packing the CUDA path consumes (`pg_refiner_bank`), built by `pack_bank` from the
reference's own data model: a per-geocell list of prototype rows (cluster lng/lat, count, member indices, mean
embedding — proto_refiner.py:288-313,359-378) plus the training embeddings/labels they index.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .geo import haversine


def pack_bank(cells: Sequence[Optional[List[dict]]], data_emb: np.ndarray, data_lnglat: np.ndarray) -> Dict[str, np.ndarray]:
    """cells[c] is None (reference: protos[c] is None, :296-303) or a list of prototype dicts with keys
    lng, lat, count, indices (list[int]) and embedding (np.float32 [D])."""
    D = data_emb.shape[1]
    cell_off = [0]
    emb, lnglat, count, moff, midx = [], [], [], [0], []
    for c in cells:
        if c:
            for p in c:
                emb.append(np.asarray(p["embedding"], np.float32))
                lnglat.append([np.float32(p["lng"]), np.float32(p["lat"])])
                count.append(int(p["count"]))
                midx.extend(int(i) for i in p["indices"])
                moff.append(len(midx))
        cell_off.append(len(emb))
    return dict(
        cell_off=np.asarray(cell_off, np.int64),
        proto_emb=(np.stack(emb) if emb else np.zeros((0, D), np.float32)).astype(np.float32),
        proto_lnglat=np.asarray(lnglat, np.float32).reshape(-1, 2),
        proto_count=np.asarray(count, np.int32),
        member_off=np.asarray(moff, np.int64),
        member_idx=np.asarray(midx, np.int64),
        data_emb=np.ascontiguousarray(data_emb, np.float32),
        data_lnglat=np.ascontiguousarray(data_lnglat, np.float32),
    )


def _euclid(matrix: torch.Tensor, vector: torch.Tensor) -> torch.Tensor:
    """:332-344 — torch.cdist(matrix, v[None]).flatten()"""
    return torch.cdist(matrix, vector.unsqueeze(0)).flatten()


@torch.no_grad()
def refiner_forward(bank: Dict[str, np.ndarray], embedding: torch.Tensor, initial_preds: torch.Tensor,
                    candidate_cells: torch.Tensor, candidate_probs: Optional[torch.Tensor], topk: int,
                    temperature: float, max_refinement: float, exact_distance: bool = False):
    """Returns (preds_LLH f32 [B,2], preds_geocell i64 [B], info dict).  :121-231.

    exact_distance=True evaluates the distances as direct float64 differences instead of torch.cdist's fp32
    (matmul-based beyond 25 rows) arithmetic — used to measure how far apart near-tied candidates are."""
    assert topk <= candidate_cells.size(1)                                        # :135-137
    if embedding.dim() == 3:
        embedding = embedding.mean(dim=1)                                         # :139-140
    if candidate_probs is None:                                                   # :143-145
        candidate_probs = torch.zeros_like(candidate_cells)
        candidate_probs[:, 0] = 1
    T = torch.tensor(temperature, dtype=torch.float32)                            # :89
    cell_off = bank["cell_off"]
    P = torch.from_numpy(bank["proto_emb"])
    E = torch.from_numpy(bank["data_emb"])
    preds, cells_out, choice = [], [], []
    B = embedding.shape[0]
    best_logit = np.zeros((B, topk), np.float32)
    best_lnglat = np.zeros((B, topk, 2), np.float32)
    best_proto = np.full((B, topk), -1, np.int32)
    gap = np.full((B, topk), np.inf, np.float64)  # distance gap between the best and second-best prototype / member
    margin = np.full((B,), np.inf, np.float64)    # relative gap between the two largest final probabilities
    for i in range(B):                                                            # :154
        emb, cands, c_probs = embedding[i], candidate_cells[i], candidate_probs[i]
        top_preds, top_d = [], []
        for j in range(topk):                                                     # :160
            cell = int(cands[j])
            lo, hi = (int(cell_off[cell]), int(cell_off[cell + 1])) if 0 <= cell < len(cell_off) - 1 else (0, 0)
            if hi <= lo:                                                          # :168-174
                top_d.append(-100000.0)
                top_preds.append([0.0, 0.0])
                best_logit[i, j] = -100000.0
                continue
            if exact_distance:
                d = (P[lo:hi].double() - emb.double()).norm(dim=1)
            else:
                d = _euclid(P[lo:hi], emb)
            logits = -d                                                           # :177
            top_d.append(torch.max(logits).item())                                # :180
            pid = int(torch.argmax(logits))                                       # :181
            if hi - lo > 1:
                srt = torch.sort(d.double()).values
                gap[i, j] = float(srt[1] - srt[0])
            p = lo + pid
            best_proto[i, j] = p
            if int(bank["proto_count"][p]) == 1:                                  # :243-244
                lng, lat = float(bank["proto_lnglat"][p, 0]), float(bank["proto_lnglat"][p, 1])
            else:                                                                 # :246-255 (argMAX: farthest member)
                idx = bank["member_idx"][int(bank["member_off"][p]): int(bank["member_off"][p + 1])]
                m = E[torch.from_numpy(idx)]
                dd = (m.double() - emb.double()).norm(dim=1) if exact_distance else _euclid(m, emb)
                mi = int(torch.argmax(dd))
                if len(idx) > 1:
                    srt = torch.sort(dd.double(), descending=True).values
                    gap[i, j] = min(gap[i, j], float(srt[0] - srt[1]))
                lng, lat = float(bank["data_lnglat"][idx[mi], 0]), float(bank["data_lnglat"][idx[mi], 1])
            top_preds.append([lng, lat])
            best_logit[i, j] = top_d[-1]
            best_lnglat[i, j] = (lng, lat)
        top_distances = torch.tensor(top_d)                                       # :187 (float32)
        ex = torch.exp(top_distances / T)                                         # :355-357, no max-subtraction
        probs = ex / torch.sum(ex, axis=0)
        cp = c_probs[:topk]
        initial_guess = torch.argmax(cp).item()                                   # :191
        final_probs = cp * probs                                                  # :192
        refined_guess = torch.argmax(final_probs).item()                          # :193
        refined = torch.tensor(top_preds[refined_guess]).unsqueeze(0)             # :198-199 (float32)
        initial = initial_preds[i].unsqueeze(0)                                   # :200
        distance = haversine(initial, refined)[0]                                 # :201
        if distance > max_refinement:                                             # :202-203
            final_probs = cp
        final_id = torch.argmax(final_probs).item()                               # :219
        if topk > 1:
            for fp in (cp * probs, cp):
                t2 = torch.topk(fp.double(), 2).values
                margin[i] = min(margin[i], float((t2[0] - t2[1]) / t2[0].abs().clamp_min(1e-300)))
            margin[i] = min(margin[i], abs(float(distance) - float(max_refinement)) / max(float(max_refinement), 1e-9))
        choice.append(final_id)
        preds.append(top_preds[final_id])                                         # :221
        cells_out.append(int(cands[final_id]))                                    # :222
        del initial_guess
    preds_LLH = torch.tensor(preds, dtype=torch.float32).reshape(B, 2)            # :229
    preds_geocell = torch.tensor(cells_out, dtype=torch.int64)                    # :230
    info = dict(best_logit=best_logit, best_lnglat=best_lnglat, best_proto=best_proto,
                choice=np.asarray(choice, np.int32), proto_gap=gap, margin=margin)
    return preds_LLH, preds_geocell, info
