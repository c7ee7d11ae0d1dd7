"""Head-only fine-tune step restated for the CPU (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows reference training/train_eval_loop.py:187 (torch.optim.AdamW(model.parameters(), lr)), :215-221
(loss.backward() per micro-batch, optimizer.step()/zero_grad() every `grad_acc_steps`) and
models/super_guessr.py:437-474 (view mean -> cell_layer -> CrossEntropyLoss on index / soft / haversine-smoothed
targets).  The optimizer itself is third-party (torch.optim.AdamW, single-tensor path); its published update is
restated in `adamw_step` in the operation order of torch/optim/adamw.py.  Pinned by tests/golden/train_head.npz,
which oracle/make_golden.py produced by running the unmodified reference module under torch autograd.
"""
from __future__ import annotations

import numpy as np

from .geo import haversine_matrix, smooth_labels


def _targets(C, labels, labels_clf, centroids, smooth):
    import torch
    if smooth:
        t = smooth_labels(haversine_matrix(torch.as_tensor(labels), torch.as_tensor(centroids).t())).numpy()
    else:
        labels_clf = np.asarray(labels_clf)
        if labels_clf.ndim == 1:                       # _to_one_hot, super_guessr.py:298-313
            t = np.zeros((labels_clf.shape[0], C), dtype=np.float64)
            t[np.arange(labels_clf.shape[0]), labels_clf] = 1.0
        else:
            t = labels_clf.astype(np.float64)
    return np.asarray(t, dtype=np.float64)


def head_loss_and_grads(emb, w, b, labels, labels_clf, centroids, smooth, panorama=True):
    """Returns loss (f64), dW [C, D], db [C], dpooled [B, D] of mean-reduced soft-target cross entropy (f64 math)."""
    emb = np.asarray(emb, dtype=np.float64)
    pooled = emb.mean(axis=1) if (panorama and emb.ndim == 3) else emb
    w64, b64 = np.asarray(w, np.float64), np.asarray(b, np.float64)
    logits = pooled.astype(np.float32).astype(np.float64) @ w64.T + b64
    m = logits.max(axis=1, keepdims=True)
    lse = m + np.log(np.exp(logits - m).sum(axis=1, keepdims=True))
    logp = logits - lse
    t = _targets(w64.shape[0], labels, labels_clf, centroids, smooth)
    B = logits.shape[0]
    loss = -(t * logp).sum() / B
    g = (np.exp(logp) * t.sum(axis=1, keepdims=True) - t) / B
    return loss, g.T @ pooled, g.sum(axis=0), g @ w64, g


def adamw_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.01):
    """torch.optim.AdamW, single-tensor path, fp32 state; returns the new (p, m, v)."""
    f = np.float32
    p, g, m, v = (np.asarray(x, dtype=f) for x in (p, g, m, v))
    p = p * f(1.0 - lr * weight_decay)
    m = m + f(1.0 - beta1) * (g - m)
    v = v * f(beta2) + (f(1.0 - beta2) * g) * g
    bc1, bc2 = 1.0 - beta1 ** step, 1.0 - beta2 ** step
    denom = np.sqrt(v) / f(np.sqrt(bc2)) + f(eps)
    p = p + f(-(lr / bc1)) * (m / denom)
    return p.astype(f), m.astype(f), v.astype(f)


def tower_gradients(sd, pixel_values, d_emb, *, patch: int, heads: int, layers: int, eps: float = 1e-5):
    """Gradients of the vision tower for a given upstream gradient `d_emb` [n, hidden] with respect to the token-mean
    embedding (models/super_guessr.py:395-398): what `loss.backward()` (training/train_eval_loop.py:216) propagates into
    HF CLIPVisionTransformer, obtained by differentiating the oracle's own fp32 restatement of its forward (oracle/vit.py)
    with torch autograd on the CPU.  Returns (embedding [n, hidden], {stripped parameter name: gradient}).
    Pinned against tests/golden/train_tower_small.npz (gradients of the unmodified reference model)."""
    import torch
    from . import vit as ovit
    names = [k for k, v in ovit._strip(sd).items() if v.is_floating_point() and "post_layernorm" not in k and "position_ids" not in k]
    leaf = {k: v.detach().clone().float().requires_grad_(True) for k, v in ovit._strip(sd).items() if k in names}
    with torch.enable_grad():
        h = ovit.vit_last_hidden_state.__wrapped__(leaf, pixel_values.float(), patch=patch, heads=heads, layers=layers, eps=eps)
        emb = h.mean(dim=1)
        grads = torch.autograd.grad(emb, [leaf[k] for k in names], grad_outputs=torch.as_tensor(d_emb, dtype=torch.float32))
    return emb.detach(), {k: g for k, g in zip(names, grads)}
