"""fp32 restatement of HF `CLIPVisionTransformer.forward(...).last_hidden_state`.

Third-party arithmetic the reference reaches at models/clip_embedder.py:63 and models/super_guessr.py:395
(`transformers` `models/clip/modeling_clip.py`, pinned 4.23.1 in env.yml:60):
  CLIPVisionEmbeddings: Conv2d(3, hidden, k=patch, s=patch, bias=False) -> flatten -> cat(class_embedding)
                        -> + position_embedding
  CLIPVisionTransformer: pre_layrnorm -> 24 x CLIPEncoderLayer (pre-LN) ; post_layernorm only on the
                        pooled CLS token, which PIGEON never uses.
  CLIPEncoderLayer: x = x + out_proj(attn(layer_norm1(x))); x = x + fc2(quick_gelu(fc1(layer_norm2(x))))
  CLIPAttention (4.23.1): q = q_proj(x) * head_dim**-0.5 ; softmax(q k^T) v, fp32, no mask, no dropout in eval
  quick_gelu(x) = x * sigmoid(1.702 x)
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F


def _strip(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    out = {}
    for k, v in sd.items():
        for pre in ("base_model.", "clip_model.", "vision_model."):
            while k.startswith(pre):
                k = k[len(pre):]
        out[k] = v
    return out


@torch.no_grad()
def vit_last_hidden_state(sd: Dict[str, torch.Tensor], pixel_values: torch.Tensor, *, patch: int, heads: int,
                          layers: int, eps: float = 1e-5, return_layers: bool = False):
    """pixel_values [N,3,H,W] fp32 -> last_hidden_state [N, tokens, hidden] fp32 (CPU or any device)."""
    sd = _strip(sd)
    x = pixel_values.to(torch.float32)
    n = x.shape[0]
    w = sd["embeddings.patch_embedding.weight"].float()
    hidden = w.shape[0]
    pe = F.conv2d(x, w, bias=None, stride=patch)                  # [N, hidden, gp, gp]
    pe = pe.flatten(2).transpose(1, 2)                            # [N, gp*gp, hidden]
    cls = sd["embeddings.class_embedding"].float().expand(n, 1, hidden)
    h = torch.cat([cls, pe], dim=1) + sd["embeddings.position_embedding.weight"].float().unsqueeze(0)
    h = F.layer_norm(h, (hidden,), sd["pre_layrnorm.weight"].float(), sd["pre_layrnorm.bias"].float(), eps)
    hd = hidden // heads
    scale = hd ** -0.5
    per_layer = []
    for i in range(layers):
        p = f"encoder.layers.{i}."
        r = h
        y = F.layer_norm(h, (hidden,), sd[p + "layer_norm1.weight"].float(), sd[p + "layer_norm1.bias"].float(), eps)
        q = F.linear(y, sd[p + "self_attn.q_proj.weight"].float(), sd[p + "self_attn.q_proj.bias"].float()) * scale
        k = F.linear(y, sd[p + "self_attn.k_proj.weight"].float(), sd[p + "self_attn.k_proj.bias"].float())
        v = F.linear(y, sd[p + "self_attn.v_proj.weight"].float(), sd[p + "self_attn.v_proj.bias"].float())
        s = h.shape[1]
        q = q.view(n, s, heads, hd).transpose(1, 2)
        k = k.view(n, s, heads, hd).transpose(1, 2)
        v = v.view(n, s, heads, hd).transpose(1, 2)
        a = torch.softmax(q @ k.transpose(-1, -2), dim=-1)
        o = (a @ v).transpose(1, 2).reshape(n, s, hidden)
        o = F.linear(o, sd[p + "self_attn.out_proj.weight"].float(), sd[p + "self_attn.out_proj.bias"].float())
        h = r + o
        r = h
        y = F.layer_norm(h, (hidden,), sd[p + "layer_norm2.weight"].float(), sd[p + "layer_norm2.bias"].float(), eps)
        y = F.linear(y, sd[p + "mlp.fc1.weight"].float(), sd[p + "mlp.fc1.bias"].float())
        y = y * torch.sigmoid(1.702 * y)
        y = F.linear(y, sd[p + "mlp.fc2.weight"].float(), sd[p + "mlp.fc2.bias"].float())
        h = r + y
        if return_layers:
            per_layer.append(h.clone())
    return (h, per_layer) if return_layers else h


@torch.no_grad()
def clip_embedding(sd, pixel_values, **kw) -> torch.Tensor:
    """reference models/clip_embedder.py:63-65 — mean over ALL tokens of last_hidden_state (pre post_layernorm)."""
    return vit_last_hidden_state(sd, pixel_values, **kw).mean(dim=1)
