"""CPU experiment for DESIGN.md §8 item 3 (LayerNorm folded into the next GEMM): does feeding the tensor cores the RAW fp16
residual row with gamma folded into the weights and correcting in the epilogue (rstd * (acc - mu * s_n) + b'_n) keep the
1e-3 budget against the fp32 reference, compared with today's fp16(LayerNorm(x)) operand?  Emulates operand rounding only
(fp16 GEMM operands, fp32 accumulation, fp32 residual stream), on the synthetic ViT-L weights, optionally with injected
outlier channels (real CLIP residual streams have a few channels two orders of magnitude above the rest)."""
import sys
import os

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vit as ovit  # noqa: E402
from pigeon_b200 import synthetic  # noqa: E402
from pigeon_b200.vit_engine import VitDims  # noqa: E402

h16 = lambda t: t.to(torch.float16).to(torch.float32)


def forward(sd, px, dims, mode, outlier=0.0):
    """mode: 'fp32' reference | 'ln16' today's path | 'fold' raw-row operand + folded gamma + epilogue correction."""
    sd = ovit._strip(sd)
    q = (lambda t: t) if mode == "fp32" else h16
    n, H, heads = px.shape[0], dims.hidden, dims.heads
    w = sd["embeddings.patch_embedding.weight"].float()
    pe = F.conv2d(q(px), q(w), stride=dims.patch_size).flatten(2).transpose(1, 2)
    x = torch.cat([sd["embeddings.class_embedding"].float().expand(n, 1, H), pe], 1) + sd["embeddings.position_embedding.weight"].float()
    if outlier:
        x[..., 7] += outlier
        x[..., 300] -= 0.6 * outlier
    x = F.layer_norm(x, (H,), sd["pre_layrnorm.weight"].float(), sd["pre_layrnorm.bias"].float(), dims.ln_eps)
    if outlier:                                     # keep the outliers in the residual stream itself
        x[..., 7] += outlier
        x[..., 300] -= 0.6 * outlier

    def ln_linear(x, g, b, W, bias):
        """LayerNorm(x) @ W^T + bias under the three operand schemes."""
        if mode == "fp32":
            return F.linear(F.layer_norm(x, (H,), g, b, dims.ln_eps), W, bias)
        if mode == "ln16":
            return F.linear(h16(F.layer_norm(x, (H,), g, b, dims.ln_eps)), h16(W), bias)
        mu = x.mean(-1, keepdim=True)
        rstd = torch.rsqrt(x.var(-1, unbiased=False, keepdim=True) + dims.ln_eps)
        Wf = h16(W * g)                             # gamma folded, then rounded like any weight
        acc = F.linear(h16(x), Wf)                  # raw fp16 row on the tensor cores, fp32 accumulation
        return rstd * (acc - mu * Wf.sum(-1)) + (bias + F.linear(b, W))

    for i in range(dims.layers):
        p = f"encoder.layers.{i}."
        g1, b1 = sd[p + "layer_norm1.weight"].float(), sd[p + "layer_norm1.bias"].float()
        Wqkv = torch.cat([sd[p + f"self_attn.{k}_proj.weight"].float() for k in "qkv"])
        bqkv = torch.cat([sd[p + f"self_attn.{k}_proj.bias"].float() for k in "qkv"])
        qkv = q(ln_linear(x, g1, b1, Wqkv, bqkv))
        s = x.shape[1]
        qq, kk, vv = (t.view(n, s, heads, 64).transpose(1, 2) for t in qkv.split(H, -1))
        a = torch.softmax(qq @ kk.transpose(-1, -2) * 0.125, -1)
        o = q((q(a) @ vv).transpose(1, 2).reshape(n, s, H))
        x = x + F.linear(o, q(sd[p + "self_attn.out_proj.weight"].float()), sd[p + "self_attn.out_proj.bias"].float())
        g2, b2 = sd[p + "layer_norm2.weight"].float(), sd[p + "layer_norm2.bias"].float()
        u = ln_linear(x, g2, b2, sd[p + "mlp.fc1.weight"].float(), sd[p + "mlp.fc1.bias"].float())
        hh = q(u * torch.sigmoid(1.702 * u))
        x = x + F.linear(hh, q(sd[p + "mlp.fc2.weight"].float()), sd[p + "mlp.fc2.bias"].float())
    return x.mean(1)


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    dims = VitDims()
    sd = synthetic.random_vit_state_dict(dims, seed=0)
    px = torch.randn(2, 3, 336, 336, generator=torch.Generator().manual_seed(1))
    rel = lambda a, b: ((a - b).norm() / b.norm()).item()
    with torch.no_grad():
        for outlier in (0.0, 30.0, 150.0):
            ref = forward(sd, px, dims, "fp32", outlier)
            a = forward(sd, px, dims, "ln16", outlier)
            b = forward(sd, px, dims, "fold", outlier)
            print(f"outlier {outlier:6.1f}: embedding rel-L2 vs fp32  fp16(LN(x)) operand {rel(a, ref):.3e}   "
                  f"raw-row + folded gamma {rel(b, ref):.3e}", flush=True)
