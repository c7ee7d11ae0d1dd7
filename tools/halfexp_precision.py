"""CPU experiment for DESIGN.md §8 item 1: would a packed-half exponential (ex2.approx.f16x2: the exponent argument
rounded to fp16 BEFORE the exponential, two results per MUFU instruction) keep the 1e-3 embedding budget?  Emulates today's
operand rounding through all 24 layers and switches only the softmax numerator:  P = fp16(2^x) (today)  vs  P = 2^fp16(x)
with x = (S - m) c + lag, lag in [0, 8) being how far the lazily updated running maximum may trail the true one."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vit as ovit  # noqa: E402
from pigeon_b200 import synthetic  # noqa: E402
from pigeon_b200.vit_engine import VitDims  # noqa: E402

h16 = lambda t: t.to(torch.float16).to(torch.float32)
LOG2E = 1.4426950408889634


def forward(sd, px, dims, mode, lag=0.0):
    sd = ovit._strip(sd)
    q = (lambda t: t) if mode == "fp32" else h16
    n, H, heads = px.shape[0], dims.hidden, dims.heads
    pe = F.conv2d(q(px), q(sd["embeddings.patch_embedding.weight"].float()), stride=dims.patch_size).flatten(2).transpose(1, 2)
    x = torch.cat([sd["embeddings.class_embedding"].float().expand(n, 1, H), pe], 1) + sd["embeddings.position_embedding.weight"].float()
    x = F.layer_norm(x, (H,), sd["pre_layrnorm.weight"].float(), sd["pre_layrnorm.bias"].float(), dims.ln_eps)
    g = torch.Generator().manual_seed(5)
    for i in range(dims.layers):
        p = f"encoder.layers.{i}."
        y = q(F.layer_norm(x, (H,), sd[p + "layer_norm1.weight"].float(), sd[p + "layer_norm1.bias"].float(), dims.ln_eps))
        W = torch.cat([sd[p + f"self_attn.{k}_proj.weight"].float() for k in "qkv"])
        b = torch.cat([sd[p + f"self_attn.{k}_proj.bias"].float() for k in "qkv"])
        qkv = q(F.linear(y, q(W), b))
        s = x.shape[1]
        qq, kk, vv = (t.view(n, s, heads, 64).transpose(1, 2) for t in qkv.split(H, -1))
        sc = qq @ kk.transpose(-1, -2) * (0.125 * LOG2E)                        # log2-domain logits
        if mode == "fp32":
            a = torch.softmax(sc / LOG2E, -1)
            o = (a @ vv)
        else:
            xs = sc - sc.max(-1, keepdim=True).values
            if lag:
                xs = xs + torch.rand(xs.shape[:-1] + (1,), generator=g) * lag   # running maximum trails by up to `lag`
            pnum = torch.exp2(h16(xs)) if mode == "halfexp" else torch.exp2(xs)
            pnum = h16(pnum)                                                    # P as the fp16 tensor-core operand
            lsum = (pnum if mode == "halfexp" else torch.exp2(xs)).sum(-1, keepdim=True)
            o = (pnum @ vv) / lsum
        o = q(o.transpose(1, 2).reshape(n, s, H))
        x = x + F.linear(o, q(sd[p + "self_attn.out_proj.weight"].float()), sd[p + "self_attn.out_proj.bias"].float())
        y = q(F.layer_norm(x, (H,), sd[p + "layer_norm2.weight"].float(), sd[p + "layer_norm2.bias"].float(), dims.ln_eps))
        u = F.linear(y, q(sd[p + "mlp.fc1.weight"].float()), sd[p + "mlp.fc1.bias"].float())
        x = x + F.linear(q(u * torch.sigmoid(1.702 * u)), q(sd[p + "mlp.fc2.weight"].float()), sd[p + "mlp.fc2.bias"].float())
    return x.mean(1)


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    dims = VitDims()
    px = torch.randn(2, 3, 336, 336, generator=torch.Generator().manual_seed(1))
    rel = lambda a, b: ((a - b).norm() / b.norm()).item()
    with torch.no_grad():
        for std in (0.02, 0.06):          # 0.06: sharper attention (larger logits) than the default synthetic weights
            sd = synthetic.random_vit_state_dict(dims, seed=0, std=std)
            ref = forward(sd, px, dims, "fp32")
            for lag in (0.0, 8.0):
                a = forward(sd, px, dims, "today", lag)
                b = forward(sd, px, dims, "halfexp", lag)
                print(f"weight std {std}: lag {lag:3.1f}: embedding rel-L2 vs fp32  P = fp16(2^x) {rel(a, ref):.3e}   "
                      f"P = 2^fp16(x) {rel(b, ref):.3e}", flush=True)
