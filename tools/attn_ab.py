"""Correctness + A/B timing of the attention kernels on one box.

    python tools/attn_ab.py check            # every variant vs a torch fp32 softmax at several geometries (exit 1 on mismatch)
    python tools/attn_ab.py time [views]     # alternating timings of the variants on the same inputs (L2 flushed)

Variants: (0, p) = pair kernel with p eighths of the exponentials on the FMA pipe, (2, p) = split kernel, (3, p) = fold kernel,
(1, 0) = first-generation kernel.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pigeon_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
VARIANTS = [(1, 0), (0, 2), (2, 2), (3, 0), (3, 2), (3, 3), (3, 4)]
if os.environ.get("ATTN_AB_VARIANTS"):   # e.g. "3:3,0:2"
    VARIANTS = [tuple(int(x) for x in v.split(":")) for v in os.environ["ATTN_AB_VARIANTS"].split(",")]


def reference(qkv, n_views, seq, heads):
    x = qkv.float().view(n_views, seq, 3, heads, 64)
    q, k, v = (x[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    att = torch.softmax(q @ k.transpose(-1, -2) * 0.125, dim=-1)
    return (att @ v).permute(0, 2, 1, 3).reshape(n_views * seq, heads * 64)


def check():
    bad = 0
    geoms = [(1, 128, 1), (1, 64, 2), (2, 577, 2), (3, 200, 1), (1, 577, 16), (2, 65, 1), (1, 129, 1), (1, 17, 4),
             (3, 577, 3), (2, 256, 5), (40, 577, 16), (2, 32, 2), (1, 33, 3), (2, 96, 1), (1, 100, 2), (1, 48, 1)]
    for (n_views, seq, heads) in geoms:
        g = torch.Generator(device="cpu").manual_seed(seq * 3 + heads)
        qkv = (torch.randn(n_views * seq, 3 * heads * 64, generator=g) * 1.5).half().to(dev)
        ref = reference(qkv, n_views, seq, heads)
        for (var, poly) in VARIANTS:
            out = ops.attention_f16(qkv, n_views, seq, heads, variant=var, poly=poly)
            torch.cuda.synchronize()
            err = ((out.float() - ref).norm() / ref.norm()).item()
            ok = err < 2e-3 and bool(torch.isfinite(out.float()).all())
            bad += not ok
            print(f"geom {(n_views, seq, heads)} variant {var} poly {poly}: rel err {err:.2e} {'ok' if ok else 'FAIL'}", flush=True)
    # growing logits: the exact-maximum / rescale path
    n_views, seq, heads = 2, 577, 2
    g = torch.Generator(device="cpu").manual_seed(77)
    x = torch.randn(n_views, seq, 3, heads, 64, generator=g)
    x[:, :, 1] *= (0.25 + 6.0 * torch.arange(seq) / seq).view(1, seq, 1, 1)
    qkv = x.reshape(n_views * seq, 3 * heads * 64).half().to(dev)
    ref = reference(qkv, n_views, seq, heads)
    for (var, poly) in VARIANTS:
        out = ops.attention_f16(qkv, n_views, seq, heads, variant=var, poly=poly)
        err = ((out.float() - ref).norm() / ref.norm()).item()
        ok = err < 2e-3 and bool(torch.isfinite(out.float()).all())
        bad += not ok
        print(f"growing logits variant {var} poly {poly}: rel err {err:.2e} {'ok' if ok else 'FAIL'}", flush=True)
    # rows whose maximum sits well above the first block's (inside the fp16 window of the fold kernel: no repair path), and the
    # log-sum-exp side output of every variant against torch
    x = torch.randn(n_views, seq, 3, heads, 64, generator=g)
    x[:, :, 1] *= (0.5 + 1.5 * torch.arange(seq) / seq).view(1, seq, 1, 1)
    qkv2 = x.reshape(n_views * seq, 3 * heads * 64).half().to(dev)
    ref2 = reference(qkv2, n_views, seq, heads)
    xf = qkv2.float().view(n_views, seq, 3, heads, 64)
    q, k = (xf[:, :, i].permute(0, 2, 1, 3) for i in range(2))
    lse_ref = torch.logsumexp(q @ k.transpose(-1, -2) * 0.125, dim=-1) * 1.4426950408889634      # [views, heads, seq], log2 units
    for (var, poly) in VARIANTS:
        for name, inp, rf in (("moderate spread", qkv2, ref2), ("growing logits", qkv, ref)):
            out, lse = ops.attention_f16(inp, n_views, seq, heads, variant=var, poly=poly, return_lse2=True)
            err = ((out.float() - rf).norm() / rf.norm()).item()
            ok = err < 2e-3 and bool(torch.isfinite(out.float()).all())
            if name == "moderate spread":
                lerr = (lse.view(n_views, heads, seq) - lse_ref).abs().max().item()
                ok = ok and lerr < 2e-2
                print(f"{name} variant {var} poly {poly}: rel err {err:.2e} lse2 max abs err {lerr:.2e} {'ok' if ok else 'FAIL'}", flush=True)
            else:
                print(f"{name} (lse2 requested) variant {var} poly {poly}: rel err {err:.2e} {'ok' if ok else 'FAIL'}", flush=True)
            bad += not ok
    # run-to-run determinism of the default variant
    for var in (0, 2, 3):
        a = ops.attention_f16(qkv, n_views, seq, heads, variant=var)
        b = ops.attention_f16(qkv, n_views, seq, heads, variant=var)
        if not torch.equal(a, b):
            bad += 1
            print(f"variant {var} is not run-to-run deterministic: FAIL")
    return bad


def timeit(fn, iters=10, warm=3):
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return sorted(ts)[len(ts) // 2]


def time_variants(views):
    qkv = torch.randn(views * 577, 3072, device=dev).half()
    flops = 4.0 * 577 * 577 * 64 * 16 * views
    for rep in range(2):
        for (var, poly) in VARIANTS:
            ms = timeit(lambda: ops.attention_f16(qkv, views, 577, 16, variant=var, poly=poly))
            print(f"views {views} variant {var} poly {poly}: {ms:.4f} ms  {flops / ms / 1e9:.0f} TF/s", flush=True)


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "check"
    if mode == "check":
        sys.exit(1 if check() else 0)
    time_variants(int(sys.argv[2]) if len(sys.argv) > 2 else 128)
