"""A/B of the attention tilings on one box: same inputs, alternating variants."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pigeon_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
views = int(sys.argv[1]) if len(sys.argv) > 1 else 128
qkv = torch.randn(views * 577, 3072, device=dev).half()


def timeit(fn, iters=10, warm=3):
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


ref = None
for rep in range(2):
    for var, poly in (("64", ""), ("32", ""), ("64s", ""), ("64h", ""), ("32c", "")):
        os.environ["PG_ATTN_VARIANT"] = var
        os.environ["PG_ATTN_POLY"] = poly
        out = ops.attention_f16(qkv, views, 577, 16)
        if ref is None:
            ref = out.clone()
        ms = timeit(lambda: ops.attention_f16(qkv, views, 577, 16))
        err = ((out.float() - ref.float()).norm() / ref.float().norm()).item()
        print(f"variant KV={var} poly={poly or 0}: {ms:.4f} ms  ({4.0 * 577 * 577 * 64 * 16 * views / ms / 1e9:.0f} TF/s)  rel diff vs KV=64: {err:.2e}", flush=True)
