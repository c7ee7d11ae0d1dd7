"""Per-kernel device timings at ViT-L/14-336 shapes (CUDA events, L2 flushed between iterations).
Development aid; bench.py is the contract benchmark."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pigeon_b200 import _lib, ops, synthetic  # noqa: E402
from pigeon_b200.vit_engine import VitDims, VitEngine  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


res = {}
views = int(sys.argv[1]) if len(sys.argv) > 1 else 128
M = views * 577
for name, N, K, epi in [("qkv", 3072, 1024, _lib.EPI_F16_BIAS), ("out_proj", 1024, 1024, _lib.EPI_F32_BIAS_RESID),
                        ("fc1", 4096, 1024, _lib.EPI_F16_BIAS_QGELU), ("fc2", 1024, 4096, _lib.EPI_F32_BIAS_RESID)]:
    a = torch.randn(M, K, device=dev).half()
    w = (torch.randn(N, K, device=dev) * 0.02).half()
    b = torch.randn(N, device=dev)
    out = torch.zeros(M, N, device=dev, dtype=torch.float16 if epi in (0, 1) else torch.float32)
    ms = timeit(lambda: ops.gemm_f16(a, w, b, epi, out=out))
    tf = 2.0 * M * N * K / ms / 1e9
    ms_t = timeit(lambda: torch.matmul(a, w.t()))
    res[f"gemm_{name}"] = dict(ms=ms, tflops=tf, torch_ms=ms_t, torch_tflops=2.0 * M * N * K / ms_t / 1e9)
    print(name, res[f"gemm_{name}"], flush=True)
    del a, w, out

qkv = torch.randn(M, 3072, device=dev).half()
ms = timeit(lambda: ops.attention_f16(qkv, views, 577, 16))
res["attention"] = dict(ms=ms, tflops=4.0 * 577 * 577 * 64 * 16 * views / ms / 1e9, us_per_view=ms * 1e3 / views)
print("attention", res["attention"], flush=True)
del qkv
x = torch.randn(M, 1024, device=dev)
g = torch.ones(1024, device=dev)
ms = timeit(lambda: ops.layernorm_f16(x, g, g, 1e-5))
res["layernorm"] = dict(ms=ms, gbs=M * 1024 * 6 / ms / 1e6)
print("layernorm", res["layernorm"], flush=True)
del x

dims = VitDims()
eng = VitEngine(synthetic.random_vit_state_dict(dims, 0), dims, device=dev, max_views_per_pass=views)
px = torch.randn(views, 3, 336, 336, device=dev).half()
ms = timeit(lambda: eng.forward(px), iters=5, warm=2)
res["vit_forward"] = dict(ms=ms, views_per_s=views / ms * 1e3, images_per_s=views / 4 / ms * 1e3,
                          tflops=381.92e9 * views / ms / 1e9)
print("vit", res["vit_forward"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/microbench.json", "w"), indent=1)
