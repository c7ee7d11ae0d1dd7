"""Probe which (LBO, SBO) descriptor strides make MN-major UMMA operand tiles of more than one 64-column chunk work."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pigeon_b200._lib import check, current_stream_ptr, load, ptr  # noqa: E402

lib = load()
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
M, N, K = 512, 512, 640
a = (torch.randn(K, M, generator=g) * 0.1).to(dev, torch.bfloat16)
w = (torch.randn(K, N, generator=g) * 0.1).to(dev, torch.bfloat16)
ref = a.double().t() @ w.double()
for lbo, sbo in ((8192, 1024), (1024, 8192), (1024, 1024), (8192, 8192), (128, 1024), (1024, 128), (16, 1024), (4096, 1024),
                 (1024, 4096), (8192, 128), (128, 8192)):
    os.environ["PG_MN_LBO"], os.environ["PG_MN_SBO"] = str(lbo), str(sbo)
    out = torch.zeros(M, N, device=dev)
    try:
        check(lib.pg_gemm_tn(ptr(a), M, ptr(w), N, ptr(out), N, M, N, K, 0, 1, current_stream_ptr()), "pg_gemm_tn")
        torch.cuda.synchronize()
        err = ((out.double() - ref).norm() / ref.norm()).item()
    except Exception as e:  # noqa: BLE001
        err = str(e)[:80]
    print(f"LBO={lbo:5d} SBO={sbo:5d}: rel err {err}", flush=True)

os.environ.pop("PG_MN_LBO"); os.environ.pop("PG_MN_SBO")   # library defaults from here on
w16 = w.to(torch.float16)
ref2 = a.double().t() @ w16.double()
for mode, (aa, ww, rr) in {1: (a, w, ref), 0: (a.to(torch.float16), w16, None)}.items():   # mixed types (2, 3) trap: illegal instruction
    rr = rr if rr is not None else aa.double().t() @ ww.double()
    MM, NN = aa.shape[1], ww.shape[1]
    out = torch.zeros(MM, NN, device=dev)
    check(lib.pg_gemm_tn(ptr(aa), MM, ptr(ww), NN, ptr(out), NN, MM, NN, K, 0, mode, current_stream_ptr()), "pg_gemm_tn")
    torch.cuda.synchronize()
    print(f"operand mode {mode}: rel err {((out.double() - rr).norm() / rr.norm()).item():.3e}", flush=True)
# accumulate + large K with a ragged tail
K2 = 36928 + 40
a2 = (torch.randn(K2, 1024, generator=g) * 0.01).to(dev, torch.bfloat16)
w2 = (torch.randn(K2, 4096, generator=g) * 0.5).to(dev, torch.bfloat16)
acc0 = torch.randn(1024, 4096, generator=g).to(dev)
out = acc0.clone()
check(lib.pg_gemm_tn(ptr(a2), 1024, ptr(w2), 4096, ptr(out), 4096, 1024, 4096, K2, 1, 1, current_stream_ptr()), "pg_gemm_tn")
torch.cuda.synchronize()
r = acc0.double() + a2.double().t() @ w2.double()
print(f"accumulate, K={K2}: rel err {((out.double() - r).norm() / r.norm()).item():.3e}")
s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s_.record()
for _ in range(5):
    check(lib.pg_gemm_tn(ptr(a2), 1024, ptr(w2), 4096, ptr(out), 4096, 1024, 4096, K2, 1, 1, current_stream_ptr()), "pg_gemm_tn")
e_.record(); torch.cuda.synchronize()
ms = s_.elapsed_time(e_) / 5
print(f"dW2-shaped weight gradient (1024 x 4096 x {K2}): {ms:.3f} ms = {2 * 1024 * 4096 * K2 / ms / 1e9:.0f} TF/s")
