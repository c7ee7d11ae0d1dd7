"""ProtoRefiner-only sweep (BASELINE.json configs[4]): P prototypes, D dims, C geocells, B queries, top-k candidates.
Synthetic bank generated on the device (count == 1 clusters).  Reports ms, achieved GB/s against the ALGORITHMIC bytes
(each touched geocell segment once + queries + candidates + outputs) and the fp32 element-pair rate."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pigeon_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=5, warm=2):
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


def make_bank(C, P, D, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    rng = np.random.default_rng(seed)
    sizes = rng.multinomial(P, np.ones(C) / C)
    cell_off = torch.from_numpy(np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64))
    emb = torch.randn(P, D, generator=g, device=dev) * 0.3
    ll = torch.rand(P, 2, generator=g, device=dev) * 90
    one = torch.zeros(1, D, device=dev)
    return ops.DeviceBank(dev, cell_off=cell_off, proto_emb=emb, proto_lnglat=ll, proto_count=torch.ones(P, dtype=torch.int32),
                          member_off=torch.arange(P + 1), member_idx=torch.zeros(P, dtype=torch.int64), data_emb=one,
                          data_lnglat=torch.zeros(1, 2)), sizes


def main():
    res = []
    for (C, P, D, B, k) in [(2076, 1_000_000, 768, 8192, 5), (2076, 1_000_000, 768, 8192, 40), (2076, 1_000_000, 1024, 8192, 5),
                            (1000, 100_000, 1024, 2048, 5), (1000, 100_000, 1024, 256, 5)]:
        bank, sizes = make_bank(C, P, D)
        rng = np.random.default_rng(3)
        cand = np.stack([rng.choice(C, size=k, replace=False) for _ in range(B)]).astype(np.int64)
        probs = -np.sort(-rng.dirichlet(np.ones(k), size=B), axis=1).astype(np.float32)
        emb = torch.randn(B, 1, D, device=dev) * 0.3
        init = torch.rand(B, 2, device=dev, dtype=torch.float64) * 90
        candt, probst = torch.from_numpy(cand).to(dev), torch.from_numpy(probs).to(dev)
        touched = np.unique(cand)
        alg_bytes = float(sizes[touched].sum()) * D * 4 + B * D * 4 + B * k * 12 + B * 32
        pairs_elems = float(sizes[cand].sum()) * D           # (prototype, query) element pairs
        row = dict(C=C, P=P, D=D, B=B, k=k, algorithmic_GB=alg_bytes / 1e9)
        for mode in ("cell_major", "query_major"):
            ops.refiner_set_schedule(1 if mode == "query_major" else 2)
            if mode == "query_major" and pairs_elems > 2e11:
                continue
            ms = timeit(lambda: ops.refiner_forward(bank, emb, init, candt, probst, k, 1.6, 1000.0))
            row[mode] = dict(ms=ms, GBps_vs_algorithmic=alg_bytes / ms / 1e6, Telem_pairs_per_s=pairs_elems / ms / 1e9)
        print(row, flush=True)
        res.append(row)
        del bank
        torch.cuda.empty_cache()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/refiner_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
