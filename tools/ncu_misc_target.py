"""Small fixed workload for `ncu --set full` captures of the kernels around the tower: geocell head (view mean + split,
head GEMM, softmax / top-k, CE loss), ProtoRefiner (pool, pair sort, cell-major scan at BASELINE.json configs[4] scale,
final stage), image pre-processing (coefficients, horizontal, vertical pass) and the AdamW update."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pigeon_b200 import ProtoRefiner, SuperGuessr, ops, synthetic  # noqa: E402
from pigeon_b200.preprocess import ClipImageProcessor  # noqa: E402
from pigeon_b200.training import AdamW  # noqa: E402

dev = torch.device("cuda:0")
what = sys.argv[1] if len(sys.argv) > 1 else "all"
torch.manual_seed(0)

if what in ("all", "head"):
    C, D, B = 1000, 1024, 256
    sg = SuperGuessr(None, panorama=True, num_candidates=50, should_smooth_labels=True, geocells=synthetic.synthetic_geocells(C, 0)).to(dev).eval()
    bank = synthetic.synthetic_bank(C, 100_000, D, seed=2, members_mean=0.0, empty_cells=5)
    ref = ProtoRefiner(topk=5, max_refinement=1000, temperature=1.6, protos=bank, device=dev).eval()
    emb = (torch.randn(B, 4, D) * 0.3).to(dev)
    labels = torch.from_numpy(synthetic.synthetic_geocells(B, 7)).to(dev)
    clf = (torch.arange(B) % C).to(dev)
    for _ in range(2):
        out = sg(embedding=emb, labels=labels, labels_clf=clf)
        ref(out.embedding, initial_preds=out.preds_LLH, candidate_cells=out.top5_geocells.indices, candidate_probs=out.top5_geocells.values)

if what in ("all", "refiner"):
    C, P, D, B, k = 2076, 1_000_000, 768, 8192, 5
    rng = np.random.default_rng(0)
    sizes = rng.multinomial(P, np.ones(C) / C)
    g = torch.Generator(device=dev).manual_seed(0)
    bank = ops.DeviceBank(dev, cell_off=torch.from_numpy(np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)),
                          proto_emb=torch.randn(P, D, generator=g, device=dev) * 0.3, proto_lnglat=torch.rand(P, 2, generator=g, device=dev) * 90,
                          proto_count=torch.ones(P, dtype=torch.int32), member_off=torch.arange(P + 1),
                          member_idx=torch.zeros(P, dtype=torch.int64), data_emb=torch.zeros(1, D), data_lnglat=torch.zeros(1, 2))
    cand = torch.from_numpy(np.stack([rng.choice(C, size=k, replace=False) for _ in range(B)]).astype(np.int64)).to(dev)
    probs = torch.from_numpy(-np.sort(-rng.dirichlet(np.ones(k), size=B), axis=1).astype(np.float32)).to(dev)
    q = torch.randn(B, 1, D, device=dev) * 0.3
    init = torch.rand(B, 2, device=dev, dtype=torch.float64) * 90
    for _ in range(2):
        ops.refiner_forward(bank, q, init, cand, probs, k, 1.6, 1000.0)

if what in ("all", "preprocess"):
    proc = ClipImageProcessor(size=336, device=dev)
    imgs = [torch.randint(0, 256, (480, 640, 3), dtype=torch.uint8, device=dev) for _ in range(64)]
    for _ in range(2):
        proc.preprocess_device(imgs)

if what in ("all", "adamw"):
    p = torch.nn.Parameter(torch.randn(16 * 1024 * 1024, device=dev))
    opt = AdamW([p], lr=2e-5)
    for _ in range(2):
        p.grad = torch.randn_like(p)
        opt.step()
torch.cuda.synchronize()
print("done")
