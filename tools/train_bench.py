"""Fine-tune step timing at the BASELINE.json configs[3] per-rank shape (cfg4: 128 four-view samples = 512 views per
GPU, ViT-L/14-336, haversine-smoothed CE, AdamW lr 2e-5) with the reference freeze policy
(models/super_guessr.py:159-160: embeddings + pre_layrnorm + last encoder layer + head trainable) or everything trainable.

    python tools/train_bench.py [--samples 128] [--steps 3] [--warmup 1] [--all-trainable] [--chunk-views 64]
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/train_bench.py   # per-rank batch fixed
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pigeon_b200 import CLIPVisionTower, SuperGuessr, VitDims, synthetic  # noqa: E402
from pigeon_b200._lib import load  # noqa: E402
from pigeon_b200.training import AdamW  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--samples", type=int, default=128)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=1)
ap.add_argument("--all-trainable", action="store_true")
ap.add_argument("--chunk-views", type=int, default=64)
ap.add_argument("--out", default="gpurun_out/train_bench.json")
args = ap.parse_args()

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:                      # torchrun: one rank per GPU, per-rank batch fixed (weak scaling), NCCL gradient averaging
    import torch.distributed as dist
    dist.init_process_group("nccl")
dims = VitDims()
tower = CLIPVisionTower(dims)
tower.load_state_dict(synthetic.random_vit_state_dict(dims, seed=0), strict=True)
C_cells = 1000
sg = SuperGuessr(tower, panorama=True, should_smooth_labels=True, num_candidates=5,
                 geocells=synthetic.synthetic_geocells(C_cells, 0)).to(dev)
if not args.all_trainable:
    for p in sg.base_model.vision_model.encoder.layers[:-1].parameters():
        p.requires_grad = False
sg.max_train_views = args.chunk_views
sg.train()
n_train = sum(p.numel() for p in sg.parameters() if p.requires_grad)
opt = AdamW(sg.parameters(), lr=2e-5)
B = args.samples
torch.manual_seed(1234 + rank)
px = torch.randn(B, 12, 336, 336, device=dev).half()
labels = torch.tensor(synthetic.synthetic_geocells(B, 5 + rank))
labels_clf = torch.randint(0, C_cells, (B,))
lib = load()


def step():
    out = sg(pixel_values=px, labels=labels, labels_clf=labels_clf)
    sg.backward(out.loss)
    opt.step()
    opt.zero_grad()
    return out.loss


for _ in range(args.warmup):
    loss = step()
torch.cuda.synchronize()
lib.pg_profile_begin()
step()
n = lib.pg_profile_end()
names = (C.c_char_p * n)()
ms = (C.c_float * n)()
cnt = (C.c_int32 * n)()
lib.pg_profile_read(names, ms, cnt, n)
prof = sorted(((names[i].decode(), float(ms[i]), int(cnt[i])) for i in range(n)), key=lambda t: -t[1])
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.time()
s.record()
losses = [step() for _ in range(args.steps)]
e.record()
torch.cuda.synchronize()
wall = (time.time() - t0) / args.steps
dt = s.elapsed_time(e) / args.steps
if world > 1:                      # slowest rank
    t = torch.tensor([dt], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = t.item()
    dist.barrier()
res = dict(config=f"fine-tune step, ViT-L/14-336, {B} four-view samples ({4 * B} views), "
                  f"{'all parameters' if args.all_trainable else 'reference freeze policy'} trainable ({n_train / 1e6:.1f} M), "
                  f"chunks of {args.chunk_views} views, AdamW",
           n_gpus=world, ms_per_step=dt, wall_ms_per_step=wall * 1e3, samples_per_s=world * B / dt * 1e3,
           grad_allreduce_bytes=4 * n_train, losses=[float(l) for l in losses],
           max_mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30,
           profile=[dict(name=a, ms=b, launches=c) for a, b, c in prof[:24]], profile_total_ms=sum(b for _, b, _ in prof))
if rank == 0:
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)
    print(json.dumps({k: v for k, v in res.items() if k != "profile"}))
    for a, b, c in prof[:16]:
        print(f"  {a:28s} {b:9.2f} ms  {c:5d} launches")
if world > 1:
    dist.destroy_process_group()
