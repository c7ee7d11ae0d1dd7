"""Data-parallel fine-tune step on 2+ GPUs (torchrun): every rank takes its shard of the golden batch, runs the training
forward/backward (the gradient all-reduce is inside SuperGuessr.backward / TowerTrainer.finalize) and rank 0 checks the
averaged gradients against tests/golden/train_tower_small.npz — the single-process full-batch gradients of the reference.
Exit code 0 on success."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl")
    import test_gpu_train_tower as T
    z, meta = T._load("train_tower_small")
    n = meta["n_views"]
    assert n % world == 0, "golden batch must split evenly"
    sg, px, dims = T._model(meta, z, dev)
    sg.train()
    lo, hi = rank * n // world, (rank + 1) * n // world
    out = sg(pixel_values=px[lo:hi].to(dev), labels=torch.tensor(z["labels"][lo:hi]),
             labels_clf=torch.tensor(z["labels_clf"][lo:hi]))
    sg.backward(out.loss)
    loss = out.loss.detach().double().reshape(1).to(dev)
    dist.all_reduce(loss)
    ok = True
    if rank == 0:
        try:
            np.testing.assert_allclose(loss.item() / world, float(z["loss"]), rtol=1e-3)
            worst = T._check_grads(sg, z, meta)
            print(f"ddp world={world}: loss ok, worst per-tensor relative gradient error {max(worst.values()):.3e}", flush=True)
        except AssertionError as e:
            print("FAILED:", e, flush=True)
            ok = False
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
