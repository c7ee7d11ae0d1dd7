"""Multi-rank result parity of the inference path on 2+ GPUs (torchrun): every rank embeds its shard of a
fixed global batch, the head outputs are exchanged with the packed NCCL all-gather, retrieval runs over (a) a replicated
bank and (b) a bank whose geocells are sharded across the ranks (second all-gather of the per-candidate partials) — and
both gathered results must equal, BIT FOR BIT, what one rank computes alone on the whole batch.  Exit code 0 on success."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _say(rank, msg):
    print(msg, flush=True)
    d = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, f"ddp_infer_check_rank{rank}.log"), "a") as f:
            f.write(msg + "\n")


def main():
    import traceback
    try:
        _main()
    except Exception:
        traceback.print_exc()
        sys.stdout.flush()
        sys.exit(2)


def _main():
    from pigeon_b200 import CLIPVisionTower, ProtoRefiner, SuperGuessr, VitDims, evaluation, synthetic
    from pigeon_b200 import dist as pdist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    dims = VitDims(image_size=112, patch_size=14, hidden=256, heads=4, intermediate=768, layers=3)
    C, P, k = 64, 3000, 5
    torch.manual_seed(1234)          # every rank must hold the SAME geocell head (nn.Linear's default init draws from the global RNG)
    tower = CLIPVisionTower(dims)
    tower.load_state_dict(synthetic.random_vit_state_dict(dims, seed=4, std=0.05))
    cells = synthetic.synthetic_geocells(C, 0)
    model = SuperGuessr(tower, panorama=True, freeze_base=True, num_candidates=12, geocells=cells).to(dev).eval()
    bank = synthetic.synthetic_bank(C, P, dims.hidden, seed=6, members_mean=2.5, empty_cells=3)
    mk = lambda shard: ProtoRefiner(topk=k, max_refinement=1e6, temperature=1.6, protos=bank, device=dev, shard_cells=shard).eval()
    Bg = 6 * world
    px = torch.randn(Bg, 12, dims.image_size, dims.image_size, generator=torch.Generator().manual_seed(9)).half().to(dev)
    labels = torch.from_numpy(synthetic.synthetic_geocells(Bg, 3)).to(dev)
    clf = (torch.arange(Bg) % C).to(dev)
    lo, hi = pdist.shard_range(Bg, rank, world)
    mine = dict(pixel_values=px[lo:hi], labels=labels[lo:hi], labels_clf=clf[lo:hi])
    ok = True
    # single-rank reference on the whole batch (no collectives)
    ll1, cell1, out1 = evaluation.predict_batch(model, mk(False), dict(pixel_values=px, labels=labels, labels_clf=clf), gather=False)
    for name, shard in (("replicated bank", False), ("cell-sharded bank", True)):
        ll, cell, out = evaluation.predict_batch(model, mk(shard), mine)
        e_emb = torch.equal(out.embedding, out1.embedding[lo:hi])
        e_top = torch.equal(out.top5_geocells.indices, out1.top5_geocells.indices[lo:hi])
        e_ll = ll.shape == ll1.shape and torch.equal(ll, ll1)
        e_cell = cell.shape == cell1.shape and torch.equal(cell, cell1)
        same = e_emb and e_top and e_ll and e_cell
        _say(rank, f"rank {rank} world {world} {name}: gathered outputs == single-rank outputs: {same} "
              f"(embedding {e_emb}, candidates {e_top}, preds_LLH {e_ll} {tuple(ll.shape)} vs {tuple(ll1.shape)}, "
              f"geocell {e_cell}; embedding max |diff| {(out.embedding - out1.embedding[lo:hi]).abs().max().item():.3e}; "
              f"preds_LLH max |diff| {(ll - ll1).abs().max().item() if ll.shape == ll1.shape else -1:.3e}, "
              f"geocell mismatches {(cell != cell1).sum().item() if cell.shape == cell1.shape else -1})")
        ok = ok and same
    # the C-ABI exchange entry point on its own communicator (include/pigeon_b200.h: pg_nccl_*, pg_allgather_embeddings): the
    # 128-byte id travels over torch.distributed, the gathered bytes must equal torch's all_gather of the same buffers
    import ctypes as Ct
    from pigeon_b200._lib import check, current_stream_ptr, load, ptr
    lib = load()
    idbuf = (Ct.c_char * 128)()
    if rank == 0:
        check(lib.pg_nccl_unique_id(idbuf), "pg_nccl_unique_id")
    idt = torch.frombuffer(bytearray(idbuf.raw), dtype=torch.uint8).to(dev)
    dist.broadcast(idt, 0)
    idbuf = (Ct.c_char * 128).from_buffer_copy(bytes(idt.cpu().numpy().tobytes()))
    comm = Ct.c_void_p()
    check(lib.pg_nccl_comm_create(idbuf, world, rank, Ct.byref(comm)), "pg_nccl_comm_create")
    send = out1.embedding[lo:hi].contiguous()
    recv = torch.empty((world,) + tuple(send.shape), dtype=send.dtype, device=dev)
    check(lib.pg_allgather_embeddings(comm, ptr(send), ptr(recv), send.numel() * send.element_size(), current_stream_ptr()),
          "pg_allgather_embeddings")
    torch.cuda.synchronize()
    theirs = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(theirs, send)
    e_gather = torch.equal(recv, torch.stack(theirs)) and torch.equal(recv.reshape(out1.embedding.shape), out1.embedding)
    lib.pg_nccl_comm_destroy(comm)
    _say(rank, f"rank {rank} world {world} pg_allgather_embeddings == torch.distributed.all_gather == whole-batch embedding: {e_gather}")
    ok = ok and e_gather
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
