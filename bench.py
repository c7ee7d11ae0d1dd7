#!/usr/bin/env python
"""Contract benchmark: four-view 336x336 images/sec through ViT-L/14 + geocell head + ProtoRefiner.

    python bench.py --gpus N --steps K --warmup W            # this repo's B200 path (N=1 default)
    torchrun --nproc-per-node N ... bench.py --gpus N ...    # one rank per GPU, NCCL all-gather before retrieval
    python bench.py --impl reference ...                     # CPU arm: the oracle port of the reference path

One step = one pass of the hot path over one batch of synthetic input:
  N = 1 : BASELINE.json configs[1] — batch 256 four-view panoramas (1024 views), fp16 operands, + head + refine
  N > 1 : configs[2] scaled by rank — 256 samples per rank, one all-gather of per-rank head outputs, every rank
          refines the gathered batch against a replicated 100k-prototype bank (weak scaling).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_VIEW = 381.92e9          # SURVEY.md §8 / BASELINE.md §2 (un-padded, 2*MAC)
GEMM_FLOP_PER_VIEW = 87.12e9 + 29.04e9 + 232.33e9 + 0.69e9
ATTN_FLOP_PER_VIEW = 32.73e9
NUM_CELLS, NUM_PROTOS, TOPK, NUM_CAND = 1000, 100_000, 5, 50
REFINER_T, REFINER_MAX_KM = 1.6, 1000.0


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], tflops=d["bf16_tflops"], tflops_sustained=d["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, tflops=1590.0, tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, f"/tmp/pg_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        # under load = samples within 40% of the busiest clock seen (idle samples before/after are dropped)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def build_models(device, seed=0, fold_layernorm=True):
    from pigeon_b200 import CLIPVisionTower, ProtoRefiner, SuperGuessr, VitDims, synthetic
    dims = VitDims()
    tower = CLIPVisionTower(dims)
    tower.load_state_dict(synthetic.random_vit_state_dict(dims, seed=seed))
    tower.max_views_per_pass = 1024
    tower.fold_layernorm = fold_layernorm
    cells = synthetic.synthetic_geocells(NUM_CELLS, 0)
    model = SuperGuessr(tower, panorama=True, freeze_base=True, num_candidates=NUM_CAND, geocells=cells).to(device).eval()
    bank = synthetic.synthetic_bank(NUM_CELLS, NUM_PROTOS, dims.hidden, seed=2, members_mean=0.0, empty_cells=5)
    refiner = ProtoRefiner(topk=TOPK, max_refinement=REFINER_MAX_KM, temperature=REFINER_T, protos=bank, device=device).eval()
    return model, refiner, dims, cells, bank


def run_ours(args):
    import torch.distributed as dist
    from pigeon_b200 import evaluation, synthetic
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    peaks = load_peaks()
    B = args.batch
    model, refiner, dims, cells, bank = build_models(dev, fold_layernorm=not args.no_ln_fold)
    g = torch.Generator().manual_seed(1 + rank)
    # synthetic panoramas, fp16, (B, 12, 336, 336): view index fastest inside a sample (dataset_preprocessing.py:199-200)
    px_host = torch.randn(B, 12, dims.image_size, dims.image_size, generator=g, dtype=torch.float32).half().pin_memory()
    px_dev = px_host.to(dev)
    labels = torch.from_numpy(synthetic.synthetic_geocells(B, 7 + rank)).to(dev)
    labels_clf = (torch.arange(B) % NUM_CELLS).to(dev)

    def step_device():
        return evaluation.predict_batch(model, refiner, dict(pixel_values=px_dev, labels=labels, labels_clf=labels_clf))

    def step_e2e():
        ll, cell, out = evaluation.predict_batch(model, refiner, dict(pixel_values=px_host, labels=labels, labels_clf=labels_clf))
        return ll.cpu(), cell.cpu(), out.loss.cpu()          # device -> host read of the step's result

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            r = fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)       # max over ranks
        return ms.item(), r

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, res = timed(step_device, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e, _ = timed(step_e2e, max(2, min(args.steps, 5)), 1)
    e2e_steps = max(2, min(args.steps, 5))

    # per-kernel-family device time of ONE step (CUDA events recorded by the library on the launch stream).
    # Every rank runs it (the step contains the all-gather); only rank 0 reports.
    prof = profile_step(model, refiner, px_dev, labels, labels_clf)
    if rank != 0:
        prof = None

    images = B * world
    value = images * args.steps / (ms / 1e3)
    out = {
        "metric": "four-view 336x336 images/sec (ViT-L/14 + geocell head + ProtoRefiner)",
        "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp16", "data": "synthetic (randn panoramas, random-init ViT-L/14-336, 1000 synthetic geocells, "
                                 "100k synthetic prototypes)",
        "config": {"workload": f"batch={B}/GPU four-view synthetic panoramas ({4 * B} views), ViT-L/14-336 embed + "
                               f"geocell head (C={NUM_CELLS}, top-{NUM_CAND}) + ProtoRefiner (P={NUM_PROTOS}, topk={TOPK}); "
                               "BASELINE.json configs[1]" + ("" if world == 1 else " per rank = configs[2] sharded"),
                   "global_batch": images, "views_per_gpu": 4 * B, "parallelism": f"dp{world}",
                   "l2": "inputs larger than L2 (0.69 GB pixels + >2 GB activations per step); no explicit flush",
                   "collective": "none" if world == 1 else "1 all_gather_into_tensor (NCCL) of per-rank head outputs per step"},
        "e2e": {"value": images * e2e_steps / (ms_e2e / 1e3), "unit": "images/s",
                "h2d_bytes_per_step": px_host.numel() * 2, "d2h_bytes_per_step": B * world * (8 + 8) + 4,
                "note": "SuperGuessr.forward + ProtoRefiner.forward with pinned HOST fp16 pixels, results read back"},
        "gpu_launches": launches_per_step(B, dims) * args.steps,
        "clocks": clocks,
        "tflops_per_gpu": FLOP_PER_VIEW * 4 * B * args.steps / (ms / 1e3) / 1e12,
    }
    if prof is not None:
        gemm_ms, attn_ms = prof["gemm_ms"], prof["attention_ms"]
        views = 4 * B
        ach = GEMM_FLOP_PER_VIEW * views / (gemm_ms / 1e3) / 1e12
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("gemm_dram_bytes_per_launch")
        out["roofline"] = {"kernel": "gemm_f16_kernel (tcgen05 GEMM, all projection/MLP launches of one step)",
                           "bound": "tensor", "achieved": ach, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                           "frac": ach / peaks["tflops_sustained"], "traffic": traffic,
                           "peak_source": peaks["source"] + ", sustained bf16 cuBLAS figure (kernel timed inside a long step)",
                           "launches": prof["gemm_launches"], "avg_launch_ms": gemm_ms / prof["gemm_launches"],
                           "share_of_step": gemm_ms / prof["total_ms"]}
        a_ach = ATTN_FLOP_PER_VIEW * views / (attn_ms / 1e3) / 1e12
        out["roofline_attention"] = {"kernel": "attention_kernel (tcgen05 QK^T / PV)", "bound": "tensor", "achieved": a_ach,
                                     "peak": peaks["tflops_sustained"], "unit": "TFLOP/s", "frac": a_ach / peaks["tflops_sustained"],
                                     "share_of_step": attn_ms / prof["total_ms"]}
        out["kernel_ms_per_step"] = {k: round(v, 3) for k, v in prof.items() if k.endswith("_ms")}
        out["gpu_launches"] = prof["launches"] * args.steps     # counted by the library's launch hooks, per step
        out["launches_by_family_per_step"] = prof["families"]
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(sample_images=args.cpu_sample)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def launches_per_step(B, dims):
    passes = -(-4 * B // 1024)
    vit = passes * (3 + dims.layers * 7 + 1)
    return vit + 3 + 2 + 3   # head (split, gemm, softmax/top-k) + loss (per-sample, mean) + refiner (pool, scan, finalize)


def profile_step(model, refiner, px_dev, labels, labels_clf):
    """One extra step with per-launch CUDA events inside the library (pg_profile_*), after the timed region."""
    from pigeon_b200 import _lib, evaluation
    lib = _lib.load()
    if not hasattr(lib, "pg_profile_begin"):
        return None
    lib.pg_profile_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    evaluation.predict_batch(model, refiner, dict(pixel_values=px_dev, labels=labels, labels_clf=labels_clf))
    e1.record()
    torch.cuda.synchronize()
    import ctypes as C
    n = lib.pg_profile_end()
    names = (C.c_char_p * n)()
    ms = (C.c_float * n)()
    counts = (C.c_int32 * n)()
    lib.pg_profile_read(names, ms, counts, n)
    d = {names[i].decode(): (ms[i], counts[i]) for i in range(n)}
    gemm = [v for k, v in d.items() if k.startswith("gemm")]
    return {"launches": sum(v[1] for v in d.values()), "families": {k: v[1] for k, v in d.items()},
            "gemm_ms": sum(v[0] for v in gemm), "gemm_launches": sum(v[1] for v in gemm),
            "attention_ms": d.get("attention", (0.0, 0))[0], "layernorm_ms": d.get("layernorm", (0.0, 0))[0],
            "other_ms": sum(v[0] for k, v in d.items() if not k.startswith("gemm") and k not in ("attention", "layernorm")),
            "total_ms": e0.elapsed_time(e1)}


def cpu_port_images(n_images: int, threads: int):
    """The oracle (CPU port of the reference path) on `n_images` four-view images; returns seconds."""
    from oracle import head as ohead, refiner as oref, vit as ovit
    from pigeon_b200 import synthetic
    from pigeon_b200.vit_engine import VitDims
    torch.set_num_threads(threads)
    dims = VitDims()
    sd = synthetic.random_vit_state_dict(dims, seed=0)
    cells = torch.from_numpy(synthetic.synthetic_geocells(NUM_CELLS, 0))
    lin = torch.nn.Linear(dims.hidden, NUM_CELLS)
    bank = synthetic.synthetic_bank(NUM_CELLS, 20_000, dims.hidden, seed=2, empty_cells=5)
    px = torch.randn(n_images * 4, 3, dims.image_size, dims.image_size, generator=torch.Generator().manual_seed(1))
    t0 = time.perf_counter()
    emb = ovit.clip_embedding(sd, px, patch=dims.patch_size, heads=dims.heads, layers=dims.layers, eps=dims.ln_eps)
    h = ohead.head_forward(emb.reshape(n_images, 4, -1), lin.weight.detach(), lin.bias.detach(), cells, NUM_CAND, True)
    oref.refiner_forward(bank, emb.reshape(n_images, 4, -1), h.pred_LLH, h.topk_idx, h.topk_val, TOPK, REFINER_T, REFINER_MAX_KM)
    return time.perf_counter() - t0


def usable_cpus() -> int:
    """Host cores this process may really use: affinity mask capped by the cgroup CPU quota."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = max(1, min(n, int(float(q) / float(p) + 0.5)))
    except (OSError, ValueError):
        pass
    return n


_BEST_THREADS = None


def best_threads() -> int:
    """Thread count that maximises fp32 GEMM throughput of the oracle's dominant op on this host (a box that
    advertises more logical CPUs than it schedules runs slower with all of them); tried: powers of two up to the
    usable core count."""
    global _BEST_THREADS
    if _BEST_THREADS is None:
        n = usable_cpus()
        cands = sorted({c for c in (4, 8, 16, 32, 64, 128, 256, n) if c <= n} | {n})
        a, b = torch.randn(2308, 1024), torch.randn(1024, 4096)
        best, best_t = n, float("inf")
        for c in cands:
            torch.set_num_threads(c)
            a @ b
            t0 = time.perf_counter()
            for _ in range(3):
                a @ b
            dt = time.perf_counter() - t0
            if dt < best_t * 0.95:
                best, best_t = c, dt
        _BEST_THREADS = best
    return _BEST_THREADS


def cpu_baseline(sample_images: int):
    threads = best_threads()
    cpu_port_images(1, threads)                       # warm-up (thread pools, allocator)
    sec = cpu_port_images(sample_images, threads)
    return {"value": sample_images / sec, "unit": "images/s", "cores": threads, "kind": "port",
            "sample": f"{sample_images} four-view images ({4 * sample_images} views) through the fp32 CPU oracle "
                      f"(ViT-L/14-336 + head + refiner), {sec:.1f} s, torch {torch.__version__} with {threads} threads "
                      f"(best of the thread counts tried; {usable_cpus()} usable logical CPUs)"}


def run_reference(args):
    """CPU arm: the oracle port of the reference path (the reference itself is Python under /root/reference and
    cannot travel to the GPU box).  One step = ONE four-view image (bounded sample of the same workload)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = best_threads()
    for _ in range(min(args.warmup, 1)):
        cpu_port_images(1, threads)
    steps = max(1, min(args.steps, 8))
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_port_images(1, threads)
    sec = time.perf_counter() - t0
    v = steps / sec
    sample = (f"{steps} steps x 1 four-view image (4 views) through the fp32 CPU oracle, {threads} threads "
              f"(best of the thread counts tried; {usable_cpus()} usable logical CPUs)")
    print(json.dumps({
        "impl": "reference", "metric": "four-view 336x336 images/sec (ViT-L/14 + geocell head + ProtoRefiner)",
        "value": v, "unit": "images/s", "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1),
        "ms_per_step": sec / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp32", "data": "synthetic",
        "config": {"workload": "1 four-view synthetic panorama per step (bounded sample of BASELINE.json configs[1]), "
                               "ViT-L/14-336 + head + ProtoRefiner on host cores"},
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="four-view samples per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=4, help="four-view images in the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ln-fold", action="store_true", help="A/B: run the LayerNorm kernels instead of the folded epilogues")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
