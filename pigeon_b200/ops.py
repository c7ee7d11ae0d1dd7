"""torch.Tensor-level wrappers over the C ABI building blocks (tests, and callers that fuse differently).

Every function requires CUDA tensors and raises `PigeonB200Error` otherwise — there is no eager fallback.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import PigeonB200Error, check, current_stream_ptr, load, ptr


def _need_cuda(*ts: torch.Tensor) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise PigeonB200Error("pigeon_b200 kernels need CUDA tensors (there is no CPU path)")
        if t is not None and not t.is_contiguous():
            raise PigeonB200Error("pigeon_b200 kernels need contiguous tensors")


def gemm_f16(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None, epilogue: int,
             out: torch.Tensor | None = None) -> torch.Tensor:
    """out = epilogue(a @ w.T + bias); a [M,K] fp16, w [N,K] fp16 (nn.Linear layout), fp32 accumulate."""
    _need_cuda(a, w, bias, out)
    assert a.dtype == torch.float16 and w.dtype == torch.float16 and a.shape[1] == w.shape[1]
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        assert epilogue != _lib.EPI_F32_BIAS_RESID, "residual epilogue accumulates into `out`"
        dt = torch.float16 if epilogue in (_lib.EPI_F16_BIAS, _lib.EPI_F16_BIAS_QGELU) else torch.float32
        out = torch.empty((M, N), dtype=dt, device=a.device)
    check(load().pg_gemm_f16(ptr(a), a.stride(0), ptr(w), w.stride(0), ptr(out), out.stride(0), ptr(bias), M, N, K,
                             epilogue, current_stream_ptr()), "pg_gemm_f16")
    return out


def layernorm_f16(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float) -> torch.Tensor:
    _need_cuda(x, gamma, beta)
    assert x.dtype == torch.float32
    rows, hidden = x.reshape(-1, x.shape[-1]).shape
    y = torch.empty(x.shape, dtype=torch.float16, device=x.device)
    check(load().pg_layernorm_f16(ptr(x), ptr(y), ptr(gamma), ptr(beta), rows, hidden, eps, current_stream_ptr()),
          "pg_layernorm_f16")
    return y


def attention_f16(qkv: torch.Tensor, n_views: int, seq: int, heads: int, variant: int | None = None,
                  poly: int = -1, return_lse2: bool = False):
    """qkv fp16 [n_views*seq, 3*heads*64] -> fp16 [n_views*seq, heads*64].

    `variant` (A/B measurements only): 0 = pair kernel, 1 = first-generation kernel, 2 = split kernel, 3 = fold kernel;
    None = the library default.  `return_lse2` (with an explicit variant): also the log2-domain log-sum-exp of the scaled
    logits, f32 [n_views*heads, seq], the side output the training forward keeps for the backward pass."""
    _need_cuda(qkv)
    assert qkv.dtype == torch.float16 and qkv.shape == (n_views * seq, 3 * heads * 64)
    out = torch.empty((n_views * seq, heads * 64), dtype=torch.float16, device=qkv.device)
    lse2 = torch.empty((n_views * heads, seq), dtype=torch.float32, device=qkv.device) if return_lse2 else None
    if variant is None:
        if return_lse2:
            raise ValueError("return_lse2 needs an explicit variant (pg_attention_f16 has no lse2 argument)")
        check(load().pg_attention_f16(ptr(qkv), ptr(out), n_views, seq, heads, current_stream_ptr()), "pg_attention_f16")
    else:
        check(load().pg_attention_f16_variant(ptr(qkv), ptr(out), ptr(lse2) if return_lse2 else None, n_views, seq, heads,
                                              variant, poly, current_stream_ptr()), "pg_attention_f16_variant")
    return (out, lse2) if return_lse2 else out


def refiner_scan(bank: "DeviceBank", emb: torch.Tensor, cand_idx: torch.Tensor, topk: int):
    """Scan stage only (cell-sharded banks): per (query, candidate) partials (best_logit [B, topk], best_lnglat
    [B, topk, 2], best_proto [B, topk]); pairs whose geocell this bank does not hold come back as -100000 / (0, 0) / -1."""
    import ctypes as C
    if emb.dim() == 2:
        emb = emb.unsqueeze(1)
    emb = emb.to(torch.float32).contiguous()
    cand_idx = cand_idx.to(torch.int64).contiguous()
    _need_cuda(emb, cand_idx)
    B, V, D = emb.shape
    if D != bank.dim:
        raise PigeonB200Error(f"embedding dim {D} != bank dim {bank.dim}")
    dev = emb.device
    lib = load()
    ws = torch.empty(lib.pg_refiner_workspace_bytes(B, topk, D, bank.num_cells), dtype=torch.uint8, device=dev)
    bl = torch.empty((B, topk), dtype=torch.float32, device=dev)
    bll = torch.empty((B, topk, 2), dtype=torch.float32, device=dev)
    bp = torch.empty((B, topk), dtype=torch.int32, device=dev)
    check(lib.pg_refiner_scan(C.byref(bank.c_struct), ptr(emb), B, V, ptr(cand_idx), cand_idx.shape[1], topk, ptr(ws),
                              ws.numel(), ptr(bl), ptr(bll), ptr(bp), current_stream_ptr()), "pg_refiner_scan")
    return bl, bll, bp


def refiner_finalize(best_logit: torch.Tensor, best_lnglat: torch.Tensor, init_lnglat: torch.Tensor,
                     cand_idx: torch.Tensor, cand_prob: torch.Tensor, topk: int, temperature: float,
                     max_refinement_km: float):
    """Final stage on merged partials -> (preds_LLH f32 [B, 2], preds_geocell i64 [B], choice i32 [B])."""
    best_logit = best_logit.to(torch.float32).contiguous()
    best_lnglat = best_lnglat.to(torch.float32).contiguous()
    init_lnglat = init_lnglat.to(torch.float64).contiguous()
    cand_idx = cand_idx.to(torch.int64).contiguous()
    cand_prob = cand_prob.to(torch.float32).contiguous()
    _need_cuda(best_logit, best_lnglat, init_lnglat, cand_idx, cand_prob)
    B = best_logit.shape[0]
    dev = best_logit.device
    out_ll = torch.empty((B, 2), dtype=torch.float32, device=dev)
    out_cell = torch.empty((B,), dtype=torch.int64, device=dev)
    choice = torch.empty((B,), dtype=torch.int32, device=dev)
    check(load().pg_refiner_finalize(ptr(best_logit), ptr(best_lnglat), ptr(init_lnglat), ptr(cand_idx), ptr(cand_prob),
                                     cand_idx.shape[1], B, topk, float(temperature), float(max_refinement_km),
                                     ptr(out_ll), ptr(out_cell), ptr(choice), current_stream_ptr()),
          "pg_refiner_finalize")
    return out_ll, out_cell, choice


def head_set_fused(on: bool) -> None:
    """A/B switch of the geocell head: False (default) three kernels, True the single fused kernel (same results)."""
    check(load().pg_head_set_fused(1 if on else 0), "pg_head_set_fused")


def refiner_set_schedule(mode: int) -> None:
    """A/B switch of the refiner scan: 0 automatic, 1 query-major, 2 cell-major, 3 tile scan, 4 slab scan (same selections)."""
    check(load().pg_refiner_set_schedule(int(mode)), "pg_refiner_set_schedule")


def head_pack_weight(weight: torch.Tensor) -> torch.Tensor:
    """cell_layer.weight f32 [C, D] -> fp16 [C, 3*D] error-compensated split consumed by `head_forward`."""
    _need_cuda(weight)
    w = weight.detach().to(torch.float32).contiguous()
    C_, D = w.shape
    w3 = torch.empty((C_, 3 * D), dtype=torch.float16, device=w.device)
    check(load().pg_head_pack_weight(ptr(w), ptr(w3), C_, D, current_stream_ptr()), "pg_head_pack_weight")
    return w3


def head_forward(emb: torch.Tensor, w3: torch.Tensor, bias: torch.Tensor, centroids: torch.Tensor, k: int):
    """emb f32 [B, V, D] -> dict(pooled, logits, probs, pred_cell, pred_lnglat, topk_val, topk_idx)."""
    _need_cuda(emb, w3, bias, centroids)
    assert emb.dtype == torch.float32 and emb.dim() == 3
    assert centroids.dtype == torch.float64 and bias.dtype == torch.float32
    B, V, D = emb.shape
    C_ = w3.shape[0]
    dev = emb.device
    lib = load()
    ws = torch.empty(lib.pg_head_workspace_bytes(B, D), dtype=torch.uint8, device=dev)
    out = dict(
        pooled=torch.empty((B, D), dtype=torch.float32, device=dev),
        logits=torch.empty((B, C_), dtype=torch.float32, device=dev),
        probs=torch.empty((B, C_), dtype=torch.float32, device=dev),
        pred_cell=torch.empty((B,), dtype=torch.int64, device=dev),
        pred_lnglat=torch.empty((B, 2), dtype=torch.float64, device=dev),
        topk_val=torch.empty((B, k), dtype=torch.float32, device=dev),
        topk_idx=torch.empty((B, k), dtype=torch.int64, device=dev),
    )
    check(lib.pg_head_forward(ptr(emb), B, V, D, ptr(w3), ptr(bias), ptr(centroids), C_, k, ptr(ws), ws.numel(),
                              ptr(out["pooled"]), ptr(out["logits"]), ptr(out["probs"]), ptr(out["pred_cell"]),
                              ptr(out["pred_lnglat"]), ptr(out["topk_val"]), ptr(out["topk_idx"]),
                              current_stream_ptr()), "pg_head_forward")
    return out


class DeviceBank:
    """CSR prototype bank resident in HBM (layout of `pg_refiner_bank`, include/pigeon_b200.h)."""

    FIELDS = ("cell_off", "proto_emb", "proto_lnglat", "proto_count", "member_off", "member_idx", "data_emb",
              "data_lnglat")
    DTYPES = dict(cell_off=torch.int64, proto_emb=torch.float32, proto_lnglat=torch.float32, proto_count=torch.int32,
                  member_off=torch.int64, member_idx=torch.int64, data_emb=torch.float32, data_lnglat=torch.float32)

    def __init__(self, device, **arrays):
        self.device = torch.device(device)
        for f in self.FIELDS:
            t = torch.as_tensor(arrays[f]).to(device=self.device, dtype=self.DTYPES[f]).contiguous()
            if t.numel() == 0:  # keep a valid (never dereferenced) pointer for empty arrays
                t = torch.zeros((1,) + tuple(t.shape[1:]), dtype=t.dtype, device=self.device)
            setattr(self, f, t)
        self.num_cells = int(arrays["cell_off"].shape[0]) - 1
        self.dim = int(self.proto_emb.shape[1])
        # |p|^2 per prototype, once per bank: the tile scan (schedule 3) scores in the |p|^2 + |q|^2 - 2 p.q form
        self.proto_sqnorm = None
        P = int(arrays["proto_emb"].shape[0])
        if self.device.type == "cuda" and P > 0:
            self.proto_sqnorm = torch.empty(P, dtype=torch.float32, device=self.device)
            with torch.cuda.device(self.device):
                check(load().pg_refiner_bank_sqnorm(ptr(self.proto_emb), P, self.dim, ptr(self.proto_sqnorm),
                                                    current_stream_ptr()), "pg_refiner_bank_sqnorm")
        self.c_struct = _lib.RefinerBank(self.num_cells, self.dim, *(ptr(getattr(self, f)) for f in self.FIELDS),
                                         ptr(self.proto_sqnorm) if self.proto_sqnorm is not None else None, P,
                                         int((self.cell_off[1:] > self.cell_off[:-1]).sum().item()))


def refiner_forward(bank: DeviceBank, emb: torch.Tensor, init_lnglat: torch.Tensor, cand_idx: torch.Tensor,
                    cand_prob: torch.Tensor, topk: int, temperature: float, max_refinement_km: float,
                    debug: bool = False):
    """ProtoRefiner.forward on the GPU. emb f32 [B, D] or [B, V, D]; returns (preds_LLH f32 [B,2], preds_geocell i64 [B])
    and, with `debug`, a dict of the per-candidate intermediates."""
    import ctypes as C
    if emb.dim() == 2:
        emb = emb.unsqueeze(1)
    emb = emb.to(torch.float32).contiguous()
    init_lnglat = init_lnglat.to(torch.float64).contiguous()
    cand_idx = cand_idx.to(torch.int64).contiguous()
    cand_prob = cand_prob.to(torch.float32).contiguous()
    _need_cuda(emb, init_lnglat, cand_idx, cand_prob)
    B, V, D = emb.shape
    if D != bank.dim:
        raise PigeonB200Error(f"embedding dim {D} != bank dim {bank.dim}")
    stride = cand_idx.shape[1]
    dev = emb.device
    lib = load()
    ws = torch.empty(lib.pg_refiner_workspace_bytes(B, topk, D, bank.num_cells), dtype=torch.uint8, device=dev)
    out_ll = torch.empty((B, 2), dtype=torch.float32, device=dev)
    out_cell = torch.empty((B,), dtype=torch.int64, device=dev)
    dbg = None
    if debug:
        dbg = dict(best_logit=torch.empty((B, topk), dtype=torch.float32, device=dev),
                   best_lnglat=torch.empty((B, topk, 2), dtype=torch.float32, device=dev),
                   best_proto=torch.empty((B, topk), dtype=torch.int32, device=dev),
                   choice=torch.empty((B,), dtype=torch.int32, device=dev))
    check(lib.pg_refiner_forward(C.byref(bank.c_struct), ptr(emb), B, V, ptr(init_lnglat), ptr(cand_idx),
                                 ptr(cand_prob), stride, topk, float(temperature), float(max_refinement_km), ptr(ws),
                                 ws.numel(), ptr(out_ll), ptr(out_cell),
                                 ptr(dbg["best_logit"]) if dbg else None, ptr(dbg["best_lnglat"]) if dbg else None,
                                 ptr(dbg["best_proto"]) if dbg else None, ptr(dbg["choice"]) if dbg else None,
                                 current_stream_ptr()), "pg_refiner_forward")
    return (out_ll, out_cell, dbg) if debug else (out_ll, out_cell)
