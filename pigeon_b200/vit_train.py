"""Fine-tune step of the vision tower (host side of pg_vit_forward_train / pg_vit_backward).

What the reference gets from autograd + DDP for `accelerator.backward(output.loss)` (training/train_eval_loop.py:216)
with the freeze policy of models/super_guessr.py:159-160 (embeddings, pre_layrnorm and the LAST encoder layer
trainable; `requires_grad` decides here too): the training-mode forward keeps its activations in one arena, the
backward runs as hand-written kernels (bf16 operands, fp32 accumulation) and leaves fp32 gradients in `param.grad`
under the HuggingFace parameter names, ready for `pigeon_b200.training.AdamW`.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch

from . import _lib
from ._lib import PigeonB200Error, check, current_stream_ptr, load, ptr


def _align(n: int, a: int = 1024) -> int:
    return (n + a - 1) // a * a


def _aligned_empty(nbytes: int, device) -> torch.Tensor:
    """uint8 buffer whose data pointer is 1024-byte aligned (the caching allocator only guarantees 512)."""
    raw = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
    off = (-raw.data_ptr()) % 1024
    return raw[off:off + nbytes]


class TowerTrainer:
    """Owns the activation arena, the transposed bf16 weights and the fused gradient buffers of one CLIPVisionTower."""

    def __init__(self, tower, max_views: int = 64):
        self.tower = tower
        self.dims = tower.dims
        self.max_views = int(max_views)
        self._lib = load()
        self._arena: Optional[torch.Tensor] = None
        self._arena_views = 0
        self._ws: Optional[torch.Tensor] = None
        self._saved = None
        self._saved_layers = None
        self._saved_views = 0
        self._wt: Optional[List[Dict[str, torch.Tensor]]] = None
        self._wt_version = None
        self._pending: Optional[Dict[str, torch.Tensor]] = None
        self._keep = []

    # ---------------------------------------------------------------------------------- which parameters train
    def layout(self):
        vm = self.tower.vision_model
        emb_params = [vm.embeddings.class_embedding, vm.embeddings.patch_embedding.weight,
                      vm.embeddings.position_embedding.weight, vm.pre_layrnorm.weight, vm.pre_layrnorm.bias]
        flags = [p.requires_grad for p in emb_params]
        if any(flags) and not all(flags):
            raise PigeonB200Error("embeddings and pre_layrnorm must be trainable or frozen together")
        layers = []
        for i, L in enumerate(vm.encoder.layers):
            f = [p.requires_grad for p in L.parameters()]
            if any(f) and not all(f):
                raise PigeonB200Error(f"encoder layer {i}: parameters must be trainable or frozen together")
            layers.append(all(f))
        return all(flags), layers

    def any_trainable(self) -> bool:
        e, layers = self.layout()
        return e or any(layers)

    # ---------------------------------------------------------------------------------- buffers
    def _build_saved(self, n_views: int):
        d = self.dims
        rows = n_views * d.tokens
        H, I = d.hidden, d.intermediate
        per_layer = [("x0", rows * H * 4), ("xn1", rows * H * 2), ("qkv", rows * 3 * H * 2),
                     ("lse2", n_views * d.heads * d.tokens * 4), ("ao", rows * H * 2), ("x1", rows * H * 4),
                     ("xn2", rows * H * 2), ("u", rows * I * 2), ("h", rows * I * 2)]
        top = [("im2col", n_views * (d.tokens - 1) * d.patch_k_pad * 2), ("e", rows * H * 4), ("x_out", rows * H * 4)]
        total = sum(_align(b) for _, b in top) + d.layers * sum(_align(b) for _, b in per_layer)
        dev = next(self.tower.parameters()).device
        if self._arena is None or self._arena.numel() < total:
            self._arena = None
            self._arena = _aligned_empty(total, dev)
        base, off = self._arena.data_ptr(), 0
        if base % 1024:
            raise PigeonB200Error("activation arena is not 1024-byte aligned")
        saved = _lib.VitSaved()
        for name, b in top:
            setattr(saved, name, base + off)
            off += _align(b)
        layers = (_lib.VitSavedLayer * d.layers)()
        for l in range(d.layers):
            for name, b in per_layer:
                setattr(layers[l], name, base + off)
                off += _align(b)
        saved.layers_host = layers
        self._saved, self._saved_layers, self._saved_views = saved, layers, n_views

    def _transposed_weights(self):
        """bf16 transposes of the four weight matrices of every layer (B operands of the backward GEMMs); only the layers whose
        parameters moved since the last call are redone (the last one, under the reference's fine-tune policy)."""
        versions = self.tower._group_versions()
        if self._wt is None:
            self._wt, self._wt_version = [None] * self.dims.layers, {}
        bf = torch.bfloat16
        with torch.no_grad():
            for l, L in enumerate(self.tower.vision_model.encoder.layers):
                if self._wt[l] is not None and self._wt_version.get(l) == versions[l]:
                    continue
                sa, mlp = L.self_attn, L.mlp
                wqkv = torch.cat([sa.q_proj.weight, sa.k_proj.weight, sa.v_proj.weight], dim=0)      # [3H, H]
                self._wt[l] = dict(w_qkv_t=wqkv.t().contiguous().to(bf), w_o_t=sa.out_proj.weight.t().contiguous().to(bf),
                                   w_fc1_t=mlp.fc1.weight.t().contiguous().to(bf),
                                   w_fc2_t=mlp.fc2.weight.t().contiguous().to(bf))
                self._wt_version[l] = versions[l]
        return self._wt

    def _pending_grads(self):
        if self._pending is not None:
            return self._pending
        d = self.dims
        dev = next(self.tower.parameters()).device
        emb_train, layers = self.layout()
        shapes = []
        if emb_train:
            shapes += [("patch_w", (d.hidden, d.patch_k_pad)), ("class_emb", (d.hidden,)), ("pos_emb", (d.tokens, d.hidden)),
                       ("pre_ln_g", (d.hidden,)), ("pre_ln_b", (d.hidden,))]
        H, I = d.hidden, d.intermediate
        for l, t in enumerate(layers):
            if t:
                shapes += [(f"{l}.ln1_g", (H,)), (f"{l}.ln1_b", (H,)), (f"{l}.w_qkv", (3 * H, H)), (f"{l}.b_qkv", (3 * H,)),
                           (f"{l}.w_o", (H, H)), (f"{l}.b_o", (H,)), (f"{l}.ln2_g", (H,)), (f"{l}.ln2_b", (H,)),
                           (f"{l}.w_fc1", (I, H)), (f"{l}.b_fc1", (I,)), (f"{l}.w_fc2", (H, I)), (f"{l}.b_fc2", (H,))]
        # ONE flat fp32 buffer (every tensor 16-byte aligned inside it): a single all-reduce averages the whole tower
        offs, total = [], 0
        for _, shp in shapes:
            offs.append(total)
            total += (int(torch.Size(shp).numel()) + 3) // 4 * 4
        self._pending_flat = torch.zeros(total, dtype=torch.float32, device=dev)
        g: Dict[str, torch.Tensor] = {}
        for (name, shp), o in zip(shapes, offs):
            g[name] = self._pending_flat[o:o + int(torch.Size(shp).numel())].view(shp)
        self._pending = g
        return g

    # ---------------------------------------------------------------------------------- forward / backward
    @torch.no_grad()
    def forward(self, pixel_values: torch.Tensor) -> torch.Tensor:
        """pixel_values CUDA [n, 3, S, S] (fp16 or fp32), n <= max_views -> token-mean embedding [n, hidden] f32;
        the activations stay in the arena until the next forward."""
        d = self.dims
        n = pixel_values.shape[0]
        if n > self.max_views:
            raise PigeonB200Error(f"{n} views exceed max_views={self.max_views} of the training arena")
        if not pixel_values.is_cuda or tuple(pixel_values.shape[1:]) != (3, d.image_size, d.image_size):
            raise ValueError(f"pixel_values must be CUDA [n, 3, {d.image_size}, {d.image_size}]")
        if pixel_values.dtype not in (torch.float16, torch.float32):
            pixel_values = pixel_values.float()
        pixel_values = pixel_values.contiguous()
        self._build_saved(n)
        eng = self.tower.engine()
        emb = torch.empty((n, d.hidden), dtype=torch.float32, device=pixel_values.device)
        check(self._lib.pg_vit_forward_train(eng._handle, ptr(pixel_values), int(pixel_values.dtype == torch.float16), n,
                                             C.byref(self._saved), ptr(emb), current_stream_ptr()), "pg_vit_forward_train")
        return emb

    def _buckets(self):
        """Gradient buckets of the flat pending buffer in the order the backward completes them: encoder layers from the
        top down, then the embeddings.  (event slot, first element, one past the last element)"""
        g, d = self._pending_grads(), self.dims
        base = self._pending_flat.data_ptr()
        off = lambda t: (t.data_ptr() - base) // 4
        out = []
        emb_train, layers = self.layout()
        for l in reversed(range(d.layers)):
            if layers[l]:
                out.append((l, off(g[f"{l}.ln1_g"]), off(g[f"{l}.b_fc2"]) + g[f"{l}.b_fc2"].numel()))
        if emb_train:
            out.append((d.layers, off(g["patch_w"]), off(g["pre_ln_b"]) + g["pre_ln_b"].numel()))
        return out

    @torch.no_grad()
    def backward(self, d_emb: torch.Tensor, overlap_world: int = 1) -> None:
        """d_emb f32 [n, hidden] for the views of the last forward; gradients accumulate in the pending buffers.

        overlap_world > 1 (the LAST chunk of a micro-batch under torch.distributed): each layer's gradient bucket is summed
        over the ranks by an NCCL all-reduce on a side stream as soon as pg_vit_backward reports the layer done, while the
        layers below are still in their backward; `finalize` then only waits for the outstanding buckets."""
        d = self.dims
        n = self._saved_views
        if self._saved is None or d_emb.shape != (n, d.hidden):
            raise PigeonB200Error("backward() must follow forward() with one gradient row per view")
        d_emb = d_emb.to(torch.float32).contiguous()
        eng = self.tower.engine()
        wt = self._transposed_weights()
        g = self._pending_grads()
        emb_train, layers = self.layout()
        lb = (_lib.VitLayerBwd * d.layers)()
        for l in range(d.layers):
            for k, t in wt[l].items():
                setattr(lb[l], k, t.data_ptr())
            if layers[l]:
                for k in ("ln1_g", "ln1_b", "w_qkv", "b_qkv", "w_o", "b_o", "ln2_g", "ln2_b", "w_fc1", "b_fc1", "w_fc2", "b_fc2"):
                    setattr(lb[l], "d_" + k, g[f"{l}.{k}"].data_ptr())
        grads = _lib.VitGrads()
        if emb_train:
            grads.d_patch_w, grads.d_class_emb, grads.d_pos_emb = g["patch_w"].data_ptr(), g["class_emb"].data_ptr(), g["pos_emb"].data_ptr()
            grads.d_pre_ln_g, grads.d_pre_ln_b = g["pre_ln_g"].data_ptr(), g["pre_ln_b"].data_ptr()
        grads.layers_host = lb
        events, ev_arr = None, None
        if overlap_world > 1:
            if getattr(self, "_bucket_events", None) is None:
                self._bucket_events = [torch.cuda.Event() for _ in range(d.layers + 1)]
                for e in self._bucket_events:
                    e.record()                                    # materialises the cudaEvent_t handle
                self._side = torch.cuda.Stream(device=d_emb.device)
            events = self._bucket_events
            ev_arr = (C.c_void_p * (d.layers + 1))(*[e.cuda_event for e in events])
            grads.layer_done_events = ev_arr
        need = int(self._lib.pg_vit_backward_workspace_bytes(eng._handle, n))
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = _aligned_empty(need, d_emb.device)
        check(self._lib.pg_vit_backward(eng._handle, C.byref(self._saved), ptr(d_emb), n, C.byref(grads), ptr(self._ws),
                                        self._ws.numel(), current_stream_ptr()), "pg_vit_backward")
        if overlap_world > 1:
            self._works = []
            with torch.cuda.stream(self._side):
                for slot, lo, hi in self._buckets():
                    self._side.wait_event(events[slot])
                    self._works.append(torch.distributed.all_reduce(self._pending_flat[lo:hi], async_op=True))

    @torch.no_grad()
    def finalize(self, world: int = 1) -> None:
        """Pending fused gradients -> `param.grad` (accumulating), averaged over `world` ranks after a summing all-reduce."""
        if self._pending is None:
            return
        g, d = self._pending, self.dims
        if world > 1:                                  # DDP: average of the per-rank gradients (train_eval_loop.py:192)
            works, self._works = getattr(self, "_works", None), None
            if works:                                  # buckets already in flight behind the last chunk's backward
                for wk in works:
                    wk.wait()                          # the current stream waits for the NCCL stream
            else:
                torch.distributed.all_reduce(self._pending_flat)
            self._pending_flat.mul_(1.0 / world)
        vm = self.tower.vision_model

        def put(p, t):
            t = t.contiguous()
            p.grad = t.clone() if p.grad is None else p.grad.add_(t)

        if "patch_w" in g:
            e = vm.embeddings
            put(e.patch_embedding.weight, g["patch_w"][:, : d.patch_k].reshape(e.patch_embedding.weight.shape))
            put(e.class_embedding, g["class_emb"])
            put(e.position_embedding.weight, g["pos_emb"])
            put(vm.pre_layrnorm.weight, g["pre_ln_g"])
            put(vm.pre_layrnorm.bias, g["pre_ln_b"])
        H = d.hidden
        for l, L in enumerate(vm.encoder.layers):
            if f"{l}.w_qkv" not in g:
                continue
            sa, mlp = L.self_attn, L.mlp
            for i, proj in enumerate((sa.q_proj, sa.k_proj, sa.v_proj)):
                put(proj.weight, g[f"{l}.w_qkv"][i * H:(i + 1) * H])
                put(proj.bias, g[f"{l}.b_qkv"][i * H:(i + 1) * H])
            put(sa.out_proj.weight, g[f"{l}.w_o"]); put(sa.out_proj.bias, g[f"{l}.b_o"])
            put(L.layer_norm1.weight, g[f"{l}.ln1_g"]); put(L.layer_norm1.bias, g[f"{l}.ln1_b"])
            put(L.layer_norm2.weight, g[f"{l}.ln2_g"]); put(L.layer_norm2.bias, g[f"{l}.ln2_b"])
            put(mlp.fc1.weight, g[f"{l}.w_fc1"]); put(mlp.fc1.bias, g[f"{l}.b_fc1"])
            put(mlp.fc2.weight, g[f"{l}.w_fc2"]); put(mlp.fc2.bias, g[f"{l}.b_fc2"])
        self._pending = None
