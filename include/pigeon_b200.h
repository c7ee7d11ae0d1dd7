/*
 * pigeon_b200 — C ABI of the B200-native PIGEON inference hot path.
 *
 * The reference (LukasHaas/PIGEON) has no FFI: its hot path sits behind Python nn.Module calls.  This header
 * is the seam a maintainer binds underneath those modules (ctypes stub in INTEGRATION.md); every entry point
 * names the reference code it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in _host; the caller owns every buffer;
 *   - no hidden device allocations: scratch comes from the caller (`*_workspace_bytes` tells how much);
 *   - every call enqueues work on `stream` (a cudaStream_t passed as void*) and returns immediately;
 *   - return value 0 = success, non-zero = failure with a message in pg_last_error() (thread-local);
 *     no exceptions cross the boundary; there is NO CPU fallback anywhere behind this ABI;
 *   - fp16 means IEEE binary16 (`__half`), "f32"/"f64" IEEE float/double, i64 = int64_t.
 */
#ifndef PIGEON_B200_H_
#define PIGEON_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PG_ABI_VERSION 3

/* ---------------------------------------------------------------------------------------------------
 * Library
 * ------------------------------------------------------------------------------------------------- */
int pg_abi_version(void);
/* Message of the last failing call on this thread ("" if none). */
const char* pg_last_error(void);
/* Number of SMs of the current device (grid sizing); < 0 on failure. */
int pg_device_sm_count(void);

/* ---------------------------------------------------------------------------------------------------
 * CLIP ViT vision tower  (HF CLIPVisionTransformer as driven by reference
 * models/clip_embedder.py:58-65 and models/super_guessr.py:386-398)
 * ------------------------------------------------------------------------------------------------- */
typedef struct pg_vit_config {
  int32_t image_size;   /* 336 */
  int32_t patch_size;   /* 14  */
  int32_t hidden;       /* 1024, multiple of 256 */
  int32_t heads;        /* 16, hidden / heads must be 64 */
  int32_t intermediate; /* 4096, multiple of 256 */
  int32_t layers;       /* 24 */
  float ln_eps;         /* 1e-5 */
  int32_t patch_k_pad;  /* padded im2col K: multiple of 64, >= 3*patch*patch (640 for 14x14) */
} pg_vit_config;

/* One encoder block (HF CLIPEncoderLayer).  Matrices are fp16 in nn.Linear layout [out, in]. */
typedef struct pg_vit_layer {
  const float* ln1_g; const float* ln1_b;       /* layer_norm1 [hidden] */
  const void* w_qkv; const float* b_qkv;        /* [3*hidden, hidden] = [Wq; Wk; Wv], bias [3*hidden] */
  const void* w_o; const float* b_o;            /* out_proj [hidden, hidden] */
  const float* ln2_g; const float* ln2_b;       /* layer_norm2 [hidden] */
  const void* w_fc1; const float* b_fc1;        /* mlp.fc1 [intermediate, hidden] */
  const void* w_fc2; const float* b_fc2;        /* mlp.fc2 [hidden, intermediate] */
  /* Optional (all six or none; NULL = pg_vit_forward runs the LayerNorm kernels): LayerNorm folded into the GEMM after it.
   *   w_*_ln   fp16 = W * gamma (column-wise), same shape as W;   cs_* f32 [out] = row sums of the fp16 values of w_*_ln;
   *   b_*_ln   f32 [out] = b + W beta.
   * With them the forward feeds the RAW fp16 residual row to the tensor cores and applies
   * rstd * (acc - mu * cs) + b_ln in the epilogue; (mu, rstd) come from moments the previous epilogue left behind. */
  const void* w_qkv_ln; const float* b_qkv_ln; const float* cs_qkv;
  const void* w_fc1_ln; const float* b_fc1_ln; const float* cs_fc1;
} pg_vit_layer;

typedef struct pg_vit_weights {
  const void* patch_w;      /* fp16 [hidden, patch_k_pad]: Conv2d weight flattened (c, ky, kx), zero padded */
  const float* class_emb;   /* [hidden] */
  const float* pos_emb;     /* [tokens, hidden], tokens = (image/patch)^2 + 1 */
  const float* pre_ln_g; const float* pre_ln_b; /* pre_layrnorm [hidden] */
  const pg_vit_layer* layers_host;              /* HOST array of `layers` entries (device pointers inside) */
} pg_vit_weights;

typedef struct pg_vit pg_vit;

/* Copies the config and the pointer tables (not the weights).  Replaces CLIPVisionModel construction at
 * reference models/clip_embedder.py:26 / evaluation/evaluate.py:36 for the forward path. */
int pg_vit_create(const pg_vit_config* cfg, const pg_vit_weights* w, pg_vit** out);
void pg_vit_destroy(pg_vit* h);
/* Scratch needed by pg_vit_forward for `n_views` images. */
size_t pg_vit_workspace_bytes(const pg_vit* h, int32_t n_views);
/* pixels [n_views, 3, image, image] (fp32 if pixels_f16 == 0 else fp16), NCHW contiguous
 *   -> emb_out f32 [n_views, hidden] = mean over all tokens of last_hidden_state (pre post_layernorm),
 *      i.e. reference models/clip_embedder.py:63-65 / models/super_guessr.py:395-398;
 *   -> hidden_out (optional, may be NULL) f32 [n_views, tokens, hidden] = last_hidden_state. */
int pg_vit_forward(pg_vit* h, const void* pixels, int32_t pixels_f16, int32_t n_views, void* workspace,
                   size_t workspace_bytes, float* emb_out, float* hidden_out, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Geocell head (reference models/super_guessr.py:437-459: view mean, cell_layer, softmax, argmax, top-k)
 * ------------------------------------------------------------------------------------------------- */
/* One-off: cell_layer.weight f32 [C, D] -> packed fp16 [C, 3*D] = [Whi | Whi | Wlo] (error-compensated split). */
int pg_head_pack_weight(const float* w, void* w3_out, int32_t C, int32_t D, void* stream);
size_t pg_head_workspace_bytes(int32_t B, int32_t D);
/* emb f32 [B, V, D] (V = 4 for panoramas, 1 otherwise)
 *   -> pooled f32 [B, D]; logits f32 [B, C]; probs f32 [B, C]; pred_cell i64 [B];
 *      pred_lnglat f64 [B, 2] = centroids[pred_cell] (lng, lat);
 *      topk_val f32 [B, k] (descending), topk_idx i64 [B, k]. */
int pg_head_forward(const float* emb, int32_t B, int32_t V, int32_t D, const void* w3, const float* bias,
                    const double* centroids, int32_t C, int32_t k, void* workspace, size_t workspace_bytes,
                    float* pooled, float* logits, float* probs, int64_t* pred_cell, double* pred_lnglat,
                    float* topk_val, int64_t* topk_idx, void* stream);
/* Measurement / deployment switch of pg_head_forward for this process: 0 (default) = three kernels (view mean + split,
 * tcgen05 GEMM + bias, softmax / arg-max / top-k); 1 = ONE kernel doing all of it (D % 64 == 0 and C <= 6144, otherwise the
 * three-kernel sequence runs): the view mean and the hi/lo split happen in the GEMM's A-operand producer warps and the CTA
 * that completes a block of 128 samples last normalises its rows.  Same results to fp32 summation order. */
int pg_head_set_fused(int32_t on);

/* Classification loss of reference models/super_guessr.py:468-474 (CrossEntropyLoss, mean over the batch).
 *   mode 0: labels_idx i64 [B] class indices;  mode 1: soft f32 [B, C] target probabilities;
 *   mode 2: haversine-smoothed targets from labels_lnglat f64 [B, 2] and centroids f64 [C, 2]
 *           (preprocessing/geo_utils.py:58-74 + preprocessing/utils.py:7-19, smoothing_km = 65).
 * per_sample f64 [B] is scratch/out; loss_out f64 [1].
 * mode 0 labels outside [0, C) are never used as an index: that sample's loss and the batch mean become NaN (and, in
 * pg_head_loss_grad, its gradient row zero).  torch's ignore_index = -100 is not implemented — the reference's labels on this
 * path are geocell indices and never negative. */
int pg_head_loss(const float* logits, int32_t B, int32_t C, int32_t mode, const int64_t* labels_idx,
                 const float* soft, const double* labels_lnglat, const double* centroids, double smoothing_km,
                 double* per_sample, double* loss_out, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Image pre-processing ("next" row N2): what `CLIPProcessor.from_pretrained(CLIP_MODEL)(images=...)` does on the
 * CPU per item in reference dataset_creation/finetune/embed_dataset.py:17-22, preprocessing/dataset_preprocessing.py:
 * 182-204, dataset_creation/benchmark/benchmark_dataset.py:100-104 (transformers 4.23.1 CLIPFeatureExtractor over
 * Pillow): resize shortest edge to `size` (BICUBIC) -> center crop -> /255 -> (x - mean) / std, channel first.
 * Bit-exact with Pillow's fixed-point resampler (uint8 stage) and with numpy float32 (normalisation).
 * ------------------------------------------------------------------------------------------------- */
typedef struct pg_image {
  const uint8_t* data;  /* DEVICE pointer, RGB interleaved (H x W x 3) */
  int32_t height, width;
  int64_t row_stride;   /* bytes between rows, >= 3 * width */
} pg_image;
/* `images` is a HOST array of n descriptors.  0 on error (see pg_last_error). */
size_t pg_preprocess_workspace_bytes(const pg_image* images, int32_t n, int32_t size);
/* out: [n, 3, size, size] f32 (out_f16 = 0) or f16 (out_f16 = 1, round-to-nearest of the f32 value); mean/std: host float[3]. */
int pg_preprocess_clip(const pg_image* images, int32_t n, int32_t size, const float* mean, const float* std,
                       void* workspace, size_t workspace_bytes, void* out, int32_t out_f16, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Fine-tune step, head-only part ("next" row N1 with a frozen or absent base model): what loss.backward() and
 * torch.optim.AdamW.step() do for cell_layer in reference training/train_eval_loop.py:187,215-221.
 * ------------------------------------------------------------------------------------------------- */
/* pg_head_loss plus its gradient: dlogits f32 [B, C] = grad_scale / B * (softmax(logits) * sum_c(target) - target)
 * (the backward of CrossEntropyLoss(reduction='mean') at super_guessr.py:474 for index, soft or smoothed targets). */
int pg_head_loss_grad(const float* logits, int32_t B, int32_t C, int32_t mode, const int64_t* labels_idx,
                      const float* soft, const double* labels_lnglat, const double* centroids, double smoothing_km,
                      double grad_scale, double* per_sample, double* loss_out, float* dlogits, void* stream);
/* Backward of cell_layer (nn.Linear, super_guessr.py:447) in fp32:
 *   dw f32 [C, D] (+)= dlogits^T . pooled;  db f32 [C] (+)= column sums of dlogits;  dpooled f32 [B, D] = dlogits . w.
 * Any of dw / db / dpooled may be NULL; `accumulate` != 0 adds into dw/db (gradient accumulation, :218-221). */
int pg_head_backward(const float* dlogits, const float* pooled, const float* w, int32_t B, int32_t C, int32_t D,
                     int32_t accumulate, float* dw, float* db, float* dpooled, void* stream);
/* torch.optim.AdamW single-tensor update on a flat fp32 parameter (amsgrad = maximize = False), `step` counts from 1;
 * grad is read as grad * grad_scale (1/world_size after a summing all-reduce). */
int pg_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr,
                  double beta1, double beta2, double eps, double weight_decay, int64_t step, double grad_scale,
                  void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Fine-tune step, vision tower ("next" row N1): the forward of pg_vit_forward with every activation the backward
 * needs kept in caller-provided buffers, and the backward itself — what autograd runs through HF
 * CLIPVisionTransformer for `accelerator.backward(output.loss)` in reference training/train_eval_loop.py:216.
 * Gradients flow in bf16 operands with fp32 accumulation; weight gradients accumulate in fp32.
 * rows = n_views * tokens.  All pointers are device pointers; the *_host arrays live on the host.
 * ------------------------------------------------------------------------------------------------- */
typedef struct pg_vit_saved_layer {
  float* x0;    /* f32 [rows, hidden]        residual stream entering the layer */
  void* xn1;    /* f16 [rows, hidden]        layer_norm1 output */
  void* qkv;    /* f16 [rows, 3*hidden]      q | k | v projections */
  float* lse2;  /* f32 [n_views*heads, tokens] log2-sum-exp of the scaled logits */
  void* ao;     /* f16 [rows, hidden]        attention output (input of out_proj) */
  float* x1;    /* f32 [rows, hidden]        residual stream after the attention block */
  void* xn2;    /* f16 [rows, hidden]        layer_norm2 output */
  void* u;      /* f16 [rows, intermediate]  fc1 pre-activation */
  void* h;      /* f16 [rows, intermediate]  quick_gelu(fc1) */
} pg_vit_saved_layer;

typedef struct pg_vit_saved {
  void* im2col;  /* f16 [n_views*patches, patch_k_pad] */
  float* e;      /* f32 [rows, hidden] class/patch + position embeddings (input of pre_layrnorm) */
  float* x_out;  /* f32 [rows, hidden] last_hidden_state */
  const pg_vit_saved_layer* layers_host;  /* [layers] */
} pg_vit_saved;

/* emb_out f32 [n_views, hidden] = token mean of last_hidden_state, as pg_vit_forward. */
int pg_vit_forward_train(pg_vit* h, const void* pixels, int32_t pixels_f16, int32_t n_views, const pg_vit_saved* saved,
                         float* emb_out, void* stream);

typedef struct pg_vit_layer_bwd {
  /* bf16 transposed weights for the data gradients: w_qkv_t [hidden, 3*hidden], w_o_t [hidden, hidden],
   * w_fc1_t [hidden, intermediate], w_fc2_t [intermediate, hidden] (row-major, = W^T of the [out, in] Linear weights) */
  const void *w_qkv_t, *w_o_t, *w_fc1_t, *w_fc2_t;
  /* f32 gradient accumulators (+=), same shapes as the parameters; all NULL = frozen layer
   * (reference freeze policy, models/super_guessr.py:159-160) */
  float *d_ln1_g, *d_ln1_b, *d_w_qkv, *d_b_qkv, *d_w_o, *d_b_o, *d_ln2_g, *d_ln2_b, *d_w_fc1, *d_b_fc1, *d_w_fc2, *d_b_fc2;
} pg_vit_layer_bwd;

typedef struct pg_vit_grads {
  /* embeddings + pre_layrnorm, f32 (+=); all NULL = frozen (the backward then stops at the lowest trainable layer) */
  float *d_patch_w;   /* [hidden, patch_k_pad] */
  float *d_class_emb; /* [hidden] */
  float *d_pos_emb;   /* [tokens, hidden] */
  float *d_pre_ln_g, *d_pre_ln_b;
  const pg_vit_layer_bwd* layers_host;  /* [layers] */
  /* Optional (may be NULL): layers + 1 cudaEvent_t handles.  pg_vit_backward records entry l on its stream as soon as the
   * weight gradients of encoder layer l are complete for this call, entry `layers` after the embedding gradients — so that
   * the caller can start the NCCL all-reduce of a layer's gradient bucket on a side stream while the layers below are
   * still in their backward (reference: DDP's bucketed overlap, training/train_eval_loop.py:192,216).  NULL entries skipped. */
  void* const* layer_done_events;
} pg_vit_grads;

size_t pg_vit_backward_workspace_bytes(const pg_vit* h, int32_t n_views);
/* d_emb f32 [n_views, hidden]: gradient of the loss with respect to the token-mean embedding of each view. */
int pg_vit_backward(pg_vit* h, const pg_vit_saved* saved, const float* d_emb, int32_t n_views, const pg_vit_grads* grads,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * ProtoRefiner (reference models/proto_refiner.py:121-255, 332-357; preprocessing/geo_utils.py:40-55)
 * ------------------------------------------------------------------------------------------------- */
typedef struct pg_refiner_bank {
  int32_t num_cells;           /* C */
  int32_t dim;                 /* D, multiple of 128, <= 1024 */
  const int64_t* cell_off;     /* [C+1] prototype range per geocell; empty range <=> protos[cell] is None */
  const float* proto_emb;      /* [P, D] */
  const float* proto_lnglat;   /* [P, 2] (lng, lat) */
  const int32_t* proto_count;  /* [P] */
  const int64_t* member_off;   /* [P+1] */
  const int64_t* member_idx;   /* [sum count] rows of data_emb/data_lnglat */
  const float* data_emb;       /* [Ntrain, D] */
  const float* data_lnglat;    /* [Ntrain, 2] */
  const float* proto_sqnorm;   /* [P] squared L2 norm of every prototype row (pg_refiner_bank_sqnorm), or NULL: the scans
                                  that need it (schedules 3, 4) are then replaced by the ones that do not */
  int64_t num_protos;          /* P = rows of proto_emb (bounds the TMA tensor map of schedule 4), or 0 */
  int32_t live_cells;          /* geocells with at least one prototype (a cell-sharded bank holds 1/N of them), or 0 = unknown:
                                  the automatic schedule sizes its work units by num_protos / live_cells */
} pg_refiner_bank;

/* |p|^2 of every row of proto_emb f32 [P, D] -> sqnorm_out f32 [P]; once per bank (the squared Euclidean distance of
 * reference models/proto_refiner.py:176-178 in the |p|^2 + |q|^2 - 2 p.q form torch.cdist itself uses). */
int pg_refiner_bank_sqnorm(const float* proto_emb, int64_t P, int32_t D, float* sqnorm_out, void* stream);

size_t pg_refiner_workspace_bytes(int64_t B, int32_t topk, int32_t D, int32_t num_cells);
/* The two stages of pg_refiner_forward as separate calls, for a CELL-SHARDED bank across GPUs (rank r holds the geocells with
 * cell % world == r, every other cell range empty): pg_refiner_scan on every rank over ALL queries -> per (query, candidate)
 * partials best_logit f32 [B, topk] (-100000 where this rank does not hold the cell), best_lnglat f32 [B, topk, 2],
 * best_proto i32 [B, topk]; the caller merges the ranks' partials by owner (one small all-gather), then pg_refiner_finalize
 * (temperature softmax x candidate probabilities, haversine gate, arg-max: reference models/proto_refiner.py:187-231). */
int pg_refiner_scan(const pg_refiner_bank* bank, const float* emb, int64_t B, int32_t V, const int64_t* cand_idx,
                    int32_t cand_stride, int32_t topk, void* workspace, size_t workspace_bytes, float* best_logit,
                    float* best_lnglat, int32_t* best_proto, void* stream);
int pg_refiner_finalize(const float* best_logit, const float* best_lnglat, const double* init_lnglat,
                        const int64_t* cand_idx, const float* cand_prob, int32_t cand_stride, int64_t B, int32_t topk,
                        float temperature, double max_refinement_km, float* out_lnglat, int64_t* out_cell, int32_t* choice,
                        void* stream);
/* Measurement aid: scan schedule of pg_refiner_forward for this process: 0 = automatic (cell-major when geocells are shared
 * by >= 2 (query, candidate) pairs on average; the slab scan when the bank carries proto_sqnorm / num_protos and its geocells
 * hold >= 256 prototypes on average over the non-empty ones), 1 = query-major, 2 = cell-major, 3 = tile scan, 4 = slab scan.  Same selections. */
int pg_refiner_set_schedule(int32_t mode);
/* emb f32 [B, V, D]; init_lnglat f64 [B, 2]; cand_idx i64 [B, cand_stride]; cand_prob f32 [B, cand_stride];
 * only the first `topk` candidates of each row are used (topk <= cand_stride).
 *   -> out_lnglat f32 [B, 2], out_cell i64 [B]   (ProtoRefiner.forward's preds_LLH, preds_geocell)
 *   -> optional debug outputs (may be NULL): best_logit f32 [B, topk] (negative Euclidean distance to the
 *      nearest prototype, -100000 for an empty cell), best_lnglat f32 [B, topk, 2], best_proto i32 [B, topk],
 *      choice i32 [B] (index into the candidate list that was emitted). */
int pg_refiner_forward(const pg_refiner_bank* bank, const float* emb, int64_t B, int32_t V,
                       const double* init_lnglat, const int64_t* cand_idx, const float* cand_prob,
                       int32_t cand_stride, int32_t topk, float temperature, double max_refinement_km,
                       void* workspace, size_t workspace_bytes, float* out_lnglat, int64_t* out_cell,
                       float* best_logit, float* best_lnglat, int32_t* best_proto, int32_t* choice, void* stream);

/* Prototype-bank builder (reference models/proto_refiner.py:359-378 `_compute_protos_for_cell`, the arithmetic of
 * `load_prototypes`): data_views f32 [N, V, D] training embeddings (V = 4 for panoramas, 1 otherwise)
 *   -> data_mean_out f32 [N, D]  (view mean, what `data_emb` of pg_refiner_bank holds)
 *   -> proto_emb_out f32 [P, D]  (mean over member_idx[member_off[p] .. member_off[p+1]) of data_mean rows). */
int pg_bank_build(const float* data_views, int64_t N, int32_t V, int32_t D, const int64_t* member_off,
                  const int64_t* member_idx, int64_t P, float* data_mean_out, float* proto_emb_out, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Per-launch device timing (measurement aid for bench.py; not part of the reference-facing surface).
 * begin: subsequent launches of this library are bracketed with CUDA events on their launch stream;
 * end: synchronises the device, stops recording and returns the number of kernel families seen;
 * read: fills `n` entries (family name, summed milliseconds, launch count).
 * ------------------------------------------------------------------------------------------------- */
void pg_profile_begin(void);
int pg_profile_end(void);
void pg_profile_read(const char** names, float* ms, int32_t* counts, int32_t n);

/* ---------------------------------------------------------------------------------------------------
 * Multi-GPU exchange: the all-gather of per-rank results that the reference does with accelerate's
 * `gather` (preprocessing/embed.py:36-37; in evaluation the per-rank head outputs before the retrieval step).  One process
 * per GPU.  NCCL is resolved at run time by soname (libnccl.so.2): no link-time dependency, never loaded by single-GPU callers.
 * ------------------------------------------------------------------------------------------------- */
/* id128: HOST buffer of 128 bytes (ncclUniqueId).  Rank 0 creates it and hands it to the other ranks by any side channel. */
int pg_nccl_unique_id(void* id128);
/* Collective over the n_ranks processes (ncclCommInitRank on the current device).  *comm is an ncclComm_t. */
int pg_nccl_comm_create(const void* id128, int32_t n_ranks, int32_t rank, void** comm);
void pg_nccl_comm_destroy(void* comm);
/* recv[r * bytes_per_rank ...] = rank r's send[0 .. bytes_per_rank), for every r, enqueued on `stream` (ncclAllGather of
 * bytes: the caller packs whatever it exchanges - embeddings f32 [B, V, D], top-k candidates, predictions - into one
 * buffer, as pigeon_b200/dist.py does).  `comm` may also be a communicator the host created itself with NCCL. */
int pg_allgather_embeddings(void* comm, const void* send, void* recv, size_t bytes_per_rank, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Building blocks (exported for unit tests and for callers that fuse differently)
 * ------------------------------------------------------------------------------------------------- */
enum {
  PG_EPI_F16_BIAS = 0,       /* out fp16 = A W^T + bias                       (q/k/v projection) */
  PG_EPI_F16_BIAS_QGELU = 1, /* out fp16 = quick_gelu(A W^T + bias)            (mlp.fc1)          */
  PG_EPI_F32_BIAS_RESID = 2, /* out f32 += A W^T + bias                        (out_proj, mlp.fc2) */
  PG_EPI_F32_BIAS = 3        /* out f32 = A W^T + bias (bias may be NULL)                          */
};
/* D[M,N] = A[M,K] (fp16, row stride lda) * W[N,K]^T (fp16, row stride ldw), fp32 accumulate on tcgen05. */
int pg_gemm_f16(const void* a, int32_t lda, const void* w, int32_t ldw, void* out, int32_t ldo, const float* bias,
                int32_t M, int32_t N, int32_t K, int32_t epilogue, void* stream);
/* pg_gemm_f16 with bf16 operands (operand_bf16 != 0: BOTH a and w are bf16 — the tensor cores reject mixed pairs) and,
 * for PG_EPI_F32_BIAS_RESID, an out-of-place residual source
 * `resid` f32 [M, ldo] (NULL = update `out` in place). */
int pg_gemm_ex(const void* a, int32_t lda, const void* w, int32_t ldw, void* out, int32_t ldo, const float* bias,
               const float* resid, int32_t M, int32_t N, int32_t K, int32_t epilogue, int32_t operand_bf16, void* stream);
/* D[M, N] (+)= a^T w with BOTH operands stored contraction-major: a [K, lda] (M contiguous), w [K, ldw] (N contiguous) —
 * the weight-gradient form dW = dY^T X fed from row-major activations (MN-major UMMA operands, no transposes).
 * M % 64 == 0, M > 128, N % 256 == 0; epilogue PG_EPI_F32_BIAS (overwrite) or PG_EPI_F32_BIAS_RESID (accumulate in place). */
int pg_gemm_tn(const void* a, int32_t lda, const void* w, int32_t ldw, float* out, int32_t ldo, int32_t M, int32_t N,
               int32_t K, int32_t accumulate, int32_t operand_bf16, void* stream);
/* LayerNorm over the last dim: x f32 [rows, hidden] -> y fp16 [rows, hidden]. */
int pg_layernorm_f16(const float* x, void* y, const float* gamma, const float* beta, int64_t rows, int32_t hidden,
                     float eps, void* stream);
/* softmax(q k^T / 8) v per (view, head); qkv fp16 [n_views*seq, 3*heads*64] -> out fp16 [n_views*seq, heads*64]. */
int pg_attention_f16(const void* qkv, void* out, int32_t n_views, int32_t seq, int32_t heads, void* stream);
/* Measurement aid: the same op with the kernel chosen explicitly.  variant 0 = "pair" kernel (persistent, two query tiles per
 * CTA, KV blocks of 128, Q in tensor memory; poly = eighths of the exponentials evaluated on the FMA pipe, < 0 = default),
 * variant 1 = first-generation kernel (one tile per CTA, KV blocks of 32), variant 2 = "split" kernel (the pair kernel with
 * sixteen softmax warps: two independent column halves per S block, merged in the epilogue).  lse2 may be NULL. */
int pg_attention_f16_variant(const void* qkv, void* out, float* lse2, int32_t n_views, int32_t seq, int32_t heads,
                             int32_t variant, int32_t poly, void* stream);
/* Same, also writing lse2 f32 [n_views*heads, seq] (log2-sum-exp of the scaled logits) for the backward pass. */
int pg_attention_f16_lse(const void* qkv, void* out, float* lse2, int32_t n_views, int32_t seq, int32_t heads, void* stream);
/* Backward of the attention core: qkv f16 [n_views*seq, 3*heads*64], d_out f32 and out f16 [n_views*seq, heads*64],
 * lse2 from the forward -> dqkv bf16 [n_views*seq, 3*heads*64].  workspace: pg_attention_backward_workspace_bytes. */
size_t pg_attention_backward_workspace_bytes(int32_t n_views, int32_t seq, int32_t heads);
int pg_attention_backward(const void* qkv, const void* out, const float* d_out, const float* lse2, void* dqkv_bf16,
                          int32_t n_views, int32_t seq, int32_t heads, void* workspace, size_t workspace_bytes, void* stream);
/* LayerNorm backward: dx (+)= dLN(x)/dx . dy;  dgamma / dbeta (both or neither NULL) += their gradients;
 * dx_bf16 (nullable): bf16 copy of the final dx, the operand of the next data-gradient GEMM. */
int pg_layernorm_backward(const float* dy, const float* x, const float* gamma, float* dx, int32_t accumulate, float* dgamma,
                          float* dbeta, void* dx_bf16, int64_t rows, int32_t hidden, float eps, void* stream);
/* du bf16 [n] = dh f32 [n] * quick_gelu'(u f16 [n]) */
int pg_dgelu_bf16(const float* dh, const void* u, void* du, int64_t n, void* stream);
/* out bf16 [cols, ldo] = transpose(src [rows, lds]); src_type 0 = f32, 1 = f16, 2 = bf16 */
int pg_transpose_to_bf16(const void* src, int32_t src_type, int64_t lds, void* out, int64_t ldo, int64_t rows, int32_t cols,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PIGEON_B200_H_ */
