// Internal (C++) interface of the geocell-head kernels; see head.cu.
#pragma once
#include <cuda_runtime.h>

namespace pg {

int view_mean_split(const float* emb, float* pooled, void* a3_f16, int B, int V, int D, cudaStream_t stream);
int weight_split(const float* w, void* w3_f16, int C, int D, cudaStream_t stream);
int softmax_topk(const float* logits, float* probs, long long* pred_cell, double* pred_lnglat, float* topk_val,
                 long long* topk_idx, const double* centroids, int B, int C, int k, cudaStream_t stream);

int ce_loss(const float* logits, int B, int C, int mode, const long long* labels_idx, const float* soft,
            const double* labels_lnglat, const double* centroids, double smoothing, double* per_sample,
            double* loss_out, float* dlogits, double grad_scale, cudaStream_t stream);

}  // namespace pg
