// Internal (C++) interface of the fine-tune-step kernels (head backward, AdamW); see train.cu.
#pragma once
#include <cuda_runtime.h>

namespace pg {

// C[M, N] = beta * C + op(A) . B with B row-major [K, N]; a_transposed: A stored [K, M], else [M, K].  fp32 FMA.
int sgemm_f32(bool a_transposed, const float* A, const float* B, float* C, int M, int N, int K, float beta,
              cudaStream_t stream);
// out[c] = beta * out[c] + sum_r x[r, c]
int column_sum_f32(const float* x, float* out, int rows, int cols, float beta, cudaStream_t stream);
// torch.optim.AdamW single-tensor update (amsgrad off, maximize off), fp32 state.
int adamw_step(float* p, const float* g, float* m, float* v, long n, double lr, double beta1, double beta2, double eps,
               double weight_decay, long step, double grad_scale, cudaStream_t stream);

}  // namespace pg
