// Internal (C++) interface of the memory-bound vision-tower kernels; see vit_misc.cu.
#pragma once
#include <cuda_runtime.h>

namespace pg {

int im2col(const void* pixels, int pixels_are_f16, void* out_f16, int n_views, int img, int patch, int kpad,
           int num_sms, cudaStream_t stream);
int layernorm_f16(const float* x, void* y_f16, const float* gamma, const float* beta, long rows, int hidden, float eps,
                  int num_sms, cudaStream_t stream);
int embed_preln(float* x, const float* cls, const float* pos, const float* gamma, const float* beta, long rows,
                int tokens, int hidden, float eps, int num_sms, cudaStream_t stream, float* e_out = nullptr,
                void* x16 = nullptr, float* stats = nullptr);   // x16 fp16 [rows, hidden] + stats float2 [rows, hidden / 128]:
                                                                // LayerNorm-folded tower (see gemm.h)
int token_mean(const float* x, float* out, int n_views, int tokens, int hidden, cudaStream_t stream);

}  // namespace pg
