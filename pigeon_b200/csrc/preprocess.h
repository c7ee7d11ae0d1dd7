// Internal (C++) interface of the image pre-processing kernels; see preprocess.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

struct pg_image;

namespace pg {

size_t preprocess_workspace_bytes(const pg_image* images, int n, int size);
int preprocess_clip(const pg_image* images, int n, int size, const float* mean, const float* stdv, void* workspace,
                    size_t workspace_bytes, void* out, int out_f16, cudaStream_t stream);

}  // namespace pg
