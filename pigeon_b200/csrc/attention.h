// Internal (C++) interface of the tcgen05 attention core; the C ABI in capi.cu wraps it.
#pragma once
#include <cuda_runtime.h>

namespace pg {

// qkv: fp16 [n_views*seq, 3*heads*64]; out: fp16 [n_views*seq, heads*64]. head_dim is 64.
int attention_f16(const void* qkv, void* out, int n_views, int seq, int heads, cudaStream_t stream);

}  // namespace pg
