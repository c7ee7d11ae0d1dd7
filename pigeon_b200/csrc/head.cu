// Geocell head around the tcgen05 GEMM: restates reference models/super_guessr.py:437 (4-view mean),
// :447-448 (cell_layer Linear + softmax), :454-455 (argmax + centroid lookup), :459 (top-k).
//
// The Linear is evaluated on the fp16 tensor-core path at fp32-faithful accuracy by an error-compensated
// split: x = hi + lo, W = Whi + Wlo (all fp16) and logits = [hi | lo | hi] . [Whi | Whi | Wlo]^T, i.e. one
// GEMM with K = 3*D whose dropped term lo.Wlo is O(2^-22).
#include "head.h"

#include <cuda_fp16.h>
#include <math.h>
#include <stdint.h>

#include "tma_host.h"

namespace pg {

namespace {

// in  : emb [B, V, D] fp32
// out : pooled [B, D] fp32 (mean over V, summed in view order then scaled, like torch.mean)
//       a3 [B, 3*D] fp16 = [hi | lo | hi]
__global__ void view_mean_split_kernel(const float* __restrict__ emb, float* __restrict__ pooled,
                                       __half* __restrict__ a3, int B, int V, int D) {
  const long total = (long)B * D;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int b = i / D, d = i % D;
    float s = 0.f;
    for (int v = 0; v < V; ++v) s += emb[((long)b * V + v) * D + d];
    const float x = s / V;
    pooled[i] = x;
    const __half hi = __float2half_rn(x);
    const __half lo = __float2half_rn(x - __half2float(hi));
    __half* row = a3 + (long)b * 3 * D;
    row[d] = hi;
    row[D + d] = lo;
    row[2 * D + d] = hi;
  }
}

// W [C, D] fp32 -> w3 [C, 3*D] fp16 = [Whi | Whi | Wlo]   (done once at handle creation)
__global__ void weight_split_kernel(const float* __restrict__ w, __half* __restrict__ w3, int C, int D) {
  const long total = (long)C * D;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = i / D, d = i % D;
    const float x = w[i];
    const __half hi = __float2half_rn(x);
    const __half lo = __float2half_rn(x - __half2float(hi));
    __half* row = w3 + (long)c * 3 * D;
    row[d] = hi;
    row[D + d] = hi;
    row[2 * D + d] = lo;
  }
}

struct ArgBest {
  float v;
  int i;
};
// "greater" with first-index tie-break; NaN ranks above everything (torch.argmax / topk semantics)
__device__ __forceinline__ bool better(float v, int i, float bv, int bi) {
  const bool vn = isnan(v), bn = isnan(bv);
  if (vn != bn) return vn;
  if (vn && bn) return i < bi;
  return (v > bv) || (v == bv && i < bi);
}
__device__ __forceinline__ ArgBest warp_best(ArgBest a) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, a.v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, a.i, o);
    if (better(ov, oi, a.v, a.i)) { a.v = ov; a.i = oi; }
  }
  return a;
}

// One CTA per sample. logits [B, C] fp32 -> probs [B, C], pred cell, pred (lng, lat) f64, top-k.
__global__ void __launch_bounds__(256)
softmax_topk_kernel(const float* __restrict__ logits, float* __restrict__ probs, long long* __restrict__ pred_cell,
                    double* __restrict__ pred_lnglat, float* __restrict__ topk_val, long long* __restrict__ topk_idx,
                    const double* __restrict__ centroids, int C, int k) {
  extern __shared__ float sp[];  // C floats
  __shared__ float red_v[8];
  __shared__ int red_i[8];
  __shared__ float bcast;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* lr = logits + (long)b * C;

  // row max
  float m = -INFINITY;
  for (int c = tid; c < C; c += 256) { const float x = lr[c]; sp[c] = x; m = fmaxf(m, x); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red_v[warp] = m;
  __syncthreads();
  if (tid == 0) { float t = red_v[0]; for (int w = 1; w < 8; ++w) t = fmaxf(t, red_v[w]); bcast = t; }
  __syncthreads();
  m = bcast;
  __syncthreads();
  // exp and sum
  float s = 0.f;
  for (int c = tid; c < C; c += 256) { const float e = expf(sp[c] - m); sp[c] = e; s += e; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) red_v[warp] = s;
  __syncthreads();
  if (tid == 0) { float t = 0.f; for (int w = 0; w < 8; ++w) t += red_v[w]; bcast = t; }
  __syncthreads();
  const float denom = bcast;
  for (int c = tid; c < C; c += 256) { const float p = sp[c] / denom; sp[c] = p; probs[(long)b * C + c] = p; }
  __syncthreads();

  // top-k by repeated block arg-max with removal; iteration 0 is the arg-max prediction
  for (int j = 0; j < k; ++j) {
    ArgBest best{-INFINITY, 0x7fffffff};
    for (int c = tid; c < C; c += 256) {
      const float p = sp[c];
      if (p != -1.f && better(p, c, best.v, best.i)) { best.v = p; best.i = c; }
    }
    best = warp_best(best);
    if (lane == 0) { red_v[warp] = best.v; red_i[warp] = best.i; }
    __syncthreads();
    if (tid == 0) {
      ArgBest t{red_v[0], red_i[0]};
      for (int w = 1; w < 8; ++w)
        if (better(red_v[w], red_i[w], t.v, t.i)) { t.v = red_v[w]; t.i = red_i[w]; }
      topk_val[(long)b * k + j] = t.v;
      topk_idx[(long)b * k + j] = t.i;
      if (j == 0) {
        pred_cell[b] = t.i;
        pred_lnglat[2 * (long)b + 0] = centroids[2 * (long)t.i + 0];
        pred_lnglat[2 * (long)b + 1] = centroids[2 * (long)t.i + 1];
      }
      sp[t.i] = -1.f;  // probabilities are >= 0, so -1 marks "taken"
    }
    __syncthreads();
  }
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_last_error("%s launch: %s", what, cudaGetErrorString(e)); return 1; }
  return 0;
}

}  // namespace

int view_mean_split(const float* emb, float* pooled, void* a3_f16, int B, int V, int D, cudaStream_t stream) {
  const long total = (long)B * D;
  int grid = (int)((total + 255) / 256);
  if (grid > 4096) grid = 4096;
  view_mean_split_kernel<<<grid, 256, 0, stream>>>(emb, pooled, reinterpret_cast<__half*>(a3_f16), B, V, D);
  return check_launch("view_mean_split");
}

int weight_split(const float* w, void* w3_f16, int C, int D, cudaStream_t stream) {
  const long total = (long)C * D;
  int grid = (int)((total + 255) / 256);
  if (grid > 4096) grid = 4096;
  weight_split_kernel<<<grid, 256, 0, stream>>>(w, reinterpret_cast<__half*>(w3_f16), C, D);
  return check_launch("weight_split");
}

int softmax_topk(const float* logits, float* probs, long long* pred_cell, double* pred_lnglat, float* topk_val,
                 long long* topk_idx, const double* centroids, int B, int C, int k, cudaStream_t stream) {
  if (k < 1 || k > C) { set_last_error("softmax_topk: k=%d out of range for C=%d", k, C); return 1; }
  if ((size_t)C * sizeof(float) > 200 * 1024) { set_last_error("softmax_topk: C=%d too large", C); return 1; }
  const size_t smem = (size_t)C * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(softmax_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_last_error("softmax_topk: smem attr: %s", cudaGetErrorString(e)); return 1; }
  }
  softmax_topk_kernel<<<B, 256, smem, stream>>>(logits, probs, pred_cell, pred_lnglat, topk_val, topk_idx, centroids,
                                                C, k);
  return check_launch("softmax_topk");
}

}  // namespace pg
