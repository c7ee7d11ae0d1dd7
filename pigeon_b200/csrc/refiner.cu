// ProtoRefiner retrieval on the GPU: restates reference models/proto_refiner.py:121-255,332-357 and
// preprocessing/geo_utils.py:40-55 as three kernels over a CSR-packed prototype bank resident in HBM.
//
//   1. pool      : (B, V, D) -> (B, D) view mean                                  (proto_refiner.py:139-140)
//   2. scan      : one warp per (query, candidate cell): Euclidean arg-min over the cell's prototypes
//                  (:176-181), then the within-cluster farthest-member pick (:233-255, arg-MAX, quirk kept)
//   3. finalize  : one thread per query: temperature softmax (no max-subtraction, :346-357) x candidate
//                  probabilities, arg-max, haversine max-refinement gate (:187-203), outputs (:219-231)
//
// No tensor cores: the scan is HBM/L2-bandwidth + FP32-FMA work; loads are 128-bit, lane-strided.
#include "refiner.h"

#include <math.h>
#include <stdint.h>

#include "prof.h"
#include "tma_host.h"

namespace pg {

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void pool_views_kernel(const float* __restrict__ emb, float* __restrict__ q, long B, int V, int D) {
  const long total = B * D;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long b = i / D;
    const int d = i % D;
    float s = 0.f;
    for (int v = 0; v < V; ++v) s += emb[(b * V + v) * D + d];
    q[i] = s / V;
  }
}

template <int NV4>
__device__ __forceinline__ float sqdist(const float4 (&q)[NV4], const float* __restrict__ row, int lane) {
  const float4* r4 = reinterpret_cast<const float4*>(row);
  float4 p[NV4];
#pragma unroll
  for (int i = 0; i < NV4; ++i) p[i] = __ldg(r4 + lane + 32 * i);
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < NV4; ++i) {
    const float a = p[i].x - q[i].x, b = p[i].y - q[i].y, c = p[i].z - q[i].z, d = p[i].w - q[i].w;
    acc = fmaf(a, a, acc);
    acc = fmaf(b, b, acc);
    acc = fmaf(c, c, acc);
    acc = fmaf(d, d, acc);
  }
  return warp_sum(acc);
}

template <int NV4>
__global__ void __launch_bounds__(256)
scan_kernel(const RefinerBank bank, const float* __restrict__ q, const long long* __restrict__ cand, int cand_stride,
            long B, int topk, float* __restrict__ best_logit, float* __restrict__ best_lnglat,
            int* __restrict__ best_proto) {
  constexpr int D = NV4 * 128;
  const int lane = threadIdx.x & 31;
  const long warps = ((long)gridDim.x * blockDim.x) >> 5;
  const long pairs = B * topk;
  for (long pair = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; pair < pairs; pair += warps) {
    const long b = pair / topk;
    const int j = pair % topk;
    const long long cell = cand[b * cand_stride + j];
    long lo = 0, hi = 0;
    if (cell >= 0 && cell < bank.num_cells) { lo = bank.cell_off[cell]; hi = bank.cell_off[cell + 1]; }
    if (hi <= lo) {
      // reference: protos[cell] is None -> logit -100000, prediction [0., 0.]   (proto_refiner.py:168-174)
      if (lane == 0) {
        best_logit[pair] = -100000.f;
        best_lnglat[2 * pair] = 0.f;
        best_lnglat[2 * pair + 1] = 0.f;
        best_proto[pair] = -1;
      }
      continue;
    }
    float4 qv[NV4];
    const float4* q4 = reinterpret_cast<const float4*>(q + b * D);
#pragma unroll
    for (int i = 0; i < NV4; ++i) qv[i] = q4[lane + 32 * i];

    float best = INFINITY;
    long bp = lo;
    for (long p = lo; p < hi; ++p) {
      const float d2 = sqdist<NV4>(qv, bank.proto_emb + p * D, lane);
      if (d2 < best) { best = d2; bp = p; }  // strict: first minimum wins (torch.argmax of -dist)
    }
    float lng = bank.proto_lnglat[2 * bp], lat = bank.proto_lnglat[2 * bp + 1];
    if (bank.proto_count[bp] != 1) {
      // within-cluster refinement: the member FARTHEST from the query (argmax of distances, :252-253)
      const long mlo = bank.member_off[bp], mhi = bank.member_off[bp + 1];
      float far = -INFINITY;
      long bm = -1;
      for (long mi = mlo; mi < mhi; ++mi) {
        const long idx = bank.member_idx[mi];
        const float d2 = sqdist<NV4>(qv, bank.data_emb + idx * D, lane);
        if (d2 > far) { far = d2; bm = idx; }
      }
      if (bm >= 0) { lng = bank.data_lnglat[2 * bm]; lat = bank.data_lnglat[2 * bm + 1]; }
    }
    if (lane == 0) {
      best_logit[pair] = -sqrtf(best);
      best_lnglat[2 * pair] = lng;
      best_lnglat[2 * pair + 1] = lat;
      best_proto[pair] = (int)bp;
    }
  }
}

// NaN ranks above everything, first index wins ties (torch.argmax).
__device__ __forceinline__ bool better(float v, float bv) {
  const bool vn = isnan(v), bn = isnan(bv);
  if (vn != bn) return vn;
  return v > bv;
}

// haversine(initial f64, refined f32) in km with the reference's type promotion (geo_utils.py:40-55):
// deg2rad of the fp32 point and the cosine of its latitude are fp32, everything else fp64.
__device__ double haversine_mixed(double lng0, double lat0, float lng1, float lat1) {
  const double kDeg = 3.14159265358979323846 / 180.0;
  const float kDegF = (float)kDeg;
  const double x_lng = lng0 * kDeg, x_lat = lat0 * kDeg;
  const float y_lng = lng1 * kDegF, y_lat = lat1 * kDegF;
  const double d_lng = (double)y_lng - x_lng, d_lat = (double)y_lat - x_lat;
  const double s_lat = sin(d_lat / 2), s_lng = sin(d_lng / 2);
  const double a = s_lat * s_lat + cos(x_lat) * (double)cosf(y_lat) * (s_lng * s_lng);
  const double c = 2 * asin(sqrt(a));
  return (6378137.0 * c) / 1000;
}

__global__ void finalize_kernel(const float* __restrict__ best_logit, const float* __restrict__ best_lnglat,
                                const long long* __restrict__ cand, const float* __restrict__ cand_prob,
                                int cand_stride, const double* __restrict__ init_lnglat, long B, int topk,
                                float temperature, double max_refinement, float* __restrict__ out_lnglat,
                                long long* __restrict__ out_cell, int* __restrict__ out_choice) {
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* lg = best_logit + b * topk;
  const float* cp = cand_prob + b * cand_stride;
  // temperature softmax without max-subtraction
  float sum = 0.f;
  for (int j = 0; j < topk; ++j) sum += expf(lg[j] / temperature);
  int refined = 0;
  float bestv = 0.f;
  int initial = 0;
  float besti = 0.f;
  for (int j = 0; j < topk; ++j) {
    const float pr = expf(lg[j] / temperature) / sum;
    const float f = cp[j] * pr;
    if (j == 0 || better(f, bestv)) { bestv = f; refined = j; }
    if (j == 0 || better(cp[j], besti)) { besti = cp[j]; initial = j; }
  }
  const float r_lng = best_lnglat[2 * (b * topk + refined)], r_lat = best_lnglat[2 * (b * topk + refined) + 1];
  const double dist = haversine_mixed(init_lnglat[2 * b], init_lnglat[2 * b + 1], r_lng, r_lat);
  // reference: `if distance > max_refinement: final_probs = c_probs[:topk]` -> falls back to the prior arg-max
  const int choice = (dist > max_refinement) ? initial : refined;
  out_lnglat[2 * b] = best_lnglat[2 * (b * topk + choice)];
  out_lnglat[2 * b + 1] = best_lnglat[2 * (b * topk + choice) + 1];
  out_cell[b] = cand[b * cand_stride + choice];
  if (out_choice) out_choice[b] = choice;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_last_error("%s launch: %s", what, cudaGetErrorString(e)); return 1; }
  return 0;
}

}  // namespace

int refiner_pool(const float* emb, float* q, long B, int V, int D, cudaStream_t stream) {
  const long total = B * D;
  int grid = (int)((total + 255) / 256);
  if (grid > 8192) grid = 8192;
  ProfScope prof("refiner_pool", stream);
  pool_views_kernel<<<grid, 256, 0, stream>>>(emb, q, B, V, D);
  return check_launch("refiner_pool");
}

int refiner_scan(const RefinerBank& bank, const float* q, const long long* cand, int cand_stride, long B, int topk,
                 float* best_logit, float* best_lnglat, int* best_proto, int num_sms, cudaStream_t stream) {
  const long pairs = B * topk;
  if (pairs == 0) return 0;
  long blocks = (pairs + 7) / 8;
  const long cap = (long)num_sms * 8;
  if (blocks > cap) blocks = cap;
  const int grid = (int)blocks;
  ProfScope prof("refiner_scan", stream);
  switch (bank.dim / 128) {
#define PG_CASE(N)                                                                                           \
  case N:                                                                                                    \
    scan_kernel<N><<<grid, 256, 0, stream>>>(bank, q, cand, cand_stride, B, topk, best_logit, best_lnglat,   \
                                             best_proto);                                                    \
    break;
    PG_CASE(1) PG_CASE(2) PG_CASE(4) PG_CASE(6) PG_CASE(8)
#undef PG_CASE
    default:
      set_last_error("refiner: embedding dim %d unsupported (need 128*{1,2,4,6,8})", bank.dim);
      return 1;
  }
  if (bank.dim % 128) { set_last_error("refiner: embedding dim %d not a multiple of 128", bank.dim); return 1; }
  return check_launch("refiner_scan");
}

int refiner_finalize(const float* best_logit, const float* best_lnglat, const long long* cand, const float* cand_prob,
                     int cand_stride, const double* init_lnglat, long B, int topk, float temperature,
                     double max_refinement, float* out_lnglat, long long* out_cell, int* out_choice,
                     cudaStream_t stream) {
  if (B == 0) return 0;
  const int grid = (int)((B + 127) / 128);
  ProfScope prof("refiner_finalize", stream);
  finalize_kernel<<<grid, 128, 0, stream>>>(best_logit, best_lnglat, cand, cand_prob, cand_stride, init_lnglat, B,
                                            topk, temperature, max_refinement, out_lnglat, out_cell, out_choice);
  return check_launch("refiner_finalize");
}

}  // namespace pg
