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
#include "ptx.cuh"
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

// ------------------------------------------------------------------------------------------------
// Cell-major scan (v2).  (query, candidate) pairs are counting-sorted by geocell; one CTA per geocell then reads
// that cell's prototype segment from HBM ONCE (re-reads per 8-query chunk come from L2) and scores every pair of the
// cell against it.  Per warp a 4-prototype x 8-query register tile: prototype rows stay in registers, the 8 query
// rows sit in shared memory interleaved in pairs so that one FFMA2 (packed fp32) advances two queries' dot products.
// Algorithmic bytes: sum over touched cells of P_c * D * 4, each once.
// ------------------------------------------------------------------------------------------------
constexpr int kQT = 8;  // queries per chunk
constexpr int kPT = 4;  // prototypes per warp register tile

__global__ void pair_hist_kernel(const RefinerBank bank, const long long* __restrict__ cand, int cand_stride, long B,
                                 int topk, int* __restrict__ cell_cnt, float* __restrict__ best_logit,
                                 float* __restrict__ best_lnglat, int* __restrict__ best_proto) {
  const long pairs = B * topk;
  for (long pair = (long)blockIdx.x * blockDim.x + threadIdx.x; pair < pairs; pair += (long)gridDim.x * blockDim.x) {
    const long b = pair / topk;
    const int j = pair % topk;
    const long long cell = cand[b * cand_stride + j];
    bool live = false;
    if (cell >= 0 && cell < bank.num_cells) live = bank.cell_off[cell + 1] > bank.cell_off[cell];
    if (live) {
      atomicAdd(&cell_cnt[cell], 1);
    } else {  // reference: protos[cell] is None -> logit -100000, prediction [0., 0.]   (proto_refiner.py:168-174)
      best_logit[pair] = -100000.f;
      best_lnglat[2 * pair] = 0.f;
      best_lnglat[2 * pair + 1] = 0.f;
      best_proto[pair] = -1;
    }
  }
}

// exclusive scan of cell_cnt[0..C) -> cell_start[0..C]; also zeroes the cursors. Single block.
__global__ void cell_scan_offsets_kernel(const int* __restrict__ cell_cnt, int* __restrict__ cell_start,
                                         int* __restrict__ cursor, int C) {
  __shared__ int carry;
  __shared__ int warp_tot[32];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < C; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int v = (i < C) ? cell_cnt[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      int t = (threadIdx.x < (blockDim.x >> 5)) ? warp_tot[threadIdx.x] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, t, o);
        if (threadIdx.x >= o) t += y;
      }
      warp_tot[threadIdx.x] = t;
    }
    __syncthreads();
    const int warp_off = (threadIdx.x >> 5) ? warp_tot[(threadIdx.x >> 5) - 1] : 0;
    if (i < C) {
      cell_start[i] = carry + warp_off + x - v;
      cursor[i] = 0;
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry += warp_off + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) cell_start[C] = carry;
}

__global__ void pair_scatter_kernel(const RefinerBank bank, const long long* __restrict__ cand, int cand_stride, long B,
                                    int topk, const int* __restrict__ cell_start, int* __restrict__ cursor,
                                    int* __restrict__ order) {
  const long pairs = B * topk;
  for (long pair = (long)blockIdx.x * blockDim.x + threadIdx.x; pair < pairs; pair += (long)gridDim.x * blockDim.x) {
    const long b = pair / topk;
    const int j = pair % topk;
    const long long cell = cand[b * cand_stride + j];
    if (cell >= 0 && cell < bank.num_cells && bank.cell_off[cell + 1] > bank.cell_off[cell])
      order[cell_start[cell] + atomicAdd(&cursor[cell], 1)] = (int)pair;
  }
}

// Shared-memory budget of the staged query set: 2 CTAs / SM.
constexpr int kStageBytes = 96 * 1024;

template <int NV4>
__global__ void __launch_bounds__(256, 2)
cell_major_scan_kernel(const RefinerBank bank, const float* __restrict__ q, const int* __restrict__ cell_start,
                       const int* __restrict__ order, int topk, float* __restrict__ best_logit,
                       float* __restrict__ best_lnglat, int* __restrict__ best_proto) {
  constexpr int D = NV4 * 128;
  constexpr int QS = (kStageBytes / (D * 4)) / kQT * kQT;   // queries staged per pass (32 at D = 768, 24 at D = 1024)
  static_assert(QS >= kQT, "stage holds at least one chunk");
  extern __shared__ float4 qs[];   // [QS/8][NV4][kQT/2][2][32] float4: chunk-major, lanes contiguous
  __shared__ float rbest_d[8][QS]; // per-warp running best of every staged query
  __shared__ int rbest_p[8][QS];   // ... as an offset from the cell's first prototype
  __shared__ float q_sqnorm[QS];
  const int cell = blockIdx.x;
  const int n_pairs = cell_start[cell + 1] - cell_start[cell];
  if (n_pairs == 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long lo = bank.cell_off[cell], hi = bank.cell_off[cell + 1];
  const int* my_order = order + cell_start[cell];

  // One pass over the cell's prototype segment per staged query set: when a cell has at most QS pairs (the common
  // case) the segment is read exactly once from HBM.
  for (int c0 = 0; c0 < n_pairs; c0 += QS) {
    const int nq = min(QS, n_pairs - c0);
    const int nch = (nq + kQT - 1) / kQT;
    __syncthreads();
    // stage the queries, two interleaved per float4 (see the FFMA2 loop):
    //   qs[(((ch * NV4 + i) * (kQT/2) + qp) * 2 + h) * 32 + lane] = for h = 0: (qA[4c], qB[4c], qA[4c+1], qB[4c+1]),
    //   for h = 1 the same for elements 4c+2, 4c+3;  A = ch*8 + 2qp, B = A + 1, c = lane + 32 i
    for (int idx = threadIdx.x; idx < nch * NV4 * (kQT / 2) * 32; idx += blockDim.x) {
      const int ln = idx & 31;
      const int qp = (idx >> 5) % (kQT / 2);
      const int i = ((idx >> 5) / (kQT / 2)) % NV4;
      const int ch = (idx >> 5) / (kQT / 2) / NV4;
      const int col4 = i * 32 + ln;
      const int qa = ch * kQT + 2 * qp;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b2 = a;
      if (qa < nq) a = reinterpret_cast<const float4*>(q + (long)(my_order[c0 + qa] / topk) * D)[col4];
      if (qa + 1 < nq) b2 = reinterpret_cast<const float4*>(q + (long)(my_order[c0 + qa + 1] / topk) * D)[col4];
      const size_t base = (size_t)(((ch * NV4 + i) * (kQT / 2) + qp) * 2) * 32 + ln;
      qs[base] = make_float4(a.x, b2.x, a.y, b2.y);
      qs[base + 32] = make_float4(a.z, b2.z, a.w, b2.w);
    }
    for (int qi = warp; qi < nch * kQT; qi += 8) {   // |q|^2, warp per query
      float acc_q = 0.f;
      if (qi < nq) {
        const float4* q4 = reinterpret_cast<const float4*>(q + (long)(my_order[c0 + qi] / topk) * D);
#pragma unroll
        for (int i = 0; i < NV4; ++i) {
          const float4 t = q4[lane + 32 * i];
          acc_q = fmaf(t.x, t.x, acc_q); acc_q = fmaf(t.y, t.y, acc_q);
          acc_q = fmaf(t.z, t.z, acc_q); acc_q = fmaf(t.w, t.w, acc_q);
        }
      }
      acc_q = warp_sum(acc_q);
      if (lane == 0) q_sqnorm[qi] = acc_q;
    }
    for (int i = lane; i < QS; i += 32) { rbest_d[warp][i] = INFINITY; rbest_p[warp][i] = 0x7fffffff; }
    __syncthreads();

    for (long p0 = lo + warp * kPT; p0 < hi; p0 += 8 * kPT) {
      // prototype tile: 4 rows in registers (all pieces requested up front), |p|^2 from the same registers
      float4 pr[NV4][kPT];
#pragma unroll
      for (int i = 0; i < NV4; ++i) {
#pragma unroll
        for (int a = 0; a < kPT; ++a) {
          const long pidx = (p0 + a < hi) ? p0 + a : hi - 1;   // clamp: duplicates are discarded below
          pr[i][a] = __ldg(reinterpret_cast<const float4*>(bank.proto_emb + pidx * D) + lane + 32 * i);
        }
      }
      float pn[kPT];
#pragma unroll
      for (int a = 0; a < kPT; ++a) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < NV4; ++i) {
          t = fmaf(pr[i][a].x, pr[i][a].x, t); t = fmaf(pr[i][a].y, pr[i][a].y, t);
          t = fmaf(pr[i][a].z, pr[i][a].z, t); t = fmaf(pr[i][a].w, pr[i][a].w, t);
        }
        pn[a] = warp_sum(t);
      }
      const float pn_mine = (lane >> 3) == 0 ? pn[0] : (lane >> 3) == 1 ? pn[1] : (lane >> 3) == 2 ? pn[2] : pn[3];
      const int pp_mine = (int)(p0 - lo) + (lane >> 3);
      const bool p_ok = (p0 + (lane >> 3)) < hi;

      for (int ch = 0; ch < nch; ++ch) {
        // d^2 = |p|^2 + |q|^2 - 2 p.q (the form torch.cdist itself uses beyond 25 rows): one FFMA2 per two (p, q)
        // element pairs.
        float2 acc[kPT][kQT / 2];
#pragma unroll
        for (int a = 0; a < kPT; ++a)
#pragma unroll
          for (int b = 0; b < kQT / 2; ++b) acc[a][b] = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < NV4; ++i) {
          const float4* qrow = qs + (size_t)((ch * NV4 + i) * (kQT / 2) * 2) * 32 + lane;
#pragma unroll
          for (int b = 0; b < kQT / 2; ++b) {
            const float4 q01 = qrow[(2 * b) * 32], q23 = qrow[(2 * b + 1) * 32];
#pragma unroll
            for (int a = 0; a < kPT; ++a) {
              acc[a][b] = ffma2(make_float2(pr[i][a].x, pr[i][a].x), make_float2(q01.x, q01.y), acc[a][b]);
              acc[a][b] = ffma2(make_float2(pr[i][a].y, pr[i][a].y), make_float2(q01.z, q01.w), acc[a][b]);
              acc[a][b] = ffma2(make_float2(pr[i][a].z, pr[i][a].z), make_float2(q23.x, q23.y), acc[a][b]);
              acc[a][b] = ffma2(make_float2(pr[i][a].w, pr[i][a].w), make_float2(q23.z, q23.w), acc[a][b]);
            }
          }
        }
        // 32 partial dot products (index = proto a * 8 + query) -> lane L ends with the warp total of index L
        float v[32];
#pragma unroll
        for (int a = 0; a < kPT; ++a)
#pragma unroll
          for (int b = 0; b < kQT / 2; ++b) { v[a * kQT + 2 * b] = acc[a][b].x; v[a * kQT + 2 * b + 1] = acc[a][b].y; }
#pragma unroll
        for (int off = 16, n = 16; off >= 1; off >>= 1, n >>= 1) {
#pragma unroll
          for (int k = 0; k < n; ++k) {
            const float send = (lane & off) ? v[k] : v[k + n];
            const float keep = (lane & off) ? v[k + n] : v[k];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
          }
        }
        // lane L: prototype (L >> 3) of the tile against query ch*8 + (L & 7); min over the 4 prototypes, first index wins
        float d2 = p_ok ? fmaxf(pn_mine + q_sqnorm[ch * kQT + (lane & 7)] - 2.f * v[0], 0.f) : INFINITY;
        int pp = pp_mine;
#pragma unroll
        for (int off = 8; off <= 16; off <<= 1) {
          const float od = __shfl_xor_sync(0xffffffffu, d2, off);
          const int op = __shfl_xor_sync(0xffffffffu, pp, off);
          if (od < d2 || (od == d2 && op < pp)) { d2 = od; pp = op; }
        }
        if (lane < kQT) {   // lanes 0..7 own the running best of query ch*8 + lane in this warp's private row
          const int qi = ch * kQT + lane;
          const float cd = rbest_d[warp][qi];
          const int cp = rbest_p[warp][qi];
          if (d2 < cd || (d2 == cd && pp < cp)) { rbest_d[warp][qi] = d2; rbest_p[warp][qi] = pp; }
        }
        __syncwarp();
      }
    }
    __syncthreads();
    // warp w finishes queries w, w+8, ...: cross-warp arg-min, then the farthest-member pick, then the outputs
    for (int qi = warp; qi < nq; qi += 8) {
      float bd = rbest_d[0][qi];
      int bpi = rbest_p[0][qi];
      for (int w = 1; w < 8; ++w) {
        const float od = rbest_d[w][qi];
        const int op = rbest_p[w][qi];
        if (od < bd || (od == bd && op < bpi)) { bd = od; bpi = op; }
      }
      const long bp = lo + bpi;
      const long pair = my_order[c0 + qi];
      const long b = pair / topk;
      float lng = bank.proto_lnglat[2 * bp], lat = bank.proto_lnglat[2 * bp + 1];
      if (bank.proto_count[bp] != 1) {
        float4 qv[NV4];
        const float4* q4 = reinterpret_cast<const float4*>(q + b * D);
#pragma unroll
        for (int i = 0; i < NV4; ++i) qv[i] = q4[lane + 32 * i];
        const long mlo = bank.member_off[bp], mhi = bank.member_off[bp + 1];
        float far = -INFINITY;
        long bm = -1;
        for (long mi = mlo; mi < mhi; ++mi) {
          const long idx = bank.member_idx[mi];
          const float dd = sqdist<NV4>(qv, bank.data_emb + idx * D, lane);
          if (dd > far) { far = dd; bm = idx; }
        }
        if (bm >= 0) { lng = bank.data_lnglat[2 * bm]; lat = bank.data_lnglat[2 * bm + 1]; }
      }
      if (lane == 0) {
        best_logit[pair] = -sqrtf(bd);
        best_lnglat[2 * pair] = lng;
        best_lnglat[2 * pair + 1] = lat;
        best_proto[pair] = (int)bp;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Bank builder (reference models/proto_refiner.py:359-378, `_compute_protos_for_cell`): prototype embedding = mean of its
// members' embeddings (members' 4-view mean first).  One warp per prototype, lanes strided over D as float4.
// ------------------------------------------------------------------------------------------------
__global__ void proto_mean_kernel(const float* __restrict__ data_emb, const long long* __restrict__ member_off,
                                  const long long* __restrict__ member_idx, long P, int D,
                                  float* __restrict__ proto_emb) {
  const int lane = threadIdx.x & 31;
  const long warps = ((long)gridDim.x * blockDim.x) >> 5;
  const int d4 = D >> 2;
  for (long p = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < P; p += warps) {
    const long lo = member_off[p], hi = member_off[p + 1];
    const float inv = 1.0f / (float)(hi - lo);
    for (int c = lane; c < d4; c += 32) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (long mi = lo; mi < hi; ++mi) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(data_emb + member_idx[mi] * D) + c);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      reinterpret_cast<float4*>(proto_emb + p * D)[c] = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
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

size_t refiner_sort_workspace_bytes(int num_cells, long pairs) {
  return (size_t)(3 * (num_cells + 1) + pairs) * sizeof(int) + 1024;
}

int refiner_scan_cell_major(const RefinerBank& bank, const float* q, const long long* cand, int cand_stride, long B,
                            int topk, void* sort_ws, float* best_logit, float* best_lnglat, int* best_proto,
                            int num_sms, cudaStream_t stream) {
  const long pairs = B * topk;
  if (pairs == 0) return 0;
  if (bank.dim % 128) { set_last_error("refiner: embedding dim %d not a multiple of 128", bank.dim); return 1; }
  const int C = bank.num_cells;
  int* cell_cnt = reinterpret_cast<int*>(sort_ws);
  int* cell_start = cell_cnt + (C + 1);
  int* cursor = cell_start + (C + 1);
  int* order = cursor + (C + 1);
  cudaError_t e = cudaMemsetAsync(cell_cnt, 0, (size_t)(C + 1) * sizeof(int), stream);
  if (e != cudaSuccess) { set_last_error("refiner: memset: %s", cudaGetErrorString(e)); return 1; }
  long blocks = (pairs + 255) / 256;
  if (blocks > (long)num_sms * 8) blocks = (long)num_sms * 8;
  {
    ProfScope prof("refiner_sort", stream);
    pair_hist_kernel<<<(int)blocks, 256, 0, stream>>>(bank, cand, cand_stride, B, topk, cell_cnt, best_logit, best_lnglat, best_proto);
    cell_scan_offsets_kernel<<<1, 1024, 0, stream>>>(cell_cnt, cell_start, cursor, C);
    pair_scatter_kernel<<<(int)blocks, 256, 0, stream>>>(bank, cand, cand_stride, B, topk, cell_start, cursor, order);
  }
  if (check_launch("refiner_sort")) return 1;
  // staged query set: QS = floor(96 KB / (D * 4)) rounded down to a multiple of 8 queries
  const size_t smem = (size_t)((kStageBytes / (bank.dim * 4)) / kQT * kQT) * bank.dim * sizeof(float);
  ProfScope prof("refiner_scan", stream);
  switch (bank.dim / 128) {
#define PG_CASE(N)                                                                                                   \
  case N: {                                                                                                          \
    auto kern = cell_major_scan_kernel<N>;                                                                           \
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);        \
    kern<<<C, 256, smem, stream>>>(bank, q, cell_start, order, topk, best_logit, best_lnglat, best_proto);           \
    break;                                                                                                           \
  }
    PG_CASE(1) PG_CASE(2) PG_CASE(4) PG_CASE(6) PG_CASE(8)
#undef PG_CASE
    default:
      set_last_error("refiner: embedding dim %d unsupported (need 128*{1,2,4,6,8})", bank.dim);
      return 1;
  }
  return check_launch("refiner_scan_cell_major");
}

int bank_build(const float* data_views, long N, int V, int D, const long long* member_off, const long long* member_idx,
               long P, float* data_mean, float* proto_emb, int num_sms, cudaStream_t stream) {
  if (D % 4) { set_last_error("bank_build: D=%d not a multiple of 4", D); return 1; }
  if (refiner_pool(data_views, data_mean, N, V, D, stream)) return 1;
  if (P == 0) return 0;
  long blocks = (P + 7) / 8;
  if (blocks > (long)num_sms * 16) blocks = (long)num_sms * 16;
  ProfScope prof("bank_proto_mean", stream);
  proto_mean_kernel<<<(int)blocks, 256, 0, stream>>>(data_mean, member_off, member_idx, P, D, proto_emb);
  return check_launch("bank_proto_mean");
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
