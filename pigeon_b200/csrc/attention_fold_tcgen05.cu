// Multi-head self-attention core for the CLIP vision tower on sm_100a, third generation ("fold" kernel).
//
//   out[v, s, h*64 + :] = softmax_s'( q[v,s,h,:] . k[v,s',h,:] / sqrt(64) ) @ v[v,s',h,:]
//
// Restates the attention core of HF CLIPAttention.forward (bmm -> fp32 softmax -> bmm; no mask, no dropout in eval)
// that the reference reaches through models/clip_embedder.py:63 and models/super_guessr.py:395.  head_dim = 64.
//
// Same job structure, K/V ring and warp roles as the pair kernel (attention_pair_tcgen05.cu); what changed is WHERE the
// softmax arithmetic happens.  ncu on the pair kernel (profiles/r02_attn_pair_stalls.txt) shows its softmax warps issuing one
// instruction every fourth cycle with the MUFU unit 43 % busy and the tensor pipe 29 %: per logit they spend an FFMA (scale,
// subtract the reference maximum), an exponential, an FADD (row sum) and a pack.  Here the tensor pipe takes two of those:
//   * the scale log2(e)/8 is multiplied into Q when the staged tile is copied to TMEM, and the row's reference value rides
//     in a 5th k-step: Q' = [c q | -ref 0 .. 0] (80 columns), K' = [k | 1 0 .. 0] (a constant shared-memory tile), so that
//     S' = Q' K'^T arrives in TMEM as c q.k - ref, the exponent itself.  ref = (maximum of the row's first KV block) + 4,
//     rounded to fp16; the first block of a job is computed without the extra k-step and shifted in registers.
//   * the row sum l = sum_j P_j . 1 is column 64 of the output accumulator: the P V MMA runs with N = 80, its MN-major B
//     operand being the V block followed (leading-dimension offset of the descriptor = distance to the constant tile) by 16
//     columns [1 0 .. 0] — the same shared-memory tile that extends K.
//   Per logit the softmax warps are left with: tcgen05.ld, ex2 (MUFU, or the FMA-pipe polynomial for a share of them), pack.
//   * no running maximum, no rescale: P = 2^(S') is fp16, so a row is exact while its largest logit stays within
//     [ref - 24, ref + 16) log2 units.  A row that overflows shows l = inf in the epilogue, which then recomputes that row
//     on the CUDA cores with an fp32 online softmax (read from global; rare, slow, exact).  Logits more than 2^-20 below the
//     row maximum lose precision or vanish, as they do in any fp16-P attention kernel.
//
// TMEM columns: tile 0 S buffers [0,64) [64,128), tile 1 [128,192) [192,256) (P aliases the first 32 columns of its S buffer
// as packed fp16), O0 | l0 [256,336) O1 | l1 [336,416) (column 64 of each = row sum), Q'0 [416,456) Q'1 [464,504).
// Warps: 0-3 softmax tile 0, 4-7 softmax tile 1, 8-11 epilogue, 12 TMA producer, 13 / 14 MMA issuers of tile 0 / 1 (13 also
// allocates TMEM), 15 idle.
#include "attention.h"
#include "attention_dev.cuh"
#include "prof.h"
#include "ptx.cuh"
#include "tma_host.h"

#include <atomic>

namespace pg {

namespace {

using namespace attn_dev;

constexpr int kHeadDim = 64;
constexpr int kBlock = 128;                            // query rows per tile
constexpr int kSub = 64;                               // kv rows per block
constexpr int kSubBytes = kSub * kHeadDim * 2;         // 8 KB: one K or V block (a Q tile is two of them)
constexpr int kQBytes = 2 * kSubBytes;
constexpr int kSlots = 8;                              // K/V ring of 16 KB slots: {K_0, K_1} or {V_j, K_(j+2)} of one tile stream
constexpr int kSlotBytes = 2 * kSubBytes;
// Warp layout, H2 = softmax warps per query row quarter and tile (1: one thread per row and 64-column block; 2: two threads
// per row, 32 columns each): 8 * H2 softmax warps, 4 epilogue warps, then TMA producer, MMA issuers of tile 0 / 1, one idle warp.
// Warpgroups (4 warps) are the unit of setmaxnreg.
template <int H2>
struct Layout {
  static constexpr int kThreads = 256 * H2 + 256;
  static constexpr int kWarpEpi = 8 * H2, kWarpTma = kWarpEpi + 4, kWarpMma = kWarpTma + 1;
};
constexpr uint32_t kColS = 0, kColO = 256, kColQ = 416;
constexpr int kOCols = 80;                             // 64 output columns, the row sum, 15 zero columns
constexpr int kQCols = 40;                             // 80 fp16 per row: 64 of c*q, then -ref and 15 zeros
constexpr int kQStride = 48;                           // column distance between the two tiles' Q'
static_assert(kQCols <= kQStride && kColQ + kQStride + kQCols <= 512, "Q' does not fit");
constexpr int kTmemCols = 512;
constexpr float kRefMargin = 4.0f;                     // ref = first-block maximum + 4: P <= 2^-4 there, overflow 20 binades above
constexpr int kOnesBytes = kSub * 128;                 // constant tile: 64 rows (kv index) x 128 B, element 0 of every row = 1.0.  Read
                                                       // K-major as the 5th k-step of K' and MN-major as columns 64..79 of V'

struct Bars {
  uint64_t kv_full[kSlots], kv_empty[kSlots];
  uint64_t qs_full[2], qs_empty[2];   // TMA -> softmax (Q staging tile landed), softmax -> TMA (copied to TMEM)
  uint64_t q_ready[2];                // softmax -> MMA : Q_i is in TMEM
  uint64_t ref_ready[2];              // softmax -> MMA : -ref of this job is in Q'_i (blocks >= 1 may be issued)
  uint64_t s_full[2][2];              // MMA -> softmax : S block complete in buffer [tile][b]
  uint64_t p_ready[2][2];             // softmax -> MMA : P written over buffer [tile][b]
  uint64_t o_full[2];                 // MMA -> epilogue: O_i and l_i complete
  uint64_t o_free[2];                 // epilogue -> MMA : O_i, l_i and the row references were read
  uint64_t l_ready[2];                // softmax -> epilogue : row references published
  uint32_t tmem_ptr;
};
constexpr int kRingBytes = kSlots * kSlotBytes;
constexpr int kOffQ = kRingBytes;
constexpr int kOffOnes = kOffQ + 2 * kQBytes;
constexpr int kOffBars = kOffOnes + kOnesBytes;
constexpr int kOffRef = kOffBars + 1024;
constexpr int kOffXmax = kOffRef + 2 * kBlock * 4;       // [tile][half][row] partial first-block maxima (H2 = 2)
constexpr int kSmemBytes = 1024 + kOffXmax + 2 * 2 * kBlock * 4;

struct FoldArgs {
  const __half* qkv;
  __half* out;
  float* lse2;
  float scale_log2;
  int seq, hidden, heads, n_views;
  int nqt;            // query tiles per (view, head)
  int npair;          // full pairs of tiles per head
  int jobs_per_view, n_jobs;
};

struct Job {
  int view, h0, h1, t0, t1;
  bool a1;       // slot 1 holds a tile
  bool shared;   // both tiles read the same K/V stream
};

__device__ __forceinline__ Job decode_job(int job, const FoldArgs& a) {
  Job j;
  j.view = job / a.jobs_per_view;
  const int jv = job - j.view * a.jobs_per_view;
  const int full = a.heads * a.npair;
  if (jv < full) {
    j.h0 = j.h1 = jv / a.npair;
    j.t0 = 2 * (jv - j.h0 * a.npair);
    j.t1 = j.t0 + 1;
    j.a1 = true;
    j.shared = true;
  } else {
    const int k = jv - full;
    j.h0 = 2 * k;
    j.h1 = 2 * k + 1;
    j.t0 = j.t1 = a.nqt - 1;
    j.a1 = j.h1 < a.heads;
    j.shared = false;
  }
  return j;
}

// One 16-column chunk of a row: exponents r[16] (+ shift) -> P as 8 packed fp16 pairs.
// POLY: bit k set -> pair k of the chunk is exponentiated on the FMA pipe.
template <int POLY, bool SHIFT>
__device__ __forceinline__ void exp_chunk(const uint32_t* r, uint32_t* pk, float2 shift2) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float2 x = make_float2(__uint_as_float(r[2 * k]), __uint_as_float(r[2 * k + 1]));
    if (SHIFT) x = fadd2(x, shift2);
    float2 p;
    if ((POLY >> k) & 1) p = exp2_poly3(x);
    else p = make_float2(ex2(x.x), ex2(x.y));
    pk[k] = pack_half2(p.x, p.y);
  }
}

// Ragged chunk: only the first `rem` of the 16 columns are valid keys; P = 0 for the others.
__device__ __forceinline__ void exp_chunk_masked(const uint32_t* r, uint32_t* pk, int rem, float shift) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float p0 = (2 * k < rem) ? ex2(__uint_as_float(r[2 * k]) + shift) : 0.f;
    const float p1 = (2 * k + 1 < rem) ? ex2(__uint_as_float(r[2 * k + 1]) + shift) : 0.f;
    pk[k] = pack_half2(p0, p1);
  }
}

__device__ __forceinline__ float2 half2_bits_to_float2(uint32_t w) {
  return __half22float2(*reinterpret_cast<const __half2*>(&w));
}

// Exact fp32 online softmax of ONE query row on the CUDA cores, straight from global memory: the repair path of a row whose
// fp16 P overflowed (l = inf).  Two passes over the keys, 32 output dimensions each, to stay inside the epilogue warps'
// register budget.  Every lane of a warp may take it independently (K / V rows are warp-uniform addresses: L1 broadcasts).
__device__ __forceinline__ void exact_row(const __half* __restrict__ qkv, int view, int token, int h, int S, int hidden,
                                          float c, __half* __restrict__ out_row, float* lse2_out) {
  const size_t ld = (size_t)3 * hidden;
  const __half* base = qkv + (size_t)view * S * ld + h * kHeadDim;
  uint32_t q[32];
  {
    const uint4* q4 = reinterpret_cast<const uint4*>(base + (size_t)token * ld);
#pragma unroll
    for (int x = 0; x < 8; ++x) {
      const uint4 v = q4[x];
      q[4 * x] = v.x; q[4 * x + 1] = v.y; q[4 * x + 2] = v.z; q[4 * x + 3] = v.w;
    }
  }
  float m_final = 0.f, l_final = 1.f;
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    float m = -INFINITY, l = 0.f;
    float o[32];
#pragma unroll
    for (int x = 0; x < 32; ++x) o[x] = 0.f;
#pragma unroll 1
    for (int t = 0; t < S; ++t) {
      const uint4* k4 = reinterpret_cast<const uint4*>(base + (size_t)t * ld + hidden);
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int x = 0; x < 8; ++x) {
        const uint4 kv = k4[x];
        const uint32_t kw[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
        for (int y = 0; y < 4; ++y) {
          const float2 a = half2_bits_to_float2(q[4 * x + y]), b = half2_bits_to_float2(kw[y]);
          s0 = fmaf(a.x, b.x, s0);
          s1 = fmaf(a.y, b.y, s1);
        }
      }
      const float s = (s0 + s1) * c;
      if (s > m) {
        const float alpha = ex2(m - s);   // 0 on the first key
        l *= alpha;
#pragma unroll
        for (int x = 0; x < 32; ++x) o[x] *= alpha;
        m = s;
      }
      const float p = ex2(s - m);
      l += p;
      const uint4* v4 = reinterpret_cast<const uint4*>(base + (size_t)t * ld + 2 * hidden + 32 * half);
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        const uint4 vv = v4[x];
        const uint32_t vw[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
        for (int y = 0; y < 4; ++y) {
          const float2 b = half2_bits_to_float2(vw[y]);
          o[8 * x + 2 * y] = fmaf(p, b.x, o[8 * x + 2 * y]);
          o[8 * x + 2 * y + 1] = fmaf(p, b.y, o[8 * x + 2 * y + 1]);
        }
      }
    }
    const float inv_l = 1.0f / l;
    uint4* o4 = reinterpret_cast<uint4*>(out_row + 32 * half);
#pragma unroll
    for (int x = 0; x < 4; ++x)
      o4[x] = make_uint4(pack_half2(o[8 * x] * inv_l, o[8 * x + 1] * inv_l), pack_half2(o[8 * x + 2] * inv_l, o[8 * x + 3] * inv_l),
                         pack_half2(o[8 * x + 4] * inv_l, o[8 * x + 5] * inv_l), pack_half2(o[8 * x + 6] * inv_l, o[8 * x + 7] * inv_l));
    m_final = m;
    l_final = l;
  }
  if (lse2_out != nullptr) *lse2_out = m_final + log2f(l_final);
}

template <int POLY, int H2>
__global__ void __launch_bounds__(Layout<H2>::kThreads, 1)
attention_fold_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const FoldArgs args) {
  constexpr int kThreads = Layout<H2>::kThreads;
  constexpr int kWarpEpi = Layout<H2>::kWarpEpi, kWarpTma = Layout<H2>::kWarpTma, kWarpMma = Layout<H2>::kWarpMma;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_kv = smem;
  uint8_t* smem_q = smem + kOffQ;
  uint8_t* smem_ones = smem + kOffOnes;
  Bars* bars = reinterpret_cast<Bars*>(smem + kOffBars);
  float* refs = reinterpret_cast<float*>(smem + kOffRef);   // [slot][row]: the row's reference (log2 units), for lse2

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = args.seq;
  const int nb = (S + kSub - 1) / kSub;                        // KV blocks (10 for S = 577)
  const int last_valid = S - (nb - 1) * kSub;                  // valid kv columns in the last block (1)
  const int last_n = (last_valid + 15) & ~15;                  // MMA N / K extent of the last block (16)
  const int n_jobs = args.n_jobs;
  const int stride = gridDim.x;

  // constant B operand (128-byte swizzle: 16-byte piece c of row r sits at piece c ^ (r & 7)): element 0 of every row = 1.0
  for (int x = threadIdx.x; x < kOnesBytes / 16; x += kThreads) {
    const int r = x >> 3, piece = x & 7;
    reinterpret_cast<uint4*>(smem_ones)[x] = make_uint4(piece == (r & 7) ? 0x00003C00u : 0u, 0u, 0u, 0u);
  }
  fence_proxy_async_smem();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    for (int s = 0; s < kSlots; ++s) {
      mbar_init(&bars->kv_full[s], 1);
      mbar_init(&bars->kv_empty[s], 2);   // two tcgen05.commit arrivals: one per MMA warp (shared stream) or both from the owner
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->qs_full[i], 1);
      mbar_init(&bars->qs_empty[i], 4 * H2);
      mbar_init(&bars->q_ready[i], 4 * H2);
      mbar_init(&bars->ref_ready[i], 4);
      for (int b = 0; b < 2; ++b) {
        mbar_init(&bars->s_full[i][b], 1);
        mbar_init(&bars->p_ready[i][b], 4 * H2);
      }
      mbar_init(&bars->o_full[i], 1);
      mbar_init(&bars->o_free[i], 4);
      mbar_init(&bars->l_ready[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == kWarpMma) {
    tmem_alloc(&bars->tmem_ptr, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_ptr;

  // register budget: 512 threads start with 128 each; the softmax warps take the share the other roles do not need
  // (8*168 + 4*96 + 4*80 = 16*128 per lane).  setmaxnreg must be executed with the SAME value by all four warps of a
  // warpgroup: warps 12-15 (TMA, MMA, idle) form one.
  // K/V ring protocol (as in the pair kernel).  A job's stream of a tile is 1 + nb slots: {K_0, K_1}, then {V_j, K_(j+2)} for
  // j = 0 .. nb-1.  Two tiles of one (view, head) share ONE stream (both MMA warps read every slot and commit once each);
  // otherwise the two streams are interleaved slot by slot (the owner commits twice), or there is a single stream.
  if (warp < kWarpEpi) {
    // ---------------------------------------------------------------- softmax warps: thread = query row = TMEM lane
   if constexpr (H2 == 2) {
    // ---- two threads per query row: warps 8 i + 0..3 take columns [0, 32) of every S block of tile i, warps 8 i + 4..7 columns
    // [32, 64); each writes its P over the first 16 columns of its own half.  768 threads: 80 registers each, no setmaxnreg.
    const int i = warp >> 3;            // tile slot
    const int half = (warp >> 2) & 1;   // column half
    const int wq = warp & 3;            // lane quarter
    const int row = wq * 32 + lane;
    const uint32_t lane_base = uint32_t(wq * 32) << 16;
    const uint32_t s_tmem0 = tmem_base + lane_base + kColS + 128 * i + 32 * half;
    const uint32_t q_tmem = tmem_base + lane_base + kColQ + kQStride * i;
    const uint32_t q_smem = smem_u32(smem_q + i * kQBytes) + row * 128;
    float* xmax = reinterpret_cast<float*>(smem + kOffXmax) + i * 2 * kBlock;
    const int pair_bar = 1 + i * 4 + wq;   // named barrier of the two warps that share these 32 rows
    const float c = args.scale_log2;
    uint32_t n_q = 0, n_blk = 0, n_job = 0;
    bool q_done = false;

    auto copy_q = [&]() {   // this half's 32 of the 64 head dimensions
      mbar_wait(&bars->qs_full[i], n_q & 1);
      ++n_q;
      uint32_t qr[16];
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        const uint4 v = lds128(q_smem + (((4 * half + ch) ^ (row & 7)) << 4));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int y = 0; y < 4; ++y) {
          const float2 f = half2_bits_to_float2(w[y]);
          qr[4 * ch + y] = pack_half2(f.x * c, f.y * c);
        }
      }
      tmem_st16p(q_tmem + 16 * half, qr);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&bars->q_ready[i]);
        mbar_arrive(&bars->qs_empty[i]);
      }
    };

    for (int job = blockIdx.x; job < n_jobs; job += stride) {
      const Job jb = decode_job(job, args);
      if (i == 1 && !jb.a1) continue;
      const int t = i ? jb.t1 : jb.t0;
      const bool warp_active = t * kBlock + wq * 32 < S;
      bool next_active = job + stride < n_jobs;
      if (next_active && i == 1) next_active = decode_job(job + stride, args).a1;
      if (!q_done) copy_q();
      q_done = false;
      float ref = 0.f;

      for (int j = 0; j < nb; ++j) {
        const int b = n_blk & 1;
        const uint32_t s_tmem = s_tmem0 + kSub * b;
        mbar_wait(&bars->s_full[i][b], (n_blk >> 1) & 1);
        tc_fence_after();
        if (j == nb - 1 && next_active) {
          copy_q();
          q_done = true;
        }
        const bool full = (j < nb - 1) || (last_valid == kSub);
        const int valid = full ? kSub : last_valid;
        int hv = valid - 32 * half;              // valid columns of this half
        hv = hv < 0 ? 0 : (hv > 32 ? 32 : hv);
        uint32_t r0[16], r1[16];
        if (j == 0) {
          // ---- reference of the job: each half takes the maximum of its own columns, the two warps of a row quarter swap
          // them through shared memory (one 64-thread named barrier per job), both derive the same fp16 reference
          float pm = -INFINITY;
          if (warp_active) {
            if (hv == 32) {
              tmem_ld16p(s_tmem, r0);
              tmem_ld16p(s_tmem + 16, r1);
              tmem_ld_wait16(r0);
              tmem_ld_wait16(r1);
              float b0 = -INFINITY, b1 = -INFINITY;
#pragma unroll
              for (int x = 0; x < 16; x += 2) {
                b0 = fmax3(b0, __uint_as_float(r0[x]), __uint_as_float(r0[x + 1]));
                b1 = fmax3(b1, __uint_as_float(r1[x]), __uint_as_float(r1[x + 1]));
              }
              pm = fmaxf(b0, b1);
            } else {
              for (int ch = 0; ch * 16 < hv; ++ch) {
                uint32_t r[16];
                tmem_ld16p(s_tmem + 16 * ch, r);
                tmem_ld_wait16(r);
#pragma unroll
                for (int x = 0; x < 16; ++x)
                  if (ch * 16 + x < hv) pm = fmaxf(pm, __uint_as_float(r[x]));
              }
            }
            xmax[half * kBlock + row] = pm;
          }
          asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
          if (warp_active) ref = __half2float(__float2half_rn(fmaxf(pm, xmax[(half ^ 1) * kBlock + row]) + kRefMargin));
          if (half == 0) {
            if (warp_active) {
              uint32_t qe[8];
              qe[0] = pack_half2(-ref, 0.f);
#pragma unroll
              for (int x = 1; x < 8; ++x) qe[x] = 0u;
              tmem_st8p(q_tmem + 32, qe);
              tmem_st_wait();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->ref_ready[i]);
          }
        }
        if (warp_active) {
          if (hv == 32) {
            uint32_t pk[16];
            if (j == 0) {
              const float2 sh2 = make_float2(-ref, -ref);
              exp_chunk<POLY, true>(r0, pk, sh2);
              exp_chunk<POLY, true>(r1, pk + 8, sh2);
            } else {
              const float2 z2 = make_float2(0.f, 0.f);
              tmem_ld16p(s_tmem, r0);
              tmem_ld16p(s_tmem + 16, r1);
              tmem_ld_wait16(r0);
              exp_chunk<POLY, false>(r0, pk, z2);
              tmem_ld_wait16(r1);
              exp_chunk<POLY, false>(r1, pk + 8, z2);
            }
            tmem_st16p(s_tmem, pk);
          } else {
            const int nfull = hv >> 4, rem = hv & 15;
            const float shift = (j == 0) ? -ref : 0.f;
            const float2 sh2 = make_float2(shift, shift);
            for (int ch = 0; ch < nfull; ++ch) {
              uint32_t r[16], pk8[8];
              tmem_ld16p(s_tmem + 16 * ch, r);
              tmem_ld_wait16(r);
              exp_chunk<0, true>(r, pk8, sh2);
              tmem_st8p(s_tmem + 8 * ch, pk8);
            }
            if (rem) {
              uint32_t r[16], pk8[8];
              tmem_ld16p(s_tmem + 16 * nfull, r);
              tmem_ld_wait16(r);
              exp_chunk_masked(r, pk8, rem, shift);
              tmem_st8p(s_tmem + 8 * nfull, pk8);
            }
          }
          tmem_st_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->p_ready[i][b]);
        ++n_blk;
      }

      if (half == 0) {
        if (n_job > 0) mbar_wait(&bars->o_free[i], (n_job - 1) & 1);
        refs[i * kBlock + row] = ref;
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->l_ready[i]);
      }
      ++n_job;
    }
   } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 168;");
    const int i = warp >> 2;        // tile slot
    const int wq = warp & 3;        // lane quarter
    const int row = wq * 32 + lane;
    const uint32_t lane_base = uint32_t(wq * 32) << 16;
    const uint32_t s_tmem0 = tmem_base + lane_base + kColS + 128 * i;
    const uint32_t q_tmem = tmem_base + lane_base + kColQ + kQStride * i;
    const uint32_t q_smem = smem_u32(smem_q + i * kQBytes) + row * 128;
    const float c = args.scale_log2;
    uint32_t n_q = 0, n_blk = 0, n_job = 0;   // n_blk: running block count of this tile (block g lives in S buffer g & 1)
    bool q_done = false;

    // staged Q tile (128-byte swizzle: 16-byte chunk ch of row r sits at chunk ch ^ (r & 7)) -> c * q -> TMEM A operand
    auto copy_q = [&]() {
      mbar_wait(&bars->qs_full[i], n_q & 1);
      ++n_q;
      uint32_t qr[32];
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        const uint4 v = lds128(q_smem + ((ch ^ (row & 7)) << 4));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int y = 0; y < 4; ++y) {
          const float2 f = half2_bits_to_float2(w[y]);
          qr[4 * ch + y] = pack_half2(f.x * c, f.y * c);
        }
      }
      tmem_st32p(q_tmem, qr);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&bars->q_ready[i]);
        mbar_arrive(&bars->qs_empty[i]);
      }
    };

    for (int job = blockIdx.x; job < n_jobs; job += stride) {
      const Job jb = decode_job(job, args);
      if (i == 1 && !jb.a1) continue;
      const int t = i ? jb.t1 : jb.t0;
      const bool warp_active = t * kBlock + wq * 32 < S;   // a warp of padding rows only keeps the barriers moving
      bool next_active = job + stride < n_jobs;
      if (next_active && i == 1) next_active = decode_job(job + stride, args).a1;
      if (!q_done) copy_q();
      q_done = false;

      float ref = 0.f;   // this row's reference, fp16-representable, log2 units

      for (int j = 0; j < nb; ++j) {
        const int b = n_blk & 1;
        const uint32_t s_tmem = s_tmem0 + kSub * b;
        mbar_wait(&bars->s_full[i][b], (n_blk >> 1) & 1);
        tc_fence_after();
        if (j == nb - 1 && next_active) {   // every S_i MMA of this job has retired: Q_i may be replaced
          copy_q();
          q_done = true;
        }
        const bool full = (j < nb - 1) || (last_valid == kSub);
        const int valid = full ? kSub : last_valid;
        // Q'[:, 64] = -ref and the go-ahead for the MMA warp to issue this job's later blocks
        auto publish_ref = [&]() {
          if (warp_active) {
            uint32_t qe[8];
            qe[0] = pack_half2(-ref, 0.f);
#pragma unroll
            for (int x = 1; x < 8; ++x) qe[x] = 0u;
            tmem_st8p(q_tmem + 32, qe);
            tmem_st_wait();
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->ref_ready[i]);
        };
        if (j == 0 && full) {
          // ---- first block of a job: all 64 logits (c q.k, log2 units) into registers once; reference = their maximum + margin
          // as an fp16 value; then P = 2^(x - ref) from the same registers
          if (!warp_active) {
            publish_ref();
          } else {
            uint32_t r0[16], r1[16], r2[16], r3[16], pk[32];
            tmem_ld16p(s_tmem, r0);
            tmem_ld16p(s_tmem + 16, r1);
            tmem_ld16p(s_tmem + 32, r2);
            tmem_ld16p(s_tmem + 48, r3);
            tmem_ld_wait16(r0);
            tmem_ld_wait16(r1);
            tmem_ld_wait16(r2);
            tmem_ld_wait16(r3);
            float b0 = -INFINITY, b1 = -INFINITY, b2 = -INFINITY, b3 = -INFINITY;
#pragma unroll
            for (int x = 0; x < 16; x += 2) {
              b0 = fmax3(b0, __uint_as_float(r0[x]), __uint_as_float(r0[x + 1]));
              b1 = fmax3(b1, __uint_as_float(r1[x]), __uint_as_float(r1[x + 1]));
              b2 = fmax3(b2, __uint_as_float(r2[x]), __uint_as_float(r2[x + 1]));
              b3 = fmax3(b3, __uint_as_float(r3[x]), __uint_as_float(r3[x + 1]));
            }
            ref = __half2float(__float2half_rn(fmaxf(fmaxf(b0, b1), fmaxf(b2, b3)) + kRefMargin));
            publish_ref();
            const float2 sh2 = make_float2(-ref, -ref);
            exp_chunk<POLY, true>(r0, pk, sh2);
            exp_chunk<POLY, true>(r1, pk + 8, sh2);
            exp_chunk<POLY, true>(r2, pk + 16, sh2);
            exp_chunk<POLY, true>(r3, pk + 24, sh2);
            tmem_st32p(s_tmem, pk);
          }
        } else if (full) {
          if (warp_active) {
            // ---- a later block of 64 valid key columns: the exponents come out of the MMA; one branch-free basic block, all 64
            // in registers before P (packed fp16) goes over the first 32 columns of the same buffer
            uint32_t pk[32];
            uint32_t ra[16], rb[16];
            const float2 z2 = make_float2(0.f, 0.f);
            tmem_ld16p(s_tmem, ra);
#pragma unroll
            for (int ch = 0; ch < 4; ch += 2) {   // software-pipelined: the next chunk's load is in flight during the arithmetic
              tmem_ld_wait16(ra);
              tmem_ld16p(s_tmem + 16 * (ch + 1), rb);
              exp_chunk<POLY, false>(ra, pk + 8 * ch, z2);
              tmem_ld_wait16(rb);
              if (ch + 2 < 4) tmem_ld16p(s_tmem + 16 * (ch + 2), ra);
              exp_chunk<POLY, false>(rb, pk + 8 * (ch + 1), z2);
            }
            tmem_st32p(s_tmem, pk);
          }
        } else {
          // ---- ragged last block (also the first one when the sequence is shorter than a block): nfull whole 16-column
          // chunks + `rem` valid columns of one more (P of chunk ch lands on S columns of chunks <= ch, already read)
          const int nfull = valid >> 4, rem = valid & 15;
          if (j == 0) {
            if (warp_active) {
              float b0 = -INFINITY;
              for (int ch = 0; ch * 16 < valid; ++ch) {
                uint32_t r[16];
                tmem_ld16p(s_tmem + 16 * ch, r);
                tmem_ld_wait16(r);
#pragma unroll
                for (int x = 0; x < 16; ++x)
                  if (ch * 16 + x < valid) b0 = fmaxf(b0, __uint_as_float(r[x]));
              }
              ref = __half2float(__float2half_rn(b0 + kRefMargin));
            }
            publish_ref();
          }
          if (warp_active) {
            const float shift = (j == 0) ? -ref : 0.f;
            const float2 sh2 = make_float2(shift, shift);
            for (int ch = 0; ch < nfull; ++ch) {
              uint32_t r[16], pk8[8];
              tmem_ld16p(s_tmem + 16 * ch, r);
              tmem_ld_wait16(r);
              exp_chunk<0, true>(r, pk8, sh2);
              tmem_st8p(s_tmem + 8 * ch, pk8);
            }
            if (rem) {
              uint32_t r[16], pk8[8];
              tmem_ld16p(s_tmem + 16 * nfull, r);
              tmem_ld_wait16(r);
              exp_chunk_masked(r, pk8, rem, shift);
              tmem_st8p(s_tmem + 8 * nfull, pk8);
            }
          }
        }
        if (!warp_active) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->p_ready[i][b]);
          ++n_blk;
          continue;
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->p_ready[i][b]);
        ++n_blk;
      }

      // publish the row references; the epilogue of this slot's previous job must have read its own first
      if (n_job > 0) mbar_wait(&bars->o_free[i], (n_job - 1) & 1);
      refs[i * kBlock + row] = ref;
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->l_ready[i]);
      ++n_job;
    }
   }
  } else if (warp < kWarpTma) {
    // ---------------------------------------------------------------- epilogue warps: O / l -> fp16 -> global
    if constexpr (H2 == 1) asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
    const int wq = warp & 3;
    const int row = wq * 32 + lane;
    const uint32_t lane_base = uint32_t(wq * 32) << 16;
    const int heads = args.hidden / kHeadDim;
    uint32_t n_e[2] = {0, 0};
    for (int job = blockIdx.x; job < n_jobs; job += stride) {
      const Job jb = decode_job(job, args);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        if (i == 1 && !jb.a1) continue;
        const int h = i ? jb.h1 : jb.h0;
        const int token = (i ? jb.t1 : jb.t0) * kBlock + row;
        mbar_wait(&bars->l_ready[i], n_e[i] & 1);
        mbar_wait(&bars->o_full[i], n_e[i] & 1);
        ++n_e[i];
        tc_fence_after();
        const float ref = refs[i * kBlock + row];
        const uint32_t o_tmem = tmem_base + lane_base + kColO + kOCols * i;
        const float l_sum = tmem_ld1(o_tmem + 64);
        const float inv_l = 1.0f / l_sum;
        uint32_t pk[32];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t o[32];
          tmem_ld32(o_tmem + 32 * hh, o);
          tmem_ld_wait();
#pragma unroll
          for (int x = 0; x < 16; ++x)
            pk[16 * hh + x] = pack_half2(__uint_as_float(o[2 * x]) * inv_l, __uint_as_float(o[2 * x + 1]) * inv_l);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->o_free[i]);
        if (token < S) {
          const size_t grow = (size_t)jb.view * S + token;
          __half* orow = args.out + grow * args.hidden + h * kHeadDim;
          float* lse = args.lse2 != nullptr ? args.lse2 + ((size_t)jb.view * heads + h) * S + token : nullptr;
          if (l_sum <= 3.0e38f) {
            uint4* o4 = reinterpret_cast<uint4*>(orow);
#pragma unroll
            for (int x = 0; x < 8; ++x) o4[x] = make_uint4(pk[4 * x], pk[4 * x + 1], pk[4 * x + 2], pk[4 * x + 3]);
            if (lse != nullptr) *lse = ref + log2f(l_sum);
          } else {
            // some P of this row left the fp16 range (a logit 20 binades above the first block's maximum): exact repair
            exact_row(args.qkv, jb.view, token, h, S, args.hidden, args.scale_log2, orow, lse);
          }
        }
      }
    }
  // (every warpgroup's setmaxnreg sits inside its own role branch, and no role calls a non-inlined function: ptxas compiles a
  // shared callee for the smallest budget and then holds every caller to it)
  } else {
   if constexpr (H2 == 1) asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
   if (warp > kWarpMma + 1) {
    // idle warp of the last warpgroup
   } else if (warp == kWarpTma) {
    // ---------------------------------------------------------------- TMA producer (warp-uniform, one elected lane issues)
    if (blockIdx.x < n_jobs) {
      uint32_t pos = 0;              // ring position
      uint32_t nq0 = 0, nq1 = 0;
      // slot contents: rows [r_a, +64) at column col_a, and (if col_b >= 0) rows [r_b, +64) at column col_b
      auto load_slot = [&](int col_a, int r_a, int col_b, int r_b) {
        const int slot = pos % kSlots;
        mbar_wait(&bars->kv_empty[slot], ((pos / kSlots) & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&bars->kv_full[slot], col_b >= 0 ? kSlotBytes : kSubBytes);
          tma_load_2d(smem_kv + slot * kSlotBytes, &tmap_qkv, &bars->kv_full[slot], col_a, r_a);
          if (col_b >= 0) tma_load_2d(smem_kv + slot * kSlotBytes + kSubBytes, &tmap_qkv, &bars->kv_full[slot], col_b, r_b);
        }
        __syncwarp();
        ++pos;
      };
      auto load_q = [&](auto I_, const Job& jb) {
        constexpr int I = decltype(I_)::value;
        uint32_t& nq = I ? nq1 : nq0;
        mbar_wait(&bars->qs_empty[I], (nq & 1) ^ 1);
        if (elect_one()) {
          const int col = (I ? jb.h1 : jb.h0) * kHeadDim, row = jb.view * S + (I ? jb.t1 : jb.t0) * kBlock;
          mbar_arrive_expect_tx(&bars->qs_full[I], kQBytes);
          tma_load_2d(smem_q + I * kQBytes, &tmap_qkv, &bars->qs_full[I], col, row);
          tma_load_2d(smem_q + I * kQBytes + kSubBytes, &tmap_qkv, &bars->qs_full[I], col, row + kSub);
        }
        __syncwarp();
        ++nq;
      };
      {
        const Job j0 = decode_job(blockIdx.x, args);
        load_q(Slot<0>{}, j0);
        if (j0.a1) load_q(Slot<1>{}, j0);
      }
      for (int job = blockIdx.x; job < n_jobs; job += stride) {
        const Job jb = decode_job(job, args);
        const int row0 = jb.view * S;
        const int nstream = (jb.shared || !jb.a1) ? 1 : 2;
        for (int step = 0; step <= nb; ++step) {          // step 0 = {K_0, K_1}; step j + 1 = {V_j, K_(j+2)}
          for (int t = 0; t < nstream; ++t) {
            const int h = t ? jb.h1 : jb.h0;
            const int kc = args.hidden + h * kHeadDim, vc = 2 * args.hidden + h * kHeadDim;
            if (step == 0) load_slot(kc, row0, nb > 1 ? kc : -1, row0 + kSub);
            else load_slot(vc, row0 + (step - 1) * kSub, step + 1 < nb ? kc : -1, row0 + (step + 1) * kSub);
          }
          if (step == 1 && job + stride < n_jobs) {       // the next job's Q tiles, once this job's first blocks are on their way
            const Job jn = decode_job(job + stride, args);
            load_q(Slot<0>{}, jn);
            if (jn.a1) load_q(Slot<1>{}, jn);
          }
        }
      }
    }
   } else {
    // ---------------------------------------------------------------- MMA issuers: warp 13 -> tile 0, warp 14 -> tile 1
    // (warp-uniform; one elected lane issues)
    const int I = warp - kWarpMma;
    uint32_t pos = 0;                                    // ring position at the start of the current job
    uint32_t g = 0;                                      // running block count of this tile: block g lives in S buffer g & 1
    uint32_t n_j = 0;                                    // jobs this tile took part in
    bool pre = false;                                    // S of this job's first block was issued during the previous job
    // smem descriptors of a slot's first (offset 0) / second (offset 8 KB) block: K is K-major (rows of 128 B, k-step = 32 B),
    // V is MN-major (row = kv index, k-step = 16 rows); the two constant tiles are K-major
    const uint64_t k_desc0 = make_smem_desc(smem_u32(smem_kv), 16, 1024, kLayoutSw128);
    const uint64_t v_desc0 = make_smem_desc(smem_u32(smem_kv), 0, 1024, kLayoutSw128);
    const uint64_t ke_desc = make_smem_desc(smem_u32(smem_ones), 16, 1024, kLayoutSw128);
    // V' = [V_j | ones tile]: the second 64-wide group of the MN-major operand starts LBO bytes after the first, i.e. at the
    // constant tile; LBO depends on the slot (descriptor bits [16,30), units of 16 bytes)
    const uint32_t ones_off16 = (smem_u32(smem_ones) - smem_u32(smem_kv)) >> 4;
    const uint32_t idesc_s = make_idesc_f16(kBlock, kSub, 0, 0);
    const uint32_t idesc_s_last = make_idesc_f16(kBlock, last_n, 0, 0);
    const uint32_t idesc_pv = make_idesc_f16(kBlock, kOCols, 0, 1);      // B (= V') is MN-major, N = 64 + 16
    const uint32_t s_base = tmem_base + kColS + 128 * I, o_base = tmem_base + kColO + kOCols * I;
    const uint32_t q_base = tmem_base + kColQ + kQStride * I;
    // TMEM column (relative to the S buffer) of the packed-fp16 P columns of k-step k: contiguous with one softmax thread per
    // row; with two, the second half's P sits over its own logits (columns 32..47)
    auto p_col = [](int k) -> uint32_t { return (H2 == 2 && k >= 2) ? 32u + 8u * (k - 2) : 8u * k; };
    // S[buf] = Q'_I K'_j^T, j = block index inside its job; block 0 without the reference k-step   (inside an elected region)
    auto issue_s = [&](uint64_t kd, int j, uint32_t buf) {
      const uint32_t idesc = (j == nb - 1) ? idesc_s_last : idesc_s;
      const uint32_t d = s_base + kSub * buf;
#pragma unroll
      for (int k = 0; k < kHeadDim / 16; ++k) umma_ts(d, q_base + 8 * k, kd + 2 * k, idesc, k != 0);
      if (j != 0) umma_ts(d, q_base + 32, ke_desc, idesc, 1);
      tc_commit(&bars->s_full[I][buf]);
    };
    for (int job = blockIdx.x; job < n_jobs; job += stride) {
      const Job jb = decode_job(job, args);
      const int nstream = (jb.shared || !jb.a1) ? 1 : 2;
      const uint32_t pos_job = pos;
      pos += (uint32_t)(nb + 1) * nstream;
      if (I == 1 && !jb.a1) continue;
      const uint32_t first = (nstream == 2) ? I : 0;
      const int ncommit = jb.shared ? 1 : 2;
      // the next job of this tile: its first S block is issued while this job's last two blocks are in the softmax warps
      bool has_next = (job + stride < n_jobs) && nb >= 2;
      uint32_t next_p = 0;
      if (has_next) {
        const Job jn = decode_job(job + stride, args);
        if (I == 1 && !jn.a1) has_next = false;
        const int ns = (jn.shared || !jn.a1) ? 1 : 2;
        next_p = pos + ((ns == 2) ? I : 0);
      }
      for (int step = 0; step <= nb; ++step) {
        const uint32_t p = pos_job + first + (uint32_t)step * nstream;
        const int slot = p % kSlots;
        const int j = step - 1;
        const uint64_t d0 = (uint64_t)(slot * (kSlotBytes >> 4)), d1 = d0 + (kSubBytes >> 4);
        if (step == 0) {
          // slot {K_0, K_1}: S_0 (unless issued during the previous job), then S'_1 once the softmax warps have put this
          // job's reference into Q'
          if (!pre) {
            mbar_wait(&bars->q_ready[I], n_j & 1);
            mbar_wait(&bars->kv_full[slot], (p / kSlots) & 1);
            tc_fence_after();
            if (elect_one()) issue_s(k_desc0 + d0, 0, g & 1);
            __syncwarp();
          }
          if (nb > 1) {
            mbar_wait(&bars->ref_ready[I], n_j & 1);
            tc_fence_after();
          }
          if (elect_one()) {
            if (nb > 1) issue_s(k_desc0 + d1, 1, (g + 1) & 1);
            tc_commit(&bars->kv_empty[slot]);
            if (ncommit == 2) tc_commit(&bars->kv_empty[slot]);
          }
          __syncwarp();
          continue;
        }
        const uint32_t gb = g + j;
        mbar_wait(&bars->p_ready[I][gb & 1], (gb >> 1) & 1);
        if (j == 0) mbar_wait(&bars->o_free[I], (n_j & 1) ^ 1);   // the previous job's O_I / l_I were read out
        mbar_wait(&bars->kv_full[slot], (p / kSlots) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t vd = v_desc0 + d0 + ((uint64_t)(ones_off16 - (uint32_t)d0) << 16);
          const uint32_t a = s_base + kSub * (gb & 1);            // P (packed fp16 over the S buffer)
          const int ksteps = (j == nb - 1) ? last_n / 16 : kSub / 16;
          umma_ts(o_base, a, vd, idesc_pv, j != 0);
          if (j == nb - 1) {
            for (int k = 1; k < ksteps; ++k) umma_ts(o_base, a + p_col(k), vd + 128 * k, idesc_pv, 1);
          } else {
#pragma unroll
            for (int k = 1; k < kSub / 16; ++k) umma_ts(o_base, a + p_col(k), vd + 128 * k, idesc_pv, 1);
          }
          if (j == nb - 1) tc_commit(&bars->o_full[I]);
          if (j + 2 < nb) issue_s(k_desc0 + d1, j + 2, gb & 1);
          tc_commit(&bars->kv_empty[slot]);
          if (ncommit == 2) tc_commit(&bars->kv_empty[slot]);
        }
        __syncwarp();
        if (has_next && j == nb - 2) {
          // block j's buffer is free again (its P V is issued): the next job's block 0 goes there.  Q of the next job
          // reaches TMEM when the softmax warps start this job's last block.
          const int slot_n = next_p % kSlots;
          mbar_wait(&bars->q_ready[I], (n_j + 1) & 1);
          mbar_wait(&bars->kv_full[slot_n], (next_p / kSlots) & 1);
          tc_fence_after();
          if (elect_one()) issue_s(k_desc0 + (uint64_t)(slot_n * (kSlotBytes >> 4)), 0, (g + nb) & 1);
          __syncwarp();
        }
      }
      pre = has_next;
      g += nb;
      ++n_j;
    }
   }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

std::atomic<int> g_attr_set[64][10];   // per device and kernel variant: dynamic shared memory opt-in done

template <int POLY, int H2>
int launch_fold(int variant, const void* qkv, void* out, int n_views, int seq, int heads, cudaStream_t stream,
                float* lse2) {
  const int hidden = heads * kHeadDim;
  CUtensorMap tm;
  if (make_tmap_f16_2d(&tm, qkv, (uint64_t)n_views * seq, 3 * hidden, 3 * hidden, kSub, kHeadDim)) return 1;
  auto kern = attention_fold_kernel<POLY, H2>;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!g_attr_set[dev][variant].load(std::memory_order_acquire)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) { set_last_error("attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 1; }
    g_attr_set[dev][variant].store(1, std::memory_order_release);
  }
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms <= 0) sms = 148;
  FoldArgs a;
  a.qkv = reinterpret_cast<const __half*>(qkv);
  a.out = reinterpret_cast<__half*>(out);
  a.lse2 = lse2;
  a.scale_log2 = 0.125f * 1.4426950408889634f;
  a.seq = seq;
  a.hidden = hidden;
  a.heads = heads;
  a.n_views = n_views;
  a.nqt = (seq + kBlock - 1) / kBlock;
  a.npair = a.nqt / 2;
  a.jobs_per_view = heads * a.npair + ((a.nqt & 1) ? (heads + 1) / 2 : 0);
  a.n_jobs = n_views * a.jobs_per_view;
  const int grid = a.n_jobs < sms ? a.n_jobs : sms;
  ProfScope prof("attention", stream);
  kern<<<grid, Layout<H2>::kThreads, kSmemBytes, stream>>>(tm, a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_last_error("attention launch: %s", cudaGetErrorString(e)); return 1; }
  return 0;
}

}  // namespace

// poly: share of the exponentials evaluated on the FMA pipe, in eighths (0 .. 4 of every 8 pairs; other values = 2).
// poly + 10: the variant with two softmax threads per query row (768 threads).
int attention_fold_f16(const void* qkv, void* out, int n_views, int seq, int heads, cudaStream_t stream, float* lse2,
                       int poly) {
  if (n_views <= 0) return 0;
  switch (poly) {
    case 0: return launch_fold<0x00, 1>(0, qkv, out, n_views, seq, heads, stream, lse2);
    case 1: return launch_fold<0x08, 1>(1, qkv, out, n_views, seq, heads, stream, lse2);
    case 3: return launch_fold<0x4A, 1>(3, qkv, out, n_views, seq, heads, stream, lse2);
    case 4: return launch_fold<0xAA, 1>(4, qkv, out, n_views, seq, heads, stream, lse2);
    case 10: return launch_fold<0x00, 2>(5, qkv, out, n_views, seq, heads, stream, lse2);
    case 12: return launch_fold<0x88, 2>(6, qkv, out, n_views, seq, heads, stream, lse2);
    case 13: return launch_fold<0x4A, 2>(7, qkv, out, n_views, seq, heads, stream, lse2);
    case 14: return launch_fold<0xAA, 2>(8, qkv, out, n_views, seq, heads, stream, lse2);
    default: return launch_fold<0x88, 1>(2, qkv, out, n_views, seq, heads, stream, lse2);
  }
}

}  // namespace pg
