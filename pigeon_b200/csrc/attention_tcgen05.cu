// Multi-head self-attention core for the CLIP vision tower on sm_100a (tcgen05 + TMEM + TMA).
//
//   out[v, s, h*64 + :] = softmax_s'( q[v,s,h,:] . k[v,s',h,:] / sqrt(64) ) @ v[v,s',h,:]
//
// Restates the attention core of HF CLIPAttention.forward (bmm -> fp32 softmax -> bmm; no mask, no
// dropout in eval) that the reference reaches through models/clip_embedder.py:63 and
// models/super_guessr.py:395.  head_dim is fixed to 64 (ViT-L/14: 16 heads x 64).
//
// Input  qkv : fp16 [n_views * S, 3 * hidden]   row = (view, token); cols = [q | k | v], head-major inside
// Output out : fp16 [n_views * S, hidden]
//
// One CTA per (q-tile of 128 rows, head, view); 192 threads.  Two tilings (AttnCfg): KV blocks of 32 with two S buffers
// (128 TMEM columns, 40 KB smem -> FOUR co-resident CTAs per SM; the default) or KV blocks of 64 with three S buffers
// (256 TMEM columns -> two CTAs per SM).
//   warp 0     TMA producer: Q tile once, then K/V tiles (64 x 64 halves, 128B swizzle) through a 6-slot ring
//   warp 1     TMEM allocator + MMA issuer (tcgen05.mma cta_group::1, M = 128)
//   warps 2-5  softmax, one TMEM lane (= one query row) per thread
//
// Single pass, flash-style, KV blocks of 64:  S_j = Q K_j^T lands in one of THREE TMEM buffers so that the MMA warp runs
// two blocks ahead of the softmax warps (its barrier round trip is longer than the softmax work of one 64-wide block).  The softmax warps keep a running row maximum m and row sum l,
// write P_j = exp2(S_j*c - m*c) as packed fp16 over the S_j columns, and the MMA warp accumulates O += P_j V_j with P
// as the TMEM A operand and V as an MN-major smem B operand.  m is only raised (and O, l rescaled in TMEM) when the
// block maximum exceeds it by more than 2^8 in the exponent domain ("lazy rescale"): P then stays <= 256, exact in fp16
// range, and the rescale path is rare.  Final O / l -> fp16.
#include "attention.h"
#include "prof.h"
#include "ptx.cuh"
#include "tma_host.h"

#include <atomic>
#include <stdlib.h>
#include <string.h>

namespace pg {

namespace {

constexpr int kHeadDim = 64;
constexpr int kBlockQ = 128;
constexpr int kQBytes = kBlockQ * kHeadDim * 2;       // 16 KB
constexpr int kThreads = 192;

// Two tilings of the same kernel:
//   <64, 3, 2>  KV blocks of 64, three S buffers, 256 TMEM columns, 2 CTAs / SM   (S0 S1 S2 [0,192)  O [192,256))
//   <32, 2, 4>  KV blocks of 32, two S buffers, 128 TMEM columns, 4 CTAs / SM     (S0 S1 [0,64)      O [64,128))
template <int KV, int NBUF, int CTAS, int POLY = 0, int SLOTS = 6, bool HALVES = false>
struct AttnCfg {
  // HALVES: a 64-wide block is read from TMEM twice, 32 columns at a time (maximum sweep, then exponential sweep), so that the
  // row never needs more than 32 logit registers (fits the 80-register budget of 4 CTAs per SM without spilling).
  static constexpr bool kHalves = HALVES;
  // POLY = n > 0: every n-th group of four exponentials is evaluated on the FMA pipe (Cody-Waite + degree-4 minimax
  // polynomial, 2.9e-6 relative) instead of MUFU.EX2, to relieve the 16/clk/SM special-function unit.
  static constexpr int kPolyStride = POLY;
  static constexpr int kBlockKV = KV;
  static constexpr int kNumSBuf = NBUF;
  static constexpr int kCtasPerSm = CTAS;
  static constexpr int kSlots = SLOTS;
  static constexpr int kTileBytes = KV * kHeadDim * 2;
  static constexpr int kOCol = NBUF * KV;
  static constexpr int kTmemCols = (NBUF * KV + kHeadDim <= 128) ? 128 : 256;
  static constexpr int kSmemBytes = kQBytes + kSlots * kTileBytes + 1024 + 256;
  static_assert(NBUF * KV + kHeadDim <= kTmemCols, "TMEM budget");
};
constexpr float kRescaleThreshold = 8.0f;             // log2 domain

struct AttnArgs {
  int seq;      // tokens per view (577)
  int hidden;   // heads * 64
  __half* out;  // [n_views*seq, hidden]
  float scale_log2;  // (1/sqrt(64)) * log2(e)
  float* lse2;       // optional [n_views * heads, seq]: log2-domain log-sum-exp of the scaled logits (for the backward pass)
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void tmem_ld16_(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// 2^x for a pair, x <= ~8, on the FMA / ALU pipes: n = round(x) via the 1.5*2^23 magic constant, f = x - n in
// [-0.5, 0.5], 2^f by a degree-4 minimax polynomial (max relative error 2.9e-6), 2^n by an exponent-field add.
__device__ __forceinline__ float2 exp2_poly(float2 x) {
  x.x = fmaxf(x.x, -125.f);
  x.y = fmaxf(x.y, -125.f);
  const float2 magic = make_float2(12582912.f, 12582912.f);
  const float2 t = fadd2(x, magic);
  const float2 n = fsub2(t, magic);
  const float2 f = fsub2(x, n);
  float2 p = ffma2(make_float2(0.009582852944731712f, 0.009582852944731712f), f,
                   make_float2(0.055906426161527634f, 0.055906426161527634f));
  p = ffma2(p, f, make_float2(0.24024099111557007f, 0.24024099111557007f));
  p = ffma2(p, f, make_float2(0.6931241750717163f, 0.6931241750717163f));
  p = ffma2(p, f, make_float2(1.0f, 1.0f));
  p.x = __int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23));
  p.y = __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23));
  return p;
}

__device__ __forceinline__ void tmem_st8_(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st32_(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

template <class Cfg>
__global__ void __launch_bounds__(kThreads, Cfg::kCtasPerSm)
attention_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnArgs args) {
  constexpr int kBlockKV = Cfg::kBlockKV, kNumSBuf = Cfg::kNumSBuf, kSlots = Cfg::kSlots, kTileBytes = Cfg::kTileBytes;
  constexpr int kOCol = Cfg::kOCol, kTmemCols = Cfg::kTmemCols;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;
  uint8_t* smem_kv = smem + kQBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kQBytes + kSlots * kTileBytes);
  uint64_t* full_bar = bars;             // [kSlots]  TMA -> MMA
  uint64_t* empty_bar = bars + kSlots;   // [kSlots]  MMA -> TMA
  uint64_t* q_full = bars + 2 * kSlots;
  uint64_t* s_full = q_full + 1;                 // [3] MMA -> softmax : S block complete in TMEM buffer b
  uint64_t* p_ready = s_full + kNumSBuf;         // [3] softmax -> MMA : P written over buffer b (4 warps arrive)
  uint64_t* pv_done = p_ready + kNumSBuf;        // [3] MMA -> softmax : P_j V_j retired (barrier j % 3)
  uint64_t* o_full = pv_done + kNumSBuf;         // MMA -> softmax : the last P V retired, O complete
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x, head = blockIdx.y, view = blockIdx.z;
  const int S = args.seq;
  const int nb = (S + kBlockKV - 1) / kBlockKV;                 // KV blocks (10 for S = 577)
  const int last_valid = S - (nb - 1) * kBlockKV;               // valid kv columns in the last block (1)
  const int last_n = (last_valid + 15) & ~15;                   // MMA N / K extent of the last block (16)
  const int row0 = view * S;                                    // first row of this view in qkv / out
  const int q_col = head * kHeadDim, k_col = args.hidden + q_col, v_col = 2 * args.hidden + q_col;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    for (int s = 0; s < kSlots; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(q_full, 1);
    for (int b = 0; b < kNumSBuf; ++b) {
      mbar_init(&s_full[b], 1);
      mbar_init(&p_ready[b], 4);
      mbar_init(&pv_done[b], 1);
    }
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, kQBytes);
#pragma unroll
      for (int part = 0; part < kBlockQ / kBlockKV; ++part)   // the tensor map's box is one KV tile (kBlockKV rows)
        tma_load_2d(smem_q + part * kTileBytes, &tmap_qkv, q_full, q_col, row0 + q_tile * kBlockQ + part * kBlockKV);
      int slot = 0;
      uint32_t phase = 0;
      auto load = [&](int col, int blk) {
        mbar_wait(&empty_bar[slot], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[slot], kTileBytes);
        tma_load_2d(smem_kv + slot * kTileBytes, &tmap_qkv, &full_bar[slot], col, row0 + blk * kBlockKV);
        if (++slot == kSlots) { slot = 0; phase ^= 1; }
      };
      // consumption order of the MMA warp: K0, K1, K2, then (V_j, K_{j+3}) for j = 0..
      for (int j = 0; j < kNumSBuf && j < nb; ++j) load(k_col, j);
      for (int j = 0; j < nb; ++j) {
        load(v_col, j);
        if (j + kNumSBuf < nb) load(k_col, j + kNumSBuf);
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      const uint32_t o_tmem = tmem_base + kOCol;
      const uint32_t q_addr = smem_u32(smem_q);
      mbar_wait(q_full, 0);
      tc_fence_after();

      auto issue_s = [&](int j) {
        const int n = (j == nb - 1) ? last_n : kBlockKV;
        const uint32_t idesc = make_idesc_f16(kBlockQ, n, 0, 0);
        const uint32_t s_tmem = tmem_base + (j % kNumSBuf) * kBlockKV;
        mbar_wait(&full_bar[slot], phase);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(smem_kv + slot * kTileBytes);
#pragma unroll
        for (int k = 0; k < kHeadDim / 16; ++k) {
          const uint64_t a_desc = make_smem_desc(q_addr + k * 32, 16, 1024, kLayoutSw128);
          const uint64_t b_desc = make_smem_desc(k_addr + k * 32, 16, 1024, kLayoutSw128);
          umma_ss(s_tmem, a_desc, b_desc, idesc, k != 0);
        }
        tc_commit(&empty_bar[slot]);
        tc_commit(&s_full[j % kNumSBuf]);
        if (++slot == kSlots) { slot = 0; phase ^= 1; }
      };
      auto issue_pv = [&](int j) {
        const int kext = (j == nb - 1) ? last_n : kBlockKV;              // contraction extent = kv rows of this block
        const uint32_t idesc = make_idesc_f16(kBlockQ, kHeadDim, 0, 1);  // B (= V) is MN-major
        const uint32_t p_tmem = tmem_base + (j % kNumSBuf) * kBlockKV;   // P aliases the S buffer, fp16 pairs
        mbar_wait(&full_bar[slot], phase);
        tc_fence_after();
        const uint32_t v_addr = smem_u32(smem_kv + slot * kTileBytes);
        for (int k = 0; k < kext / 16; ++k) {
          // V tile: row = kv index (128 B each, 64 halves of head_dim), 8-row groups 1024 B apart.
          const uint64_t b_desc = make_smem_desc(v_addr + k * 16 * 128, 1024, 1024, kLayoutSw128);
          umma_ts(o_tmem, p_tmem + k * 8, b_desc, idesc, (j | k) != 0);
        }
        tc_commit(&empty_bar[slot]);
        tc_commit(&pv_done[j % kNumSBuf]);
        if (++slot == kSlots) { slot = 0; phase ^= 1; }
      };

      for (int j = 0; j < kNumSBuf && j < nb; ++j) issue_s(j);
      for (int j = 0; j < nb; ++j) {
        mbar_wait(&p_ready[j % kNumSBuf], (j / kNumSBuf) & 1);
        tc_fence_after();
        issue_pv(j);
        if (j + kNumSBuf < nb) issue_s(j + kNumSBuf);  // overwrites P_j only after P_j V_j (in-order tensor pipe)
      }
      tc_commit(o_full);
    }
  } else {
    // ---------------------------------------------------------------- softmax warps
    const int q = warp & 3;
    const uint32_t lane_base = uint32_t(q * 32) << 16;
    const uint32_t o_tmem = tmem_base + lane_base + kOCol;
    const int q_row = q_tile * kBlockQ + q * 32 + lane;  // token index inside the view
    const float c = args.scale_log2;

    float m = -INFINITY;   // reference maximum currently used in the exponent (raw logit units)
    float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;

    for (int j = 0; j < nb; ++j) {
      const uint32_t s_tmem = tmem_base + lane_base + (j % kNumSBuf) * kBlockKV;
      const bool tail = (j == nb - 1) && (last_valid < kBlockKV);
      mbar_wait(&s_full[j % kNumSBuf], (j / kNumSBuf) & 1);
      tc_fence_after();

      uint32_t r[Cfg::kHalves ? 32 : kBlockKV];
      float bm = -INFINITY;
      if (!tail && Cfg::kHalves) {
#pragma unroll
        for (int h = 0; h < kBlockKV / 32; ++h) {
          tmem_ld32(s_tmem + 32 * h, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
          tmem_ld_wait();
          float b0 = -INFINITY, b1 = -INFINITY, b2 = -INFINITY, b3 = -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            b0 = fmaxf(b0, __uint_as_float(r[i]));
            b1 = fmaxf(b1, __uint_as_float(r[i + 1]));
            b2 = fmaxf(b2, __uint_as_float(r[i + 2]));
            b3 = fmaxf(b3, __uint_as_float(r[i + 3]));
          }
          bm = fmaxf(bm, fmaxf(fmaxf(b0, b1), fmaxf(b2, b3)));
        }
      } else if (!tail) {
#pragma unroll
        for (int h = 0; h < kBlockKV / 32; ++h) tmem_ld32(s_tmem + 32 * h, *reinterpret_cast<uint32_t(*)[32]>(&r[32 * h]));
        tmem_ld_wait();
        float b0 = -INFINITY, b1 = -INFINITY, b2 = -INFINITY, b3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < kBlockKV; i += 4) {
          b0 = fmaxf(b0, __uint_as_float(r[i]));
          b1 = fmaxf(b1, __uint_as_float(r[i + 1]));
          b2 = fmaxf(b2, __uint_as_float(r[i + 2]));
          b3 = fmaxf(b3, __uint_as_float(r[i + 3]));
        }
        bm = fmaxf(fmaxf(b0, b1), fmaxf(b2, b3));
      } else {
        for (int c0 = 0; c0 < last_n; c0 += 16) {   // sweep 1 of the ragged block: maximum over the valid columns
          uint32_t t[16];
          tmem_ld16_(s_tmem + c0, t);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (c0 + i < last_valid) bm = fmaxf(bm, __uint_as_float(t[i]));
        }
      }

      // running maximum with lazy rescale
      const float m_new = fmaxf(m, bm);
      if (j == 0) {
        m = m_new;
      } else {
        const bool need = (m_new - m) * c > kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {
          // rare: raise m for every row of this warp and rescale its O rows and l.  O is quiescent once P_{j-1} V_{j-1}
          // retired, and P_j V_j cannot be issued before this warp reports p_ready.
          // barrier (j-1) % 3 has completed exactly (j-1)/3 phases once P_{j-4} V_{j-4} retired, which S_j complete implies
          mbar_wait(&pv_done[(j - 1) % kNumSBuf], ((j - 1) / kNumSBuf) & 1);
          tc_fence_after();
          const float alpha = ex2((m - m_new) * c);
          uint32_t o[32];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            tmem_ld32(o_tmem + 32 * h, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st32_(o_tmem + 32 * h, o);
          }
          tmem_st_wait();
          l0 *= alpha; l1 *= alpha; l2 *= alpha; l3 *= alpha;
          m = m_new;
        }
      }
      const float mc = m * c;

      if (!tail && Cfg::kHalves) {
        const float2 c2 = make_float2(c, c), nmc2 = make_float2(-mc, -mc);
        float2 l01 = make_float2(l0, l1), l23 = make_float2(l2, l3);
#pragma unroll
        for (int h = 0; h < kBlockKV / 32; ++h) {
          tmem_ld32(s_tmem + 32 * h, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float2 x01 = ffma2(make_float2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), c2, nmc2);
            const float2 x23 = ffma2(make_float2(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])), c2, nmc2);
            const float2 p01 = make_float2(ex2(x01.x), ex2(x01.y));
            const float2 p23 = make_float2(ex2(x23.x), ex2(x23.y));
            l01 = fadd2(l01, p01);
            l23 = fadd2(l23, p23);
            r[i >> 1] = pack_half2(p01.x, p01.y);
            r[(i >> 1) + 1] = pack_half2(p23.x, p23.y);
          }
          // P of this half lands on S columns already consumed: [16h, 16h + 16) of the buffer
          tmem_st16(s_tmem + 16 * h, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
        }
        l0 = l01.x; l1 = l01.y; l2 = l23.x; l3 = l23.y;
      } else if (!tail) {
        const float2 c2 = make_float2(c, c), nmc2 = make_float2(-mc, -mc);
        float2 l01 = make_float2(l0, l1), l23 = make_float2(l2, l3);
#pragma unroll
        for (int i = 0; i < kBlockKV; i += 4) {   // FFMA2 / FADD2: two columns per instruction; P packed in place
          const float2 x01 = ffma2(make_float2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), c2, nmc2);
          const float2 x23 = ffma2(make_float2(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])), c2, nmc2);
          float2 p01, p23;
          if (Cfg::kPolyStride > 0 && ((i >> 2) % (Cfg::kPolyStride > 0 ? Cfg::kPolyStride : 1)) == (Cfg::kPolyStride - 1)) {
            p01 = exp2_poly(x01);
            p23 = exp2_poly(x23);
          } else {
            p01 = make_float2(ex2(x01.x), ex2(x01.y));
            p23 = make_float2(ex2(x23.x), ex2(x23.y));
          }
          l01 = fadd2(l01, p01);
          l23 = fadd2(l23, p23);
          r[i >> 1] = pack_half2(p01.x, p01.y);
          r[(i >> 1) + 1] = pack_half2(p23.x, p23.y);
        }
        l0 = l01.x; l1 = l01.y; l2 = l23.x; l3 = l23.y;
        // P overwrites S columns already held in registers by this thread
        if constexpr (kBlockKV == 64) tmem_st32_(s_tmem, r);
        else tmem_st16(s_tmem, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
      } else {
        for (int c0 = 0; c0 < last_n; c0 += 16) {   // sweep 2: reload the chunk (P of earlier chunks never reaches it)
          uint32_t t[16], pk[8];
          tmem_ld16_(s_tmem + c0, t);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            const float p0 = (c0 + i < last_valid) ? ex2(fmaf(__uint_as_float(t[i]), c, -mc)) : 0.f;
            const float p1 = (c0 + i + 1 < last_valid) ? ex2(fmaf(__uint_as_float(t[i + 1]), c, -mc)) : 0.f;
            l0 += p0; l1 += p1;
            pk[i >> 1] = pack_half2(p0, p1);
          }
          tmem_st8_(s_tmem + (c0 >> 1), pk);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[j % kNumSBuf]);
    }

    // epilogue: O / l -> fp16 -> global
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float l_sum = (l0 + l1) + (l2 + l3);
    const float inv_l = 1.0f / l_sum;
    if (args.lse2 != nullptr && q_row < S)
      args.lse2[((size_t)view * (args.hidden / kHeadDim) + head) * S + q_row] = m * c + log2f(l_sum);
    __half* orow = args.out + (size_t)(row0 + q_row) * args.hidden + q_col;
#pragma unroll
    for (int c0 = 0; c0 < kHeadDim; c0 += 32) {
      uint32_t o[32];
      tmem_ld32(o_tmem + c0, o);
      tmem_ld_wait();
      if (q_row < S) {
        uint4* o4 = reinterpret_cast<uint4*>(orow + c0);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 v;
          v.x = pack_half2(__uint_as_float(o[8 * i + 0]) * inv_l, __uint_as_float(o[8 * i + 1]) * inv_l);
          v.y = pack_half2(__uint_as_float(o[8 * i + 2]) * inv_l, __uint_as_float(o[8 * i + 3]) * inv_l);
          v.z = pack_half2(__uint_as_float(o[8 * i + 4]) * inv_l, __uint_as_float(o[8 * i + 5]) * inv_l);
          v.w = pack_half2(__uint_as_float(o[8 * i + 6]) * inv_l, __uint_as_float(o[8 * i + 7]) * inv_l);
          o4[i] = v;
        }
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <class Cfg>
int launch_attention(const void* qkv, void* out, int n_views, int seq, int heads, cudaStream_t stream, float* lse2) {
  const int hidden = heads * kHeadDim;
  CUtensorMap tm;
  if (make_tmap_f16_2d(&tm, qkv, (uint64_t)n_views * seq, 3 * hidden, 3 * hidden, Cfg::kBlockKV, kHeadDim)) return 1;
  auto kern = attention_kernel<Cfg>;
  static std::atomic<int> attr_set_dev[64];   // per device (the attribute is per context)
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!attr_set_dev[dev].load(std::memory_order_acquire)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) { set_last_error("attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 1; }
    attr_set_dev[dev].store(1, std::memory_order_release);
  }
  AttnArgs a;
  a.seq = seq;
  a.hidden = hidden;
  a.out = reinterpret_cast<__half*>(out);
  a.scale_log2 = 0.125f * 1.4426950408889634f;
  a.lse2 = lse2;
  dim3 grid((seq + kBlockQ - 1) / kBlockQ, heads, n_views);
  ProfScope prof("attention", stream);
  kern<<<grid, kThreads, Cfg::kSmemBytes, stream>>>(tm, a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_last_error("attention launch: %s", cudaGetErrorString(e)); return 1; }
  return 0;
}

}  // namespace

// Variant switches are read ONCE per process (PG_ATTN_VARIANT: unset = fold kernel for inference and pair kernel when the
// log-sum-exp side output is requested | "fold" | "pair" | "split" | "legacy" | "64" | "64s" | "64h" | "32c";
// PG_ATTN_POLY: eighths of the exponentials on the FMA pipe for the fold / pair / split kernels, or 4 / 2 = every 4th / 2nd
// group for the legacy kernel).
struct AttnSwitches {
  int variant = -1;  // -1 default, 0 pair, 1 legacy 32/2/4, 2 "64", 3 "64s", 4 "64h", 5 "32c", 6 split, 7 fold
  int poly = -1;
  AttnSwitches() {
    const char* v = getenv("PG_ATTN_VARIANT");
    if (v) {
      if (!strcmp(v, "pair")) variant = 0;
      else if (!strcmp(v, "legacy") || !strcmp(v, "32")) variant = 1;
      else if (!strcmp(v, "split")) variant = 6;
      else if (!strcmp(v, "fold")) variant = 7;
      else if (!strcmp(v, "64")) variant = 2;
      else if (!strcmp(v, "64s")) variant = 3;
      else if (!strcmp(v, "64h")) variant = 4;
      else if (!strcmp(v, "32c")) variant = 5;
    }
    const char* pe = getenv("PG_ATTN_POLY");
    if (pe && pe[0] >= '0' && pe[0] <= '9') poly = pe[0] - '0';
  }
};

int attention_f16_variant(const void* qkv, void* out, int n_views, int seq, int heads, cudaStream_t stream, float* lse2,
                          int variant, int poly) {
  if (n_views <= 0) return 0;
  if (variant == 0) return attention_pair_f16(qkv, out, n_views, seq, heads, stream, lse2, poly < 0 ? 2 : poly);
  if (variant == 2) return attention_split_f16(qkv, out, n_views, seq, heads, stream, lse2, poly < 0 ? 2 : poly);
  if (variant == 3) return attention_fold_f16(qkv, out, n_views, seq, heads, stream, lse2, poly < 0 ? 2 : poly);
  if (n_views > 65535) { set_last_error("attention (legacy kernel): n_views %d > 65535 (grid.z)", n_views); return 1; }
  if (poly == 4) return launch_attention<AttnCfg<32, 2, 4, 4>>(qkv, out, n_views, seq, heads, stream, lse2);
  if (poly == 2) return launch_attention<AttnCfg<32, 2, 4, 2>>(qkv, out, n_views, seq, heads, stream, lse2);
  return launch_attention<AttnCfg<32, 2, 4>>(qkv, out, n_views, seq, heads, stream, lse2);
}

int attention_f16(const void* qkv, void* out, int n_views, int seq, int heads, cudaStream_t stream, float* lse2) {
  if (n_views <= 0) return 0;
  static const AttnSwitches sw;
  if (sw.variant < 0) {
    // Training keeps the pair kernel: its lse2 is the log-sum-exp of exactly the fp16 q.k products the backward recomputes
    // (the fold kernel rounds c*q to fp16 once more, 1e-3 in log2 units).
    if (lse2 == nullptr) return attention_fold_f16(qkv, out, n_views, seq, heads, stream, nullptr, sw.poly < 0 ? 2 : sw.poly);
    return attention_pair_f16(qkv, out, n_views, seq, heads, stream, lse2, sw.poly < 0 ? 2 : sw.poly);
  }
  if (sw.variant != 0 && sw.variant != 7 && n_views > 65535) {
    set_last_error("attention (legacy kernel): n_views %d > 65535 (grid.z)", n_views);
    return 1;
  }
  switch (sw.variant) {
    case 0: return attention_pair_f16(qkv, out, n_views, seq, heads, stream, lse2, sw.poly < 0 ? 2 : sw.poly);
    case 6: return attention_split_f16(qkv, out, n_views, seq, heads, stream, lse2, sw.poly < 0 ? 2 : sw.poly);
    case 7: return attention_fold_f16(qkv, out, n_views, seq, heads, stream, lse2, sw.poly < 0 ? 2 : sw.poly);
    case 2: return launch_attention<AttnCfg<64, 3, 2>>(qkv, out, n_views, seq, heads, stream, lse2);
    case 3: return launch_attention<AttnCfg<64, 1, 4, 0, 4>>(qkv, out, n_views, seq, heads, stream, lse2);
    case 4: return launch_attention<AttnCfg<64, 1, 4, 0, 4, true>>(qkv, out, n_views, seq, heads, stream, lse2);
    case 5: return launch_attention<AttnCfg<32, 2, 3>>(qkv, out, n_views, seq, heads, stream, lse2);
    default: break;
  }
  if (sw.poly == 4) return launch_attention<AttnCfg<32, 2, 4, 4>>(qkv, out, n_views, seq, heads, stream, lse2);
  if (sw.poly == 2) return launch_attention<AttnCfg<32, 2, 4, 2>>(qkv, out, n_views, seq, heads, stream, lse2);
  return launch_attention<AttnCfg<32, 2, 4>>(qkv, out, n_views, seq, heads, stream, lse2);
}

}  // namespace pg
