// Multi-head self-attention core for the CLIP vision tower on sm_100a, second generation ("pair" kernel).
//
//   out[v, s, h*64 + :] = softmax_s'( q[v,s,h,:] . k[v,s',h,:] / sqrt(64) ) @ v[v,s',h,:]
//
// Restates the attention core of HF CLIPAttention.forward (bmm -> fp32 softmax -> bmm; no mask, no dropout in eval)
// that the reference reaches through models/clip_embedder.py:63 and models/super_guessr.py:395.  head_dim = 64.
//
// Why a second kernel: the first one (attention_tcgen05.cu) issued S = Q K^T as 128 x 32 x 16 MMAs with both operands in
// shared memory (320 B/clk of operand reads against the 128 B/clk the SM delivers), synchronised softmax and MMA warps every
// 32 columns and was MUFU-bound at one ex2 per logit; ncu put it at 23 % tensor pipe.  This one:
//   * persistent, ONE CTA per SM, 512 threads; a CTA works on "jobs" of TWO 128-row query tiles (slots 0 / 1) that ping-pong:
//     while the softmax warps of one slot work, the tensor pipe serves the other.  Two tiles of the same (view, head) share one
//     K/V stream; the odd last tiles (577 = 4 * 128 + 65) of two neighbouring heads are paired with separate streams.
//   * KV blocks of 64 with TWO S buffers per tile: S_i[b] = Q_i K_j^T is 4 MMAs of 128 x 64 x 16, P_i V_j 4 MMAs of
//     128 x 64 x 16; the MMA warp runs two blocks ahead, so the softmax warps of a tile find their next S block complete and
//     never wait for the tensor round trip.  Q lives in TENSOR MEMORY (A operand): every MMA reads only its B operand (K or
//     V, 64 B/clk) from shared memory.
//   * softmax threads (one per query row, 2 x 4 warps) process 64 logits per barrier round trip; the row maximum is computed
//     for the first block only: later blocks reuse it and check the block sum instead (P = 2^(x - m) stays below 2^15, exact
//     in fp16 range; a larger sum triggers the exact path: true maximum, O and l rescaled in TMEM, block redone).
//   * a tunable share of the exponentials runs on the FMA pipe (Cody-Waite + degree-3 minimax, 7.5e-5 relative — below the
//     fp16 rounding of P) to relieve the 16 / clk / SM MUFU unit.
//   * Q tiles arrive by TMA into a staging buffer one job ahead and are copied to TMEM while the previous job finishes; the
//     epilogue (O / l -> fp16 -> global) runs on its own 4 warps, so consecutive jobs overlap.
//
// TMEM columns: tile 0 S buffers [0,64) [64,128), tile 1 [128,192) [192,256) (P aliases the first 32 columns of its S buffer
// as packed fp16), O0 [256,320) O1 [320,384), Q0 [384,416) Q1 [416,448).
// Warps: 0-3 softmax tile 0, 4-7 softmax tile 1, 8-11 epilogue, 12 TMA producer, 13 / 14 MMA issuers of tile 0 / 1 (13 also
// allocates TMEM), 15 idle.
#include "attention.h"
#include "prof.h"
#include "ptx.cuh"
#include "tma_host.h"

#include <atomic>
#include <type_traits>

namespace pg {

namespace {

constexpr int kHeadDim = 64;
constexpr int kBlock = 128;                            // query rows per tile
constexpr int kSub = 64;                               // kv rows per block
constexpr int kSubBytes = kSub * kHeadDim * 2;         // 8 KB: one K or V block (a Q tile is two of them)
constexpr int kQBytes = 2 * kSubBytes;
constexpr int kSlots = 8;                              // K/V ring of 16 KB slots: {K_0, K_1} or {V_j, K_(j+2)} of one tile stream
constexpr int kSlotBytes = 2 * kSubBytes;
constexpr int kThreads = 512;                          // 16 warps: setmaxnreg is a per-warpgroup (4 warps) operation
constexpr int kWarpEpi = 8, kWarpTma = 12, kWarpMma = 13;   // MMA issuers: warp 13 (tile 0), warp 14 (tile 1)
constexpr uint32_t kColS = 0, kColO = 256, kColQ = 384;
constexpr int kTmemCols = 512;
constexpr float kRescaleThreshold = 8.0f;              // log2 domain, ragged last block (exact maximum, lazy rescale)
constexpr float kSumLimit = 32768.f;                   // block sum that triggers the exact-maximum path (P < 2^15)

struct Bars {
  uint64_t kv_full[kSlots], kv_empty[kSlots];
  uint64_t qs_full[2], qs_empty[2];   // TMA -> softmax (Q staging tile landed), softmax -> TMA (copied to TMEM)
  uint64_t q_ready[2];                // softmax -> MMA : Q_i is in TMEM
  uint64_t s_full[2][2];              // MMA -> softmax : S block complete in buffer [tile][b]
  uint64_t p_ready[2][2];             // softmax -> MMA : P written over buffer [tile][b]
  uint64_t pv_done[2];                // MMA -> softmax : P_i V retired (O_i quiescent), one phase per block
  uint64_t o_full[2];                 // MMA -> epilogue: O_i complete
  uint64_t o_free[2];                 // epilogue -> MMA / softmax : O_i and the row statistics were read
  uint64_t l_ready[2];                // softmax -> epilogue : row statistics published
  uint32_t tmem_ptr;
};
constexpr int kRingBytes = kSlots * kSlotBytes;
constexpr int kSmemBytes = 1024 + kRingBytes + 2 * kQBytes + 1024 /* Bars */ + 2 * 2 * kBlock * 4;

struct PairArgs {
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

__device__ __forceinline__ Job decode_job(int job, const PairArgs& a) {
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

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float y;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c));
  return y;
}

// 2^x for a pair on the FMA / ALU pipes, x <= ~16: n = round(x) via the 1.5 * 2^23 magic constant, f = x - n in [-0.5, 0.5],
// 2^f by a degree-3 minimax polynomial (max relative error 7.5e-5), 2^n by an exponent-field add.
__device__ __forceinline__ float2 exp2_poly3(float2 x) {
  x.x = fmaxf(x.x, -125.f);
  x.y = fmaxf(x.y, -125.f);
  const float2 magic = make_float2(12582912.f, 12582912.f);
  const float2 t = fadd2(x, magic);
  const float2 n = fsub2(t, magic);
  const float2 f = fsub2(x, n);
  float2 p = ffma2(make_float2(0.0551716685295105f, 0.0551716685295105f), f,
                   make_float2(0.2426111400127411f, 0.2426111400127411f));
  p = ffma2(p, f, make_float2(0.6932609677314758f, 0.6932609677314758f));
  p = ffma2(p, f, make_float2(0.9999280571937561f, 0.9999280571937561f));
  p.x = __int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23));
  p.y = __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23));
  return p;
}

__device__ __forceinline__ void tmem_ld16p(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// tcgen05.wait::ld that also "touches" the 16 destination registers, so that the compiler cannot move their uses above it
// when loads are software-pipelined (the next load is issued between this wait and the arithmetic on r).
__device__ __forceinline__ void tmem_ld_wait16(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_st8p(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st16p(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st32p(uint32_t taddr, const uint32_t* r) {
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
__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}

// One 16-column chunk of a row: logits r[16] -> P as 8 packed fp16 pairs, row-sum contributions into acc0 / acc1.
// POLY: bit k set -> pair k of the chunk is exponentiated on the FMA pipe.
template <int POLY>
__device__ __forceinline__ void exp_chunk(const uint32_t* r, uint32_t* pk, float2 c2, float2 nmc2, float2& acc0,
                                          float2& acc1) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float2 x = ffma2(make_float2(__uint_as_float(r[2 * k]), __uint_as_float(r[2 * k + 1])), c2, nmc2);
    float2 p;
    if ((POLY >> k) & 1) p = exp2_poly3(x);
    else p = make_float2(ex2(x.x), ex2(x.y));
    if (k & 1) acc1 = fadd2(acc1, p);
    else acc0 = fadd2(acc0, p);
    pk[k] = pack_half2(p.x, p.y);
  }
}

// Ragged chunk: only the first `rem` of the 16 columns are valid keys; P = 0 for the others.
__device__ __forceinline__ void exp_chunk_masked(const uint32_t* r, uint32_t* pk, int rem, float c, float nmc, float& acc) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float p0 = (2 * k < rem) ? ex2(fmaf(__uint_as_float(r[2 * k]), c, nmc)) : 0.f;
    const float p1 = (2 * k + 1 < rem) ? ex2(fmaf(__uint_as_float(r[2 * k + 1]), c, nmc)) : 0.f;
    acc += p0 + p1;
    pk[k] = pack_half2(p0, p1);
  }
}

// One lane of a converged warp (the single-thread roles run warp-uniform so that descriptors and addresses stay in uniform
// registers; only the tcgen05 / TMA instructions themselves are issued under this predicate).
__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(p));
  return p != 0;
}

template <int I>
using Slot = std::integral_constant<int, I>;

template <int POLY>
__global__ void __launch_bounds__(kThreads, 1)
attention_pair_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const PairArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_kv = smem;
  uint8_t* smem_q = smem + kRingBytes;
  Bars* bars = reinterpret_cast<Bars*>(smem + kRingBytes + 2 * kQBytes);
  float* lm = reinterpret_cast<float*>(smem + kRingBytes + 2 * kQBytes + 1024);   // [slot][l | m][row]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = args.seq;
  const int nb = (S + kSub - 1) / kSub;                        // KV blocks (10 for S = 577)
  const int last_valid = S - (nb - 1) * kSub;                  // valid kv columns in the last block (1)
  const int last_n = (last_valid + 15) & ~15;                  // MMA N / K extent of the last block (16)
  const int n_jobs = args.n_jobs;
  const int stride = gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    for (int s = 0; s < kSlots; ++s) {
      mbar_init(&bars->kv_full[s], 1);
      mbar_init(&bars->kv_empty[s], 2);   // two tcgen05.commit arrivals: one per MMA warp (shared stream) or both from the owner
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->qs_full[i], 1);
      mbar_init(&bars->qs_empty[i], 4);
      mbar_init(&bars->q_ready[i], 4);
      for (int b = 0; b < 2; ++b) {
        mbar_init(&bars->s_full[i][b], 1);
        mbar_init(&bars->p_ready[i][b], 4);
      }
      mbar_init(&bars->pv_done[i], 1);
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
  // warpgroup: warps 12-15 (TMA, MMA, two idle) form one.
  // K/V ring protocol.  A job's stream of a tile is 1 + nb slots: {K_0, K_1}, then {V_j, K_(j+2)} for j = 0 .. nb-1.  Two tiles
  // of one (view, head) share ONE stream (both MMA warps read every slot and commit once each); otherwise the two streams are
  // interleaved slot by slot (the owner commits twice), or there is a single stream (tile 1 absent).  Every role derives the
  // same ring positions from the job list.
  if (warp < kWarpEpi) {
    // ---------------------------------------------------------------- softmax warps: thread = query row = TMEM lane
    asm volatile("setmaxnreg.inc.sync.aligned.u32 168;");
    const int i = warp >> 2;        // tile slot
    const int wq = warp & 3;        // lane quarter
    const int row = wq * 32 + lane;
    const uint32_t lane_base = uint32_t(wq * 32) << 16;
    const uint32_t s_tmem0 = tmem_base + lane_base + kColS + 128 * i;
    const uint32_t o_tmem = tmem_base + lane_base + kColO + 64 * i;
    const uint32_t q_tmem = tmem_base + lane_base + kColQ + 32 * i;
    const uint32_t q_smem = smem_u32(smem_q + i * kQBytes) + row * 128;
    const float c = args.scale_log2;
    uint32_t n_q = 0, n_blk = 0, n_job = 0;   // n_blk: running block count of this tile (block g lives in S buffer g & 1)
    bool q_done = false;

    // staged Q tile (128-byte swizzle: 16-byte chunk ch of row r sits at chunk ch ^ (r & 7)) -> TMEM A operand
    auto copy_q = [&]() {
      mbar_wait(&bars->qs_full[i], n_q & 1);
      ++n_q;
      uint32_t qr[32];
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        const uint4 v = lds128(q_smem + ((ch ^ (row & 7)) << 4));
        qr[4 * ch] = v.x; qr[4 * ch + 1] = v.y; qr[4 * ch + 2] = v.z; qr[4 * ch + 3] = v.w;
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

      float m = -INFINITY;   // reference maximum used in the exponent (raw logit units)
      float l = 0.f;         // running row sum of P

      // rare: raise the reference maximum to m_new, rescale this row's O and l.  O_i is quiescent once P_i V of the previous
      // block retired, and the next P_i V cannot be issued before this warp reports p_ready.
      auto rescale = [&](float m_new) {
        mbar_wait(&bars->pv_done[i], (n_blk - 1) & 1);
        tc_fence_after();
        const float alpha = ex2((m - m_new) * c);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t o[32];
          tmem_ld32(o_tmem + 32 * h, o);
          tmem_ld_wait();
#pragma unroll
          for (int x = 0; x < 32; ++x) o[x] = __float_as_uint(__uint_as_float(o[x]) * alpha);
          tmem_st32p(o_tmem + 32 * h, o);
        }
        tmem_st_wait();
        l *= alpha;
      };

      for (int j = 0; j < nb; ++j) {
        const int b = n_blk & 1;
        const uint32_t s_tmem = s_tmem0 + kSub * b;
        mbar_wait(&bars->s_full[i][b], (n_blk >> 1) & 1);
        tc_fence_after();
        if (j == nb - 1 && next_active) {   // every S_i MMA of this job has retired: Q_i may be replaced
          copy_q();
          q_done = true;
        }
        if (!warp_active) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->p_ready[i][b]);
          ++n_blk;
          continue;
        }
        if (j < nb - 1 || last_valid == kSub) {
          // ---- a block of 64 valid key columns: one branch-free basic block, P held in registers until the sum is checked
          uint32_t pk[32];
          float bs;
          bool exact = false;   // m is known to be >= every logit of this block
          if (j == 0) {
            // reference maximum of a job: of the first 16 logits only (any reference works as long as the block sums stay
            // below kSumLimit; the class token's key is column 0)
            uint32_t r[16];
            tmem_ld16p(s_tmem, r);
            tmem_ld_wait16(r);
            const float b0 = fmax3(fmax3(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2])),
                                   __uint_as_float(r[3]), __uint_as_float(r[4]));
            const float b1 = fmax3(fmax3(__uint_as_float(r[5]), __uint_as_float(r[6]), __uint_as_float(r[7])),
                                   __uint_as_float(r[8]), __uint_as_float(r[9]));
            const float b2 = fmax3(fmax3(__uint_as_float(r[10]), __uint_as_float(r[11]), __uint_as_float(r[12])),
                                   __uint_as_float(r[13]), __uint_as_float(r[14]));
            m = fmax3(b0, b1, fmaxf(b2, __uint_as_float(r[15])));
          }
          for (;;) {
            const float mc = m * c;
            const float2 c2 = make_float2(c, c), nmc2 = make_float2(-mc, -mc);
            float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
            uint32_t ra[16], rb[16];
            tmem_ld16p(s_tmem, ra);
#pragma unroll
            for (int ch = 0; ch < 4; ch += 2) {   // software-pipelined: the next chunk's load is in flight during the arithmetic
              tmem_ld_wait16(ra);
              tmem_ld16p(s_tmem + 16 * (ch + 1), rb);
              exp_chunk<POLY>(ra, pk + 8 * ch, c2, nmc2, acc0, acc1);
              tmem_ld_wait16(rb);
              if (ch + 2 < 4) tmem_ld16p(s_tmem + 16 * (ch + 2), ra);
              exp_chunk<POLY>(rb, pk + 8 * (ch + 1), c2, nmc2, acc0, acc1);
            }
            bs = (acc0.x + acc0.y) + (acc1.x + acc1.y);
            if (!exact && __any_sync(0xffffffffu, !(bs < kSumLimit))) {
              // rare: some P may not fit fp16.  Take the exact maximum of the block, rescale O and l, redo the block.
              float b0 = -INFINITY, b1 = -INFINITY;
              for (int h = 0; h < 4; ++h) {
                uint32_t r[16];
                tmem_ld16p(s_tmem + 16 * h, r);
                tmem_ld_wait16(r);
#pragma unroll
                for (int x = 0; x < 16; x += 4) {
                  b0 = fmax3(b0, __uint_as_float(r[x]), __uint_as_float(r[x + 1]));
                  b1 = fmax3(b1, __uint_as_float(r[x + 2]), __uint_as_float(r[x + 3]));
                }
              }
              const float m_new = fmax3(m, b0, b1);
              if (j > 0) rescale(m_new);
              m = m_new;
              exact = true;
              continue;
            }
            break;
          }
          l += bs;
          // P (packed fp16) over the first 32 columns of the S buffer: every logit of the block is already in registers
          tmem_st32p(s_tmem, pk);
        } else {
          // ---- ragged last block: nfull whole 16-column chunks + `rem` valid columns of one more.  Exact maximum first
          // (lazy rescale: P <= 2^8), then P chunk by chunk (P of chunk ch lands on S columns of chunks <= ch, already read).
          const int nfull = last_valid >> 4, rem = last_valid & 15;
          float b0 = -INFINITY, b1 = -INFINITY, b2 = -INFINITY, b3 = -INFINITY;
          for (int ch = 0; ch < nfull; ++ch) {
            uint32_t r[16];
            tmem_ld16p(s_tmem + 16 * ch, r);
            tmem_ld_wait16(r);
            b0 = fmax3(b0, __uint_as_float(r[0]), __uint_as_float(r[1]));
            b1 = fmax3(b1, __uint_as_float(r[2]), __uint_as_float(r[3]));
            b2 = fmax3(b2, __uint_as_float(r[4]), __uint_as_float(r[5]));
            b3 = fmax3(b3, __uint_as_float(r[6]), __uint_as_float(r[7]));
            b0 = fmax3(b0, __uint_as_float(r[8]), __uint_as_float(r[9]));
            b1 = fmax3(b1, __uint_as_float(r[10]), __uint_as_float(r[11]));
            b2 = fmax3(b2, __uint_as_float(r[12]), __uint_as_float(r[13]));
            b3 = fmax3(b3, __uint_as_float(r[14]), __uint_as_float(r[15]));
          }
          if (rem) {
            uint32_t r[16];
            tmem_ld16p(s_tmem + 16 * nfull, r);
            tmem_ld_wait16(r);
#pragma unroll
            for (int x = 0; x < 16; ++x)
              if (x < rem) b0 = fmaxf(b0, __uint_as_float(r[x]));
          }
          const float m_new = fmaxf(fmaxf(m, fmaxf(b0, b1)), fmaxf(b2, b3));
          if (j == 0) {
            m = m_new;
          } else if (__any_sync(0xffffffffu, (m_new - m) * c > kRescaleThreshold)) {
            rescale(m_new);
            m = m_new;
          }
          const float mc = m * c;
          const float2 c2 = make_float2(c, c), nmc2 = make_float2(-mc, -mc);
          float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
          for (int ch = 0; ch < nfull; ++ch) {
            uint32_t r[16], pk8[8];
            tmem_ld16p(s_tmem + 16 * ch, r);
            tmem_ld_wait16(r);
            exp_chunk<0>(r, pk8, c2, nmc2, acc0, acc1);
            tmem_st8p(s_tmem + 8 * ch, pk8);
          }
          float bs = (acc0.x + acc0.y) + (acc1.x + acc1.y);
          if (rem) {
            uint32_t r[16], pk8[8];
            tmem_ld16p(s_tmem + 16 * nfull, r);
            tmem_ld_wait16(r);
            exp_chunk_masked(r, pk8, rem, c, -mc, bs);
            tmem_st8p(s_tmem + 8 * nfull, pk8);
          }
          l += bs;
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->p_ready[i][b]);
        ++n_blk;
      }

      // publish the row statistics; the epilogue of this slot's previous job must have read its own first
      if (n_job > 0) mbar_wait(&bars->o_free[i], (n_job - 1) & 1);
      lm[i * 2 * kBlock + row] = l;
      lm[i * 2 * kBlock + kBlock + row] = m;
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->l_ready[i]);
      ++n_job;
    }
  } else if (warp < kWarpTma) {
    // ---------------------------------------------------------------- epilogue warps: O / l -> fp16 -> global
    asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
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
        const float l_sum = lm[i * 2 * kBlock + row];
        const float m = lm[i * 2 * kBlock + kBlock + row];
        const float inv_l = 1.0f / l_sum;
        const uint32_t o_tmem = tmem_base + lane_base + kColO + 64 * i;
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
          uint4* o4 = reinterpret_cast<uint4*>(args.out + grow * args.hidden + h * kHeadDim);
#pragma unroll
          for (int x = 0; x < 8; ++x) o4[x] = make_uint4(pk[4 * x], pk[4 * x + 1], pk[4 * x + 2], pk[4 * x + 3]);
          if (args.lse2 != nullptr)
            args.lse2[((size_t)jb.view * heads + h) * S + token] = m * args.scale_log2 + log2f(l_sum);
        }
      }
    }
  // (every warpgroup's setmaxnreg sits inside its own role branch, and no role calls a non-inlined function: ptxas compiles a
  // shared callee for the smallest budget and then holds every caller to it)
  } else {
   asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
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
    bool pre = false;                                    // S of this job's first two blocks was issued during the previous job
    // smem descriptors of a slot's first (offset 0) / second (offset 8 KB) block: K is K-major (rows of 128 B, k-step = 32 B),
    // V is MN-major (row = kv index, k-step = 16 rows)
    const uint64_t k_desc0 = make_smem_desc(smem_u32(smem_kv), 16, 1024, kLayoutSw128);
    const uint64_t v_desc0 = make_smem_desc(smem_u32(smem_kv), 1024, 1024, kLayoutSw128);
    const uint32_t idesc_s = make_idesc_f16(kBlock, kSub, 0, 0);
    const uint32_t idesc_s_last = make_idesc_f16(kBlock, last_n, 0, 0);
    const uint32_t idesc_pv = make_idesc_f16(kBlock, kHeadDim, 0, 1);    // B (= V) is MN-major
    const uint32_t s_base = tmem_base + kColS + 128 * I, o_base = tmem_base + kColO + 64 * I, q_base = tmem_base + kColQ + 32 * I;
    // S[buf] = Q_I K_j^T, j = block index inside its job   (inside an elected region)
    auto issue_s = [&](uint64_t kd, int j, uint32_t buf) {
      const uint32_t idesc = (j == nb - 1) ? idesc_s_last : idesc_s;
      const uint32_t d = s_base + kSub * buf;
#pragma unroll
      for (int k = 0; k < kHeadDim / 16; ++k) umma_ts(d, q_base + 8 * k, kd + 2 * k, idesc, k != 0);
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
      // the next job of this tile: its first S blocks are issued while this job's last two blocks are in the softmax warps
      bool has_next = (job + stride < n_jobs) && nb >= 2;
      uint32_t next_p = 0;
      int next_commit = 0;
      if (has_next) {
        const Job jn = decode_job(job + stride, args);
        if (I == 1 && !jn.a1) has_next = false;
        const int ns = (jn.shared || !jn.a1) ? 1 : 2;
        next_p = pos + ((ns == 2) ? I : 0);
        next_commit = jn.shared ? 1 : 2;
      }
      for (int step = pre ? 1 : 0; step <= nb; ++step) {
        const uint32_t p = pos_job + first + (uint32_t)step * nstream;
        const int slot = p % kSlots;
        const int j = step - 1;
        if (step == 0) {
          mbar_wait(&bars->q_ready[I], n_j & 1);
        } else {
          const uint32_t gb = g + j;
          mbar_wait(&bars->p_ready[I][gb & 1], (gb >> 1) & 1);
          if (j == 0) mbar_wait(&bars->o_free[I], (n_j & 1) ^ 1);   // the previous job's O_I was read out
        }
        mbar_wait(&bars->kv_full[slot], (p / kSlots) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t d0 = (uint64_t)(slot * (kSlotBytes >> 4)), d1 = d0 + (kSubBytes >> 4);
          if (step == 0) {
            issue_s(k_desc0 + d0, 0, g & 1);
            if (nb > 1) issue_s(k_desc0 + d1, 1, (g + 1) & 1);
          } else {
            const uint64_t vd = v_desc0 + d0;
            const uint32_t a = s_base + kSub * ((g + j) & 1);       // P (packed fp16 over the S buffer)
            umma_ts(o_base, a, vd, idesc_pv, j != 0);
            if (j == nb - 1) {
              for (int k = 1; k < last_n / 16; ++k) umma_ts(o_base, a + 8 * k, vd + 128 * k, idesc_pv, 1);
            } else {
#pragma unroll
              for (int k = 1; k < kSub / 16; ++k) umma_ts(o_base, a + 8 * k, vd + 128 * k, idesc_pv, 1);
            }
            tc_commit(&bars->pv_done[I]);
            if (j == nb - 1) tc_commit(&bars->o_full[I]);
            if (j + 2 < nb) issue_s(k_desc0 + d1, j + 2, (g + j) & 1);
          }
          tc_commit(&bars->kv_empty[slot]);
          if (ncommit == 2) tc_commit(&bars->kv_empty[slot]);
        }
        __syncwarp();
        if (has_next && j >= nb - 2) {
          // block j's buffer is free again (its P V is issued): the next job's block j - (nb - 2) goes there.  Q of the next job
          // reaches TMEM when the softmax warps start this job's last block.
          const int slot_n = next_p % kSlots;
          if (j == nb - 2) {
            mbar_wait(&bars->q_ready[I], (n_j + 1) & 1);
            mbar_wait(&bars->kv_full[slot_n], (next_p / kSlots) & 1);
            tc_fence_after();
          }
          if (elect_one()) {
            const uint64_t d0 = (uint64_t)(slot_n * (kSlotBytes >> 4));
            if (j == nb - 2) {
              issue_s(k_desc0 + d0, 0, (g + nb) & 1);
            } else {
              issue_s(k_desc0 + d0 + (kSubBytes >> 4), 1, (g + nb + 1) & 1);
              tc_commit(&bars->kv_empty[slot_n]);
              if (next_commit == 2) tc_commit(&bars->kv_empty[slot_n]);
            }
          }
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

std::atomic<int> g_attr_set[64][5];   // per device and kernel variant: dynamic shared memory opt-in done

template <int POLY>
int launch_pair(int variant, const void* qkv, void* out, int n_views, int seq, int heads, cudaStream_t stream,
                float* lse2) {
  const int hidden = heads * kHeadDim;
  CUtensorMap tm;
  if (make_tmap_f16_2d(&tm, qkv, (uint64_t)n_views * seq, 3 * hidden, 3 * hidden, kSub, kHeadDim)) return 1;
  auto kern = attention_pair_kernel<POLY>;
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
  PairArgs a;
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
  kern<<<grid, kThreads, kSmemBytes, stream>>>(tm, a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_last_error("attention launch: %s", cudaGetErrorString(e)); return 1; }
  return 0;
}

}  // namespace

// poly: share of the exponentials evaluated on the FMA pipe, in eighths (0, 1, 2, 3 or 4 of every 8 pairs; other values = 2).
int attention_pair_f16(const void* qkv, void* out, int n_views, int seq, int heads, cudaStream_t stream, float* lse2,
                       int poly) {
  if (n_views <= 0) return 0;
  switch (poly) {
    case 0: return launch_pair<0x00>(0, qkv, out, n_views, seq, heads, stream, lse2);
    case 1: return launch_pair<0x08>(1, qkv, out, n_views, seq, heads, stream, lse2);
    case 3: return launch_pair<0x4A>(3, qkv, out, n_views, seq, heads, stream, lse2);
    case 4: return launch_pair<0xAA>(4, qkv, out, n_views, seq, heads, stream, lse2);
    default: return launch_pair<0x88>(2, qkv, out, n_views, seq, heads, stream, lse2);
  }
}

}  // namespace pg
