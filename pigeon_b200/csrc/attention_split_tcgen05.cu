// Multi-head self-attention core for the CLIP vision tower on sm_100a, third generation ("split" kernel).
//
//   out[v, s, h*64 + :] = softmax_s'( q[v,s,h,:] . k[v,s',h,:] / sqrt(64) ) @ v[v,s',h,:]
//
// Restates the attention core of HF CLIPAttention.forward (bmm -> fp32 softmax -> bmm; no mask, no dropout in eval)
// that the reference reaches through models/clip_embedder.py:63 and models/super_guessr.py:395.  head_dim = 64.
//
// Same job structure as attention_pair_tcgen05.cu (persistent, one CTA per SM, two 128-row query tiles per job sharing one
// K/V stream, KV blocks of 64, two S buffers per tile, MMA warps two blocks ahead, next job's first S blocks issued during
// the current job's tail).  The pair kernel ended up bound by the instruction rate of its EIGHT softmax warps (two per
// scheduler: MUFU pipe 69 % busy inside the sweeps, idle during every barrier round trip).  This kernel runs SIXTEEN:
//   * every 64-column S block is split between two warps per 32 query rows: half A = columns [0,32), half B = [32,64).  A
//     (tile, half) is an independent flash-attention stream: its own reference maximum m, row sum l and its OWN accumulator
//     O_half += P_half V_half (K extent 32) in tensor memory — no communication between the two warps of a row, not even in
//     the rare rescale path.  The epilogue merges the halves like split-KV decoding:
//         O = (O_A w_A + O_B w_B) / (l_A w_A + l_B w_B),   w_X = 2^((m_X - max(m_A, m_B)) c).
//   * tensor memory is full with 2 x (2 S buffers + O_A + O_B) = 512 columns, so Q stays in SHARED memory (the TMA staging
//     tile is the A operand itself, double-buffered per tile; no copy through registers).  S = Q K^T then reads 6 KB of
//     operands per 128 x 64 x 16 MMA (48 instead of 32 cycles) — the tensor pipe has the slack.
//
// TMEM columns of tile i (base 256 i): S buffers [0,64) [64,128) (P_A = packed fp16 over columns [0,16) of its buffer,
// P_B over [32,48)), O_A [128,192), O_B [192,256).
// Warps: 0-15 softmax (tile = w >> 3, half = (w >> 2) & 1, lane quarter = w & 3), 16-19 epilogue, 20 TMA producer,
// 21 / 22 MMA issuers of tile 0 / 1 (21 also allocates TMEM), 23 idle.
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
constexpr int kHalf = 32;                              // kv columns of a block handled by one softmax warp per row quarter
constexpr int kSubBytes = kSub * kHeadDim * 2;         // 8 KB: one K or V block (a Q tile is two of them)
constexpr int kQBytes = 2 * kSubBytes;
constexpr int kSlots = 8;                              // K/V ring of 16 KB slots: {K_0, K_1} or {V_j, K_(j+2)} of one tile stream
constexpr int kSlotBytes = 2 * kSubBytes;
constexpr int kThreads = 768;                          // 24 warps: setmaxnreg is a per-warpgroup (4 warps) operation
constexpr int kSoftmaxWarps = 16, kWarpEpi = 16, kWarpTma = 20, kWarpMma = 21;   // MMA issuers: warp 21 (tile 0), 22 (tile 1)
constexpr uint32_t kTileCols = 256, kColO = 128;       // per tile: S buffers at +0 / +64, O_A at +128, O_B at +192
constexpr int kTmemCols = 512;
constexpr float kRescaleThreshold = 8.0f;              // log2 domain, ragged last block (exact maximum, lazy rescale)
constexpr float kSumLimit = 32768.f;                   // block sum that triggers the exact-maximum path (P < 2^15)

struct Bars {
  uint64_t kv_full[kSlots], kv_empty[kSlots];
  uint64_t q_full[2][2], q_empty[2][2];   // TMA <-> MMA: Q staging tile [tile][job parity] landed / no longer read
  uint64_t s_full[2][2];                  // MMA -> softmax : S block complete in buffer [tile][b]
  uint64_t p_ready[2][2][2];              // softmax -> MMA : P written over buffer [tile][b] by half [h]
  uint64_t pv_done[2][2];                 // MMA -> softmax : P V of [tile][half] retired (O quiescent), one phase per block
  uint64_t o_full[2];                     // MMA -> epilogue: both O halves of the tile complete
  uint64_t o_free[2];                     // epilogue -> MMA / softmax : O and the row statistics were read
  uint64_t l_ready[2];                    // softmax -> epilogue : row statistics published (8 warps)
  uint32_t tmem_ptr;
};
constexpr int kRingBytes = kSlots * kSlotBytes;
constexpr int kLmFloats = 2 * 2 * 2 * kBlock;          // [tile][half][l | m][row]
constexpr int kSmemBytes = 1024 + kRingBytes + 4 * kQBytes + 1024 /* Bars */ + kLmFloats * 4;

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
attention_split_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const PairArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_kv = smem;
  uint8_t* smem_q = smem + kRingBytes;                                     // [tile][job parity] 16 KB each
  Bars* bars = reinterpret_cast<Bars*>(smem + kRingBytes + 4 * kQBytes);
  float* lm = reinterpret_cast<float*>(smem + kRingBytes + 4 * kQBytes + 1024);   // [tile][half][l | m][row]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = args.seq;
  const int nb = (S + kSub - 1) / kSub;                        // KV blocks (10 for S = 577)
  const int last_valid = S - (nb - 1) * kSub;                  // valid kv columns in the last block (1)
  const int last_n = (last_valid + 15) & ~15;                  // MMA N extent of the last S block (16)
  const int n_jobs = args.n_jobs;
  const int stride = gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    for (int s = 0; s < kSlots; ++s) {
      mbar_init(&bars->kv_full[s], 1);
      mbar_init(&bars->kv_empty[s], 2);   // two tcgen05.commit arrivals: one per MMA warp (shared stream) or both from the owner
    }
    for (int i = 0; i < 2; ++i) {
      for (int b = 0; b < 2; ++b) {
        mbar_init(&bars->q_full[i][b], 1);
        mbar_init(&bars->q_empty[i][b], 1);
        mbar_init(&bars->s_full[i][b], 1);
        mbar_init(&bars->p_ready[i][b][0], 4);
        mbar_init(&bars->p_ready[i][b][1], 4);
        mbar_init(&bars->pv_done[i][b], 1);
      }
      mbar_init(&bars->o_full[i], 1);
      mbar_init(&bars->o_free[i], 4);
      mbar_init(&bars->l_ready[i], 8);
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

  // register budget: 768 threads start with 80 each (16*96 + 4*40 + 4*56 = 24*80 per lane).  Every warpgroup's setmaxnreg sits
  // inside its own role branch, and no role calls a non-inlined function (ptxas compiles a shared callee for the smallest
  // budget and then holds every caller to it).
  // K/V ring protocol: as in attention_pair_tcgen05.cu — a job's stream of a tile is 1 + nb slots: {K_0, K_1}, then
  // {V_j, K_(j+2)}; two tiles of one (view, head) share ONE stream (each MMA warp commits once), otherwise the two streams are
  // interleaved slot by slot (the owner commits twice).  Every role derives the same ring positions from the job list.
  if (warp < kSoftmaxWarps) {
    // ---------------------------------------------------------------- softmax warps: thread = (query row, column half)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 96;");
    const int i = warp >> 3;          // tile
    const int hh = (warp >> 2) & 1;   // column half of every S block
    const int wq = warp & 3;          // lane quarter
    const int row = wq * 32 + lane;
    const uint32_t lane_base = uint32_t(wq * 32) << 16;
    const uint32_t s_tmem0 = tmem_base + lane_base + kTileCols * i + kHalf * hh;       // this half's logits inside buffer 0
    const uint32_t o_tmem = tmem_base + lane_base + kTileCols * i + kColO + 64 * hh;   // this half's accumulator
    const float c = args.scale_log2;
    uint32_t n_blk = 0, n_job = 0;    // n_blk: running block count of this tile (block g lives in S buffer g & 1)

    for (int job = blockIdx.x; job < n_jobs; job += stride) {
      const Job jb = decode_job(job, args);
      if (i == 1 && !jb.a1) continue;
      const int t = i ? jb.t1 : jb.t0;
      const bool warp_active = t * kBlock + wq * 32 < S;   // a warp of padding rows only keeps the barriers moving

      float m = -INFINITY;   // reference maximum used in the exponent (raw logit units)
      float l = 0.f;         // running row sum of P over this half's columns

      // rare: raise the reference maximum to m_new, rescale this row's O half and l.  The accumulator is quiescent once the
      // previous block's P V of this half retired, and the next one cannot be issued before this warp reports p_ready.
      auto rescale = [&](float m_new) {
        mbar_wait(&bars->pv_done[i][hh], (n_blk - 1) & 1);
        tc_fence_after();
        const float alpha = ex2((m - m_new) * c);
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          uint32_t o[16];
          tmem_ld16p(o_tmem + 16 * h, o);
          tmem_ld_wait16(o);
#pragma unroll
          for (int x = 0; x < 16; ++x) o[x] = __float_as_uint(__uint_as_float(o[x]) * alpha);
          tmem_st16p(o_tmem + 16 * h, o);
        }
        tmem_st_wait();
        l *= alpha;
      };

      for (int j = 0; j < nb; ++j) {
        const int b = n_blk & 1;
        const uint32_t s_tmem = s_tmem0 + kSub * b;
        mbar_wait(&bars->s_full[i][b], (n_blk >> 1) & 1);
        tc_fence_after();
        // valid key columns of this half in this block
        const int v = (j < nb - 1) ? kHalf : min(max(last_valid - kHalf * hh, 0), kHalf);
        if (!warp_active || v == 0) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->p_ready[i][b][hh]);
          ++n_blk;
          continue;
        }
        if (v == kHalf) {
          // ---- 32 valid key columns: one branch-free basic block, P held in registers until the sum is checked
          uint32_t pk[16];
          float bs;
          bool exact = false;   // m is known to be >= every logit of this block
          uint32_t ra[16], rb[16];
          tmem_ld16p(s_tmem, ra);
          tmem_ld16p(s_tmem + 16, rb);
          if (j == 0) {
            // reference maximum of a job: of this half's first 16 logits only (any reference works as long as the block sums
            // stay below kSumLimit; the class token's key is column 0)
            tmem_ld_wait16(ra);
            const float b0 = fmax3(fmax3(__uint_as_float(ra[0]), __uint_as_float(ra[1]), __uint_as_float(ra[2])),
                                   __uint_as_float(ra[3]), __uint_as_float(ra[4]));
            const float b1 = fmax3(fmax3(__uint_as_float(ra[5]), __uint_as_float(ra[6]), __uint_as_float(ra[7])),
                                   __uint_as_float(ra[8]), __uint_as_float(ra[9]));
            const float b2 = fmax3(fmax3(__uint_as_float(ra[10]), __uint_as_float(ra[11]), __uint_as_float(ra[12])),
                                   __uint_as_float(ra[13]), __uint_as_float(ra[14]));
            m = fmax3(b0, b1, fmaxf(b2, __uint_as_float(ra[15])));
          }
          for (;;) {
            const float mc = m * c;
            const float2 c2 = make_float2(c, c), nmc2 = make_float2(-mc, -mc);
            float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
            tmem_ld_wait16(ra);
            exp_chunk<POLY>(ra, pk, c2, nmc2, acc0, acc1);
            tmem_ld_wait16(rb);
            exp_chunk<POLY>(rb, pk + 8, c2, nmc2, acc0, acc1);
            bs = (acc0.x + acc0.y) + (acc1.x + acc1.y);
            if (!exact && __any_sync(0xffffffffu, !(bs < kSumLimit))) {
              // rare: some P may not fit fp16.  Take the exact maximum of the block, rescale O and l, redo the block.
              float b0 = -INFINITY, b1 = -INFINITY;
#pragma unroll
              for (int x = 0; x < 16; x += 4) {
                b0 = fmax3(b0, __uint_as_float(ra[x]), __uint_as_float(ra[x + 1]));
                b1 = fmax3(b1, __uint_as_float(ra[x + 2]), __uint_as_float(ra[x + 3]));
                b0 = fmax3(b0, __uint_as_float(rb[x]), __uint_as_float(rb[x + 1]));
                b1 = fmax3(b1, __uint_as_float(rb[x + 2]), __uint_as_float(rb[x + 3]));
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
          // P (packed fp16) over the first 16 columns of this half's logits: they are all in registers
          tmem_st16p(s_tmem, pk);
        } else {
          // ---- ragged last block: nfull whole 16-column chunks + `rem` valid columns of one more.  Exact maximum first
          // (lazy rescale: P <= 2^8), then P chunk by chunk (P of chunk ch lands on S columns of chunks <= ch, already read).
          const int nfull = v >> 4, rem = v & 15;
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
        if (lane == 0) mbar_arrive(&bars->p_ready[i][b][hh]);
        ++n_blk;
      }

      // publish the row statistics; the epilogue of this tile's previous job must have read its own first
      if (n_job > 0) mbar_wait(&bars->o_free[i], (n_job - 1) & 1);
      lm[((i * 2 + hh) * 2 + 0) * kBlock + row] = l;
      lm[((i * 2 + hh) * 2 + 1) * kBlock + row] = m;
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->l_ready[i]);
      ++n_job;
    }
  } else if (warp < kWarpTma) {
    // ---------------------------------------------------------------- epilogue warps: merge the halves, O / l -> fp16 -> global
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    const int wq = warp & 3;
    const int row = wq * 32 + lane;
    const uint32_t lane_base = uint32_t(wq * 32) << 16;
    const int heads = args.hidden / kHeadDim;
    const float c = args.scale_log2;
    uint32_t n_e0 = 0, n_e1 = 0;
    for (int job = blockIdx.x; job < n_jobs; job += stride) {
      const Job jb = decode_job(job, args);
      for (int i = 0; i < 2; ++i) {
        if (i == 1 && !jb.a1) continue;
        const uint32_t n_e = i ? n_e1 : n_e0;
        const int h = i ? jb.h1 : jb.h0;
        const int token = (i ? jb.t1 : jb.t0) * kBlock + row;
        mbar_wait(&bars->l_ready[i], n_e & 1);
        mbar_wait(&bars->o_full[i], n_e & 1);
        if (i) ++n_e1; else ++n_e0;
        tc_fence_after();
        const float la = lm[((i * 2 + 0) * 2 + 0) * kBlock + row], ma = lm[((i * 2 + 0) * 2 + 1) * kBlock + row];
        const float lb = lm[((i * 2 + 1) * 2 + 0) * kBlock + row], mb = lm[((i * 2 + 1) * 2 + 1) * kBlock + row];
        // a half without a single valid key column (S <= 32) has l == 0 and an accumulator that was never written
        const bool use_a = la > 0.f, use_b = lb > 0.f;
        const float mm = fmaxf(use_a ? ma : -INFINITY, use_b ? mb : -INFINITY);
        const float wa = use_a ? ex2((ma - mm) * c) : 0.f, wb = use_b ? ex2((mb - mm) * c) : 0.f;
        const float l_sum = la * wa + lb * wb;
        const float inv_l = 1.0f / l_sum;
        const float fa = wa * inv_l, fb = wb * inv_l;
        const uint32_t o_tmem = tmem_base + lane_base + kTileCols * i + kColO;
        const bool store = token < S;
        uint4* o4 = reinterpret_cast<uint4*>(args.out + ((size_t)jb.view * S + (store ? token : 0)) * args.hidden + h * kHeadDim);
#pragma unroll
        for (int q8 = 0; q8 < 8; ++q8) {   // 8 output columns at a time
          uint32_t oa[8], ob[8];
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(oa[0]), "=r"(oa[1]), "=r"(oa[2]), "=r"(oa[3]), "=r"(oa[4]), "=r"(oa[5]), "=r"(oa[6]), "=r"(oa[7])
                       : "r"(o_tmem + 8 * q8) : "memory");
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(ob[0]), "=r"(ob[1]), "=r"(ob[2]), "=r"(ob[3]), "=r"(ob[4]), "=r"(ob[5]), "=r"(ob[6]), "=r"(ob[7])
                       : "r"(o_tmem + 64 + 8 * q8) : "memory");
          tmem_ld_wait();
          float f[8];
#pragma unroll
          for (int x = 0; x < 8; ++x) {
            const float va = use_a ? __uint_as_float(oa[x]) * fa : 0.f;
            f[x] = use_b ? fmaf(__uint_as_float(ob[x]), fb, va) : va;
          }
          if (store)
            o4[q8] = make_uint4(pack_half2(f[0], f[1]), pack_half2(f[2], f[3]), pack_half2(f[4], f[5]), pack_half2(f[6], f[7]));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->o_free[i]);
        if (store && args.lse2 != nullptr)
          args.lse2[((size_t)jb.view * heads + h) * S + token] = mm * c + log2f(l_sum);
      }
    }
  } else {
   asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
   if (warp > kWarpMma + 1) {
    // idle warp of the last warpgroup
   } else if (warp == kWarpTma) {
    // ---------------------------------------------------------------- TMA producer (warp-uniform, one elected lane issues)
    if (blockIdx.x < n_jobs) {
      uint32_t pos = 0;              // ring position
      uint32_t nq0 = 0, nq1 = 0;     // Q tiles loaded per tile (job ordinal -> staging buffer parity)
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
        const int par = nq & 1;
        mbar_wait(&bars->q_empty[I][par], ((nq >> 1) & 1) ^ 1);
        if (elect_one()) {
          const int col = (I ? jb.h1 : jb.h0) * kHeadDim, row = jb.view * S + (I ? jb.t1 : jb.t0) * kBlock;
          uint8_t* dst = smem_q + (I * 2 + par) * kQBytes;
          mbar_arrive_expect_tx(&bars->q_full[I][par], kQBytes);
          tma_load_2d(dst, &tmap_qkv, &bars->q_full[I][par], col, row);
          tma_load_2d(dst + kSubBytes, &tmap_qkv, &bars->q_full[I][par], col, row + kSub);
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
    // ---------------------------------------------------------------- MMA issuers: warp 21 -> tile 0, warp 22 -> tile 1
    // (warp-uniform; one elected lane issues)
    const int I = warp - kWarpMma;
    uint32_t pos = 0;                                    // ring position at the start of the current job
    uint32_t g = 0;                                      // running block count of this tile: block g lives in S buffer g & 1
    uint32_t n_j = 0;                                    // jobs this tile took part in (job ordinal -> Q staging buffer parity)
    bool pre = false;                                    // S of this job's first two blocks was issued during the previous job
    // smem descriptors: Q and K are K-major (rows of 128 B, k-step = 32 B), V is MN-major (row = kv index, k-step = 16 rows)
    const uint64_t q_desc0 = make_smem_desc(smem_u32(smem_q) + I * 2 * kQBytes, 16, 1024, kLayoutSw128);
    const uint64_t k_desc0 = make_smem_desc(smem_u32(smem_kv), 16, 1024, kLayoutSw128);
    const uint64_t v_desc0 = make_smem_desc(smem_u32(smem_kv), 1024, 1024, kLayoutSw128);
    const uint32_t idesc_s = make_idesc_f16(kBlock, kSub, 0, 0);
    const uint32_t idesc_s_last = make_idesc_f16(kBlock, last_n, 0, 0);
    const uint32_t idesc_pv = make_idesc_f16(kBlock, kHeadDim, 0, 1);    // B (= V) is MN-major
    const uint32_t s_base = tmem_base + kTileCols * I, o_base = tmem_base + kTileCols * I + kColO;
    // valid key columns of the two halves of the last block -> number of 16-row k-steps of their P V
    const int ks_a_last = (min(last_valid, kHalf) + 15) >> 4, ks_b_last = (max(last_valid - kHalf, 0) + 15) >> 4;
    // S[buf] = Q_I K_j^T with Q of job ordinal `jq`, j = block index inside its job   (inside an elected region)
    auto issue_s = [&](uint64_t kd, int j, uint32_t buf, uint32_t jq) {
      const uint32_t idesc = (j == nb - 1) ? idesc_s_last : idesc_s;
      const uint32_t d = s_base + kSub * buf;
      const uint64_t qd = q_desc0 + (uint64_t)((jq & 1) * (kQBytes >> 4));
#pragma unroll
      for (int k = 0; k < kHeadDim / 16; ++k) umma_ss(d, qd + 2 * k, kd + 2 * k, idesc, k != 0);
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
        const uint32_t buf = (g + (uint32_t)(j < 0 ? 0 : j)) & 1;
        if (step == 0) {
          mbar_wait(&bars->q_full[I][n_j & 1], (n_j >> 1) & 1);
        } else {
          const uint32_t gb = g + j;
          mbar_wait(&bars->p_ready[I][gb & 1][0], (gb >> 1) & 1);
          if (j == 0) mbar_wait(&bars->o_free[I], (n_j & 1) ^ 1);   // the previous job's O halves were read out
        }
        mbar_wait(&bars->kv_full[slot], (p / kSlots) & 1);
        tc_fence_after();
        const uint64_t d0 = (uint64_t)(slot * (kSlotBytes >> 4)), d1 = d0 + (kSubBytes >> 4);
        if (step == 0) {
          if (elect_one()) {
            issue_s(k_desc0 + d0, 0, g & 1, n_j);
            if (nb > 1) issue_s(k_desc0 + d1, 1, (g + 1) & 1, n_j);
            tc_commit(&bars->kv_empty[slot]);
            if (ncommit == 2) tc_commit(&bars->kv_empty[slot]);
          }
          __syncwarp();
        } else {
          const uint64_t vd = v_desc0 + d0;
          const uint32_t pa = s_base + kSub * buf;                 // P_A (packed fp16) of this block; P_B sits 32 columns on
          const int ks_a = (j == nb - 1) ? ks_a_last : 2, ks_b = (j == nb - 1) ? ks_b_last : 2;
          if (elect_one()) {                                      // O_A += P_A V[0:32)
            for (int k = 0; k < ks_a; ++k) umma_ts(o_base, pa + 8 * k, vd + 128 * k, idesc_pv, (j | k) != 0);
            tc_commit(&bars->pv_done[I][0]);
          }
          __syncwarp();
          mbar_wait(&bars->p_ready[I][buf][1], ((g + j) >> 1) & 1);
          tc_fence_after();
          if (elect_one()) {                                      // O_B += P_B V[32:64)
            for (int k = 0; k < ks_b; ++k)
              umma_ts(o_base + 64, pa + kHalf + 8 * k, vd + 128 * (2 + k), idesc_pv, (j | k) != 0);
            tc_commit(&bars->pv_done[I][1]);
            if (j == nb - 1) {
              tc_commit(&bars->o_full[I]);
              if (!has_next) tc_commit(&bars->q_empty[I][n_j & 1]);    // (otherwise released after the next job's S blocks)
            }
            if (j + 2 < nb) issue_s(k_desc0 + d1, j + 2, buf, n_j);
            tc_commit(&bars->kv_empty[slot]);
            if (ncommit == 2) tc_commit(&bars->kv_empty[slot]);
          }
          __syncwarp();
        }
        if (has_next && j >= nb - 2) {
          // block j's buffer is free again (both P V of it are issued): the next job's block j - (nb - 2) goes there
          const int slot_n = next_p % kSlots;
          if (j == nb - 2) {
            mbar_wait(&bars->q_full[I][(n_j + 1) & 1], ((n_j + 1) >> 1) & 1);
            mbar_wait(&bars->kv_full[slot_n], (next_p / kSlots) & 1);
            tc_fence_after();
          }
          if (elect_one()) {
            const uint64_t d0n = (uint64_t)(slot_n * (kSlotBytes >> 4));
            if (j == nb - 2) {
              issue_s(k_desc0 + d0n, 0, (g + nb) & 1, n_j + 1);
            } else {
              issue_s(k_desc0 + d0n + (kSubBytes >> 4), 1, (g + nb + 1) & 1, n_j + 1);
              tc_commit(&bars->kv_empty[slot_n]);
              if (next_commit == 2) tc_commit(&bars->kv_empty[slot_n]);
              tc_commit(&bars->q_empty[I][n_j & 1]);     // every S MMA of THIS job retired before this commit fires
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
int launch_split(int variant, const void* qkv, void* out, int n_views, int seq, int heads, cudaStream_t stream,
                float* lse2) {
  const int hidden = heads * kHeadDim;
  CUtensorMap tm;
  if (make_tmap_f16_2d(&tm, qkv, (uint64_t)n_views * seq, 3 * hidden, 3 * hidden, kSub, kHeadDim)) return 1;
  auto kern = attention_split_kernel<POLY>;
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
int attention_split_f16(const void* qkv, void* out, int n_views, int seq, int heads, cudaStream_t stream, float* lse2,
                       int poly) {
  if (n_views <= 0) return 0;
  switch (poly) {
    case 0: return launch_split<0x00>(0, qkv, out, n_views, seq, heads, stream, lse2);
    case 1: return launch_split<0x08>(1, qkv, out, n_views, seq, heads, stream, lse2);
    case 3: return launch_split<0x4A>(3, qkv, out, n_views, seq, heads, stream, lse2);
    case 4: return launch_split<0xAA>(4, qkv, out, n_views, seq, heads, stream, lse2);
    default: return launch_split<0x88>(2, qkv, out, n_views, seq, heads, stream, lse2);
  }
}

}  // namespace pg
