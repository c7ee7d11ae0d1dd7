// Backward of the multi-head attention core on sm_100a (tcgen05 + TMEM + TMA), head_dim 64 — part of the fine-tune
// step ("next" row N1; what autograd runs for HF CLIPAttention in reference training/train_eval_loop.py:216).
//
//   S = scale Q K^T,  P = softmax(S),  O = P V          (forward, attention_tcgen05.cu; it keeps lse2 = log2 sum exp)
//   dV = P^T dO,  dP = dO V^T,  dS = P o (dP - delta),  delta = rowsum(dO o O),  dQ = scale dS K,  dK = scale dS^T Q
//
// Two launches of one kernel template, each "row tile" of 128 rows owning its accumulators in TMEM for the whole loop,
// so that no atomics and no transposes are needed and every MMA uses an operand form the forward kernel already uses:
//   MODE_DQ   CTA = (128 query rows, head, view), loops over KV blocks of 64:
//             S = Q K_j^T (f16), dP = dO V_j^T (bf16)  ->  dS  ->  dQ += dS K_j        (dS: TMEM A operand, K_j: MN-major B)
//   MODE_DKV  CTA = (128 key/value rows, head, view), loops over query blocks of 64:
//             S^T = K Q_j^T (f16), dP^T = V dO_j^T (bf16)  ->  P^T, dS^T  ->  dV += P^T dO_j,  dK += dS^T Q_j
// The logits are recomputed from the SAME fp16 q/k the forward used (so P is reproduced exactly); everything that carries
// gradient magnitude (dO, dS, and the q/k/v copies they multiply) is bf16 — gradients underflow fp16's range.
//
//   warp 0     TMA producer: the row tile's two operands once, then three 64 x 64 tiles per block through a 2-slot ring
//   warp 1     TMEM allocator + MMA issuer
//   warps 2-5  one TMEM lane (= one row of the row tile) per thread: P / dS from S, dP, lse2, delta; written back to TMEM
//              as packed bf16 over the columns they were read from
// TMEM (256 columns): S [0,64)  dP [64,128)  acc1 [128,192) (dQ | dV)  acc2 [192,256) (dK).  Correctness-first version:
// one S/dP buffer (the MMA warp and the row threads alternate); two CTAs per SM overlap each other's bubbles.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "attention.h"
#include "prof.h"
#include "ptx.cuh"
#include "tma_host.h"

namespace pg {

namespace {

constexpr int kHeadDim = 64;
constexpr int kRows = 128;                             // row tile
constexpr int kRowTileBytes = kRows * kHeadDim * 2;    // 16 KB
constexpr int kTmemCols = 256;
constexpr uint32_t kBf16Fmt = (1u << 7) | (1u << 10);

// Column-block width and number of S/dP buffers in TMEM.  <64, 1>: one buffer, the MMA warp and the row threads alternate
// (first version).  <32, 2>: two buffers of 32 columns — the S/dP MMAs of block j+1 run while the row threads work on
// block j (default).  Both fit 256 TMEM columns: NBUF * 2 * BLK for S/dP, then acc1 and acc2 (64 columns each).
template <int BLK, int NBUF, int CW = 4>
struct BwdCfg {
  // CW: row-thread warps.  8 = two warps per TMEM lane quarter, each taking one 32-column half of a 64-wide block
  // (its packed outputs stay inside the columns it read, so the two warps never touch each other's data).
  static constexpr int kComputeWarps = CW;
  static constexpr int kThreadsCfg = 64 + 32 * CW;
  static_assert(CW == 4 || (CW == 8 && BLK == 64), "8 row warps need 64-wide blocks");
  static constexpr int kBlk = BLK, kNumBuf = NBUF;
  static constexpr int kTileBytes = BLK * kHeadDim * 2;
  static constexpr int kStageBytes = 3 * kTileBytes;
  static constexpr int kStages = (BLK == 64) ? 2 : 4;
  static constexpr int kAcc1Col = NBUF * 2 * BLK, kAcc2Col = kAcc1Col + kHeadDim;
  static constexpr int kSmemBytes = 2 * kRowTileBytes + kStages * kStageBytes + 2 * 128 * 4 + 1024 + 256;
  static_assert(kAcc2Col + kHeadDim <= kTmemCols, "TMEM budget");
};

enum { MODE_DQ = 0, MODE_DKV = 1 };

struct BwdArgs {
  int seq, hidden;
  const float* lse2;    // [n_views * heads, seq]
  const float* delta;   // [n_views * heads, seq]
  __nv_bfloat16* dqkv;  // [n_views * seq, 3 * hidden]
  float scale;          // 1 / sqrt(64)
  float scale_log2;     // scale * log2(e)
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;"
      : "=l"(d)
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ float4 lds_f32x4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

template <int MODE, class Cfg>
__global__ void __launch_bounds__(Cfg::kThreadsCfg, 2)
attention_bwd_kernel(const __grid_constant__ CUtensorMap tmap_f16, const __grid_constant__ CUtensorMap tmap_bf16,
                     const __grid_constant__ CUtensorMap tmap_do, const BwdArgs args) {
  constexpr int kBlk = Cfg::kBlk, kNumBuf = Cfg::kNumBuf, kStages = Cfg::kStages, kCW = Cfg::kComputeWarps;
  constexpr int kTileBytes = Cfg::kTileBytes, kStageBytes = Cfg::kStageBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_r16 = smem;                              // row tile, fp16 (Q | K)
  uint8_t* smem_rbf = smem + kRowTileBytes;              // row tile, bf16 (dO | V)
  uint8_t* smem_st = smem + 2 * kRowTileBytes;           // ring: [C16 | Cbf1 | Cbf2] per slot
  float* cbuf = reinterpret_cast<float*>(smem_st + kStages * kStageBytes);   // [2][128]: lse2 | delta of a column block
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(cbuf) + 2 * 128 * 4);
  uint64_t* full_bar = bars;                 // [kStages] TMA -> MMA
  uint64_t* empty_bar = bars + kStages;      // [kStages] MMA -> TMA
  uint64_t* r_full = bars + 2 * kStages;     // row tile landed
  uint64_t* s_full = r_full + 1;             // [kNumBuf] MMA -> rows : S and dP of a block complete in buffer b
  uint64_t* ds_ready = s_full + kNumBuf;     // [kNumBuf] rows -> MMA : P / dS written over buffer b (4 warps arrive)
  uint64_t* acc_full = ds_ready + kNumBuf;   // MMA -> rows : accumulators complete
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, head = blockIdx.y, view = blockIdx.z;
  const int S = args.seq;
  const int heads = args.hidden / kHeadDim;
  const int nb = (S + kBlk - 1) / kBlk;
  const int row0 = view * S;
  const int q_col = head * kHeadDim, k_col = args.hidden + q_col, v_col = 2 * args.hidden + q_col;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_f16);
    tma_prefetch_desc(&tmap_bf16);
    tma_prefetch_desc(&tmap_do);
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(r_full, 1);
    for (int b = 0; b < kNumBuf; ++b) { mbar_init(&s_full[b], 1); mbar_init(&ds_ready[b], kCW); }
    mbar_init(acc_full, 1);
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
  constexpr uint32_t kAcc1Col = Cfg::kAcc1Col, kAcc2Col = Cfg::kAcc2Col;
  // buffer b: S at columns [2 b BLK, +BLK), dP at [2 b BLK + BLK, +BLK)

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      const int r_row = row0 + tile * kRows;
      mbar_arrive_expect_tx(r_full, 2 * kRowTileBytes);
#pragma unroll
      for (int part = 0; part < kRows / kBlk; ++part) {
        if (MODE == MODE_DQ) {
          tma_load_2d(smem_r16 + part * kTileBytes, &tmap_f16, r_full, q_col, r_row + part * kBlk);
          tma_load_2d(smem_rbf + part * kTileBytes, &tmap_do, r_full, q_col, r_row + part * kBlk);
        } else {
          tma_load_2d(smem_r16 + part * kTileBytes, &tmap_f16, r_full, k_col, r_row + part * kBlk);
          tma_load_2d(smem_rbf + part * kTileBytes, &tmap_bf16, r_full, v_col, r_row + part * kBlk);
        }
      }
      for (int j = 0; j < nb; ++j) {
        const int slot = j % kStages;
        mbar_wait(&empty_bar[slot], ((j / kStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&full_bar[slot], kStageBytes);
        uint8_t* st = smem_st + slot * kStageBytes;
        const int c_row = row0 + j * kBlk;
        if (MODE == MODE_DQ) {
          tma_load_2d(st, &tmap_f16, &full_bar[slot], k_col, c_row);                       // K_j  fp16  (S)
          tma_load_2d(st + kTileBytes, &tmap_bf16, &full_bar[slot], v_col, c_row);         // V_j  bf16  (dP)
          tma_load_2d(st + 2 * kTileBytes, &tmap_bf16, &full_bar[slot], k_col, c_row);     // K_j  bf16  (dQ)
        } else {
          tma_load_2d(st, &tmap_f16, &full_bar[slot], q_col, c_row);                       // Q_j  fp16  (S^T)
          tma_load_2d(st + kTileBytes, &tmap_do, &full_bar[slot], q_col, c_row);           // dO_j bf16  (dP^T, dV)
          tma_load_2d(st + 2 * kTileBytes, &tmap_bf16, &full_bar[slot], q_col, c_row);     // Q_j  bf16  (dK)
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_f16(kRows, kBlk, 0, 0);                  // fp16 x fp16, both K-major
      const uint32_t idesc_dp = make_idesc_f16(kRows, kBlk, 0, 0) | kBf16Fmt;      // bf16 x bf16, both K-major
      const uint32_t idesc_acc = make_idesc_f16(kRows, kHeadDim, 0, 1) | kBf16Fmt; // A from TMEM, B MN-major
      const uint32_t r16_addr = smem_u32(smem_r16), rbf_addr = smem_u32(smem_rbf);
      mbar_wait(r_full, 0);
      tc_fence_after();
      auto issue_sdp = [&](int j) {      // S (fp16) and dP (bf16) of block j into buffer j % kNumBuf
        const int slot = j % kStages;
        const uint32_t buf = tmem_base + (j % kNumBuf) * 2 * kBlk;
        mbar_wait(&full_bar[slot], (j / kStages) & 1);
        tc_fence_after();
        const uint32_t c16 = smem_u32(smem_st + slot * kStageBytes), cb1 = c16 + kTileBytes;
#pragma unroll
        for (int k = 0; k < kHeadDim / 16; ++k)
          umma_ss(buf, make_smem_desc(r16_addr + k * 32, 16, 1024, kLayoutSw128),
                  make_smem_desc(c16 + k * 32, 16, 1024, kLayoutSw128), idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < kHeadDim / 16; ++k)
          umma_ss(buf + kBlk, make_smem_desc(rbf_addr + k * 32, 16, 1024, kLayoutSw128),
                  make_smem_desc(cb1 + k * 32, 16, 1024, kLayoutSw128), idesc_dp, k != 0);
        tc_commit(&s_full[j % kNumBuf]);
      };
      auto issue_acc = [&](int j) {      // accumulate with the packed bf16 P / dS the row threads left in buffer j % kNumBuf
        const int slot = j % kStages;
        const uint32_t buf = tmem_base + (j % kNumBuf) * 2 * kBlk;
        const uint32_t cb1 = smem_u32(smem_st + slot * kStageBytes) + kTileBytes, cb2 = cb1 + kTileBytes;
        mbar_wait(&ds_ready[j % kNumBuf], (j / kNumBuf) & 1);
        tc_fence_after();
        if (MODE == MODE_DQ) {
#pragma unroll
          for (int k = 0; k < kBlk / 16; ++k)      // dQ += dS K_j : contraction over the block's kv rows
            umma_ts(tmem_base + kAcc1Col, buf + kBlk + 32 * (k >> 1) + 8 * (k & 1),
                    make_smem_desc(cb2 + k * 16 * 128, 1024, 1024, kLayoutSw128), idesc_acc, (j | k) != 0);
        } else {
#pragma unroll
          for (int k = 0; k < kBlk / 16; ++k)      // dV += P^T dO_j : contraction over the block's query rows
            umma_ts(tmem_base + kAcc1Col, buf + 32 * (k >> 1) + 8 * (k & 1),
                    make_smem_desc(cb1 + k * 16 * 128, 1024, 1024, kLayoutSw128), idesc_acc, (j | k) != 0);
#pragma unroll
          for (int k = 0; k < kBlk / 16; ++k)      // dK += dS^T Q_j
            umma_ts(tmem_base + kAcc2Col, buf + kBlk + 32 * (k >> 1) + 8 * (k & 1),
                    make_smem_desc(cb2 + k * 16 * 128, 1024, 1024, kLayoutSw128), idesc_acc, (j | k) != 0);
        }
        tc_commit(&empty_bar[slot]);
      };
      for (int j = 0; j < kNumBuf && j < nb; ++j) issue_sdp(j);
      for (int j = 0; j < nb; ++j) {
        issue_acc(j);
        if (j + kNumBuf < nb) issue_sdp(j + kNumBuf);   // overwrites buffer j % kNumBuf after its accumulate MMAs (in order)
      }
      tc_commit(acc_full);
    }
  } else {
    // ---------------------------------------------------------------- row threads
    const int qd = warp & 3;                                  // TMEM lane quarter this warp may touch
    const int cw = warp - 2;                                  // 0 .. kCW-1
    const int ct = cw * 32 + lane;                            // index among the row threads (the first 128 stage cbuf)
    const int h_lo = (kCW == 8) ? (cw >> 2) : 0;              // 8 warps: this warp owns ONE 32-column half of each block
    const int h_hi = (kCW == 8) ? h_lo + 1 : kBlk / 32;
    const uint32_t lane_base = uint32_t(qd * 32) << 16;
    const int r = tile * kRows + qd * 32 + lane;              // row index inside the view
    const bool row_valid = r < S;
    const float c = args.scale_log2, scale = args.scale;
    const size_t stat0 = ((size_t)view * heads + head) * S;
    float lse_r = 0.f, delta_r = 0.f;
    if (MODE == MODE_DQ && row_valid) { lse_r = args.lse2[stat0 + r]; delta_r = args.delta[stat0 + r]; }
    // fast path constants (MODE_DQ): dS*scale = 2^(S c - lse - 3) * (dP - delta); rows past the view get P = 0
    const float nl_r = row_valid ? -(lse_r + 3.0f) : -1e30f;
    const float2 c2 = make_float2(c, c), nl2 = make_float2(nl_r, nl_r), nd2 = make_float2(-delta_r, -delta_r);
    const float2 sc2 = make_float2(scale, scale);
    const uint32_t cbuf_addr = smem_u32(cbuf);

    for (int j = 0; j < nb; ++j) {
      const uint32_t cb = cbuf_addr + (j & 1) * 512;
      if (MODE == MODE_DKV) {
        // per-column statistics of this query block: cb[0, kBlk) = -lse2 (-1e30 past the view: P becomes exactly 0),
        // cb[64, 64 + kBlk) = -delta * scale
        const int qi = j * kBlk + (ct & 63);
        if (ct < 128) {
          float v = (ct < 64) ? -1e30f : 0.f;
          if ((ct & 63) < kBlk && qi < S) v = (ct < 64) ? -args.lse2[stat0 + qi] : -args.delta[stat0 + qi] * scale;
          sts_f32(cb + ct * 4, v);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kCW) : "memory");
      }
      // MODE_DQ masks kv columns past the view per element; only its last, ragged block needs that
      const bool ragged = (MODE == MODE_DQ) && (j == nb - 1) && (S % kBlk != 0);
      const uint32_t buf = tmem_base + lane_base + (j % kNumBuf) * 2 * kBlk;   // S at buf, dP at buf + kBlk
      mbar_wait(&s_full[j % kNumBuf], (j / kNumBuf) & 1);
      tc_fence_after();
      for (int h = h_lo; h < h_hi; ++h) {
        uint32_t sv[32], dv[32], pk_p[16], pk_ds[16];
        tmem_ld32(buf + 32 * h, sv);
        tmem_ld32(buf + kBlk + 32 * h, dv);
        tmem_ld_wait();
        if (!ragged) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {   // packed fp32 pairs: 4 columns per iteration
            const float2 s01 = make_float2(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1]));
            const float2 s23 = make_float2(__uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3]));
            const float2 d01 = make_float2(__uint_as_float(dv[i]), __uint_as_float(dv[i + 1]));
            const float2 d23 = make_float2(__uint_as_float(dv[i + 2]), __uint_as_float(dv[i + 3]));
            if (MODE == MODE_DQ) {
              const float2 x01 = ffma2(s01, c2, nl2), x23 = ffma2(s23, c2, nl2);
              const float2 p01 = make_float2(ex2(x01.x), ex2(x01.y)), p23 = make_float2(ex2(x23.x), ex2(x23.y));
              const float2 g01 = fmul2(p01, fadd2(d01, nd2)), g23 = fmul2(p23, fadd2(d23, nd2));
              pk_ds[i >> 1] = pack_bf16x2(g01.x, g01.y);
              pk_ds[(i >> 1) + 1] = pack_bf16x2(g23.x, g23.y);
            } else {
              const float4 nl = lds_f32x4(cb + (32 * h + i) * 4), nd = lds_f32x4(cb + (64 + 32 * h + i) * 4);
              const float2 x01 = ffma2(s01, c2, make_float2(nl.x, nl.y)), x23 = ffma2(s23, c2, make_float2(nl.z, nl.w));
              const float2 p01 = make_float2(ex2(x01.x), ex2(x01.y)), p23 = make_float2(ex2(x23.x), ex2(x23.y));
              const float2 g01 = fmul2(p01, ffma2(d01, sc2, make_float2(nd.x, nd.y)));
              const float2 g23 = fmul2(p23, ffma2(d23, sc2, make_float2(nd.z, nd.w)));
              pk_p[i >> 1] = pack_bf16x2(p01.x, p01.y);
              pk_p[(i >> 1) + 1] = pack_bf16x2(p23.x, p23.y);
              pk_ds[i >> 1] = pack_bf16x2(g01.x, g01.y);
              pk_ds[(i >> 1) + 1] = pack_bf16x2(g23.x, g23.y);
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {   // MODE_DQ, last block: per-element masks
            float ds[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int col = j * kBlk + 32 * h + i + e;        // index of the kv column inside the view
              const bool ok = row_valid && (col < S);
              const float pe = ex2(fmaf(__uint_as_float(sv[i + e]), c, -lse_r));
              ds[e] = ok ? pe * (__uint_as_float(dv[i + e]) - delta_r) * scale : 0.f;
            }
            pk_ds[i >> 1] = pack_bf16x2(ds[0], ds[1]);
          }
        }
        // packed bf16 pairs over columns this thread has just consumed: P at S[32h, 32h+16), dS at dP[32h, 32h+16)
        if (MODE == MODE_DKV) tmem_st16(buf + 32 * h, pk_p);
        tmem_st16(buf + kBlk + 32 * h, pk_ds);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ds_ready[j % kNumBuf]);
    }

    // epilogue: accumulators -> bf16 -> dqkv
    mbar_wait(acc_full, 0);
    tc_fence_after();
    __nv_bfloat16* orow = args.dqkv + (size_t)(row0 + r) * (3 * args.hidden);
    constexpr int kNumAcc = (MODE == MODE_DQ) ? 1 : 2;
#pragma unroll
    for (int a = 0; a < kNumAcc; ++a) {
      // MODE_DQ: acc1 = dQ -> q section;  MODE_DKV: acc1 = dV -> v section, acc2 = dK -> k section
      const int out_col = (MODE == MODE_DQ) ? q_col : (a == 0 ? v_col : k_col);
      const uint32_t acc = tmem_base + lane_base + (a == 0 ? kAcc1Col : kAcc2Col);
      for (int c0 = (kCW == 8 ? 32 * (cw >> 2) : 0); c0 < (kCW == 8 ? 32 * (cw >> 2) + 32 : kHeadDim); c0 += 32) {
        uint32_t o[32];
        tmem_ld32(acc + c0, o);
        tmem_ld_wait();
        if (row_valid) {
          uint4* o4 = reinterpret_cast<uint4*>(orow + out_col + c0);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 v;
            v.x = pack_bf16x2(__uint_as_float(o[8 * i + 0]), __uint_as_float(o[8 * i + 1]));
            v.y = pack_bf16x2(__uint_as_float(o[8 * i + 2]), __uint_as_float(o[8 * i + 3]));
            v.z = pack_bf16x2(__uint_as_float(o[8 * i + 4]), __uint_as_float(o[8 * i + 5]));
            v.w = pack_bf16x2(__uint_as_float(o[8 * i + 6]), __uint_as_float(o[8 * i + 7]));
            o4[i] = v;
          }
        }
        __syncwarp();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int MODE, class Cfg>
int launch_bwd(const CUtensorMap& t16, const CUtensorMap& tbf, const CUtensorMap& tdo, const BwdArgs& a, int n_views,
               int heads, cudaStream_t stream) {
  constexpr int kSmemBytes = Cfg::kSmemBytes;
  auto kern = attention_bwd_kernel<MODE, Cfg>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) { set_last_error("attention_backward: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 1; }
    attr_set = true;
  }
  dim3 grid((a.seq + kRows - 1) / kRows, heads, n_views);
  ProfScope prof(MODE == MODE_DQ ? "attention_bwd_dq" : "attention_bwd_dkv", stream);
  kern<<<grid, Cfg::kThreadsCfg, kSmemBytes, stream>>>(t16, tbf, tdo, a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_last_error("attention_backward launch: %s", cudaGetErrorString(e)); return 1; }
  return 0;
}

template <class Cfg>
int run_backward(const void* qkv_f16, const void* qkv_bf16, const void* d_out_bf16, const float* lse2, const float* delta,
                 void* dqkv_bf16, int n_views, int seq, int heads, cudaStream_t stream) {
  const int hidden = heads * kHeadDim;
  const uint64_t rows = (uint64_t)n_views * seq;
  CUtensorMap t16, tbf, tdo;
  if (make_tmap_f16_2d(&t16, qkv_f16, rows, 3 * hidden, 3 * hidden, Cfg::kBlk, kHeadDim)) return 1;
  if (make_tmap_f16_2d(&tbf, qkv_bf16, rows, 3 * hidden, 3 * hidden, Cfg::kBlk, kHeadDim)) return 1;
  if (make_tmap_f16_2d(&tdo, d_out_bf16, rows, hidden, hidden, Cfg::kBlk, kHeadDim)) return 1;
  BwdArgs a;
  a.seq = seq; a.hidden = hidden; a.lse2 = lse2; a.delta = delta;
  a.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv_bf16);
  a.scale = 0.125f;
  a.scale_log2 = 0.125f * 1.4426950408889634f;
  if (launch_bwd<MODE_DQ, Cfg>(t16, tbf, tdo, a, n_views, heads, stream)) return 1;
  return launch_bwd<MODE_DKV, Cfg>(t16, tbf, tdo, a, n_views, heads, stream);
}

}  // namespace

int attention_backward(const void* qkv_f16, const void* qkv_bf16, const void* d_out_bf16, const float* lse2,
                       const float* delta, void* dqkv_bf16, int n_views, int seq, int heads, cudaStream_t stream) {
  if (n_views <= 0) return 0;
  // A/B switch, read once per process: "64" = single-buffered 64-wide blocks, 4 row warps; "32" = two 32-wide buffers
  static const int variant = [] {
    const char* v = getenv("PG_ATTN_BWD");
    return (v && v[0] == '6') ? 64 : (v && v[0] == '3') ? 32 : 0;
  }();
  if (variant == 64)
    return run_backward<BwdCfg<64, 1>>(qkv_f16, qkv_bf16, d_out_bf16, lse2, delta, dqkv_bf16, n_views, seq, heads, stream);
  if (variant == 32)
    return run_backward<BwdCfg<32, 2>>(qkv_f16, qkv_bf16, d_out_bf16, lse2, delta, dqkv_bf16, n_views, seq, heads, stream);
  // default: 64-wide blocks, EIGHT row warps (two per TMEM lane quarter) -> 16 row warps per SM with two CTAs
  return run_backward<BwdCfg<64, 1, 8>>(qkv_f16, qkv_bf16, d_out_bf16, lse2, delta, dqkv_bf16, n_views, seq, heads, stream);
}

}  // namespace pg
