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
// Persistent: 2 CTAs per SM (80 KB smem, 256 TMEM columns each), each looping over work items
// (q-tile of 128 rows, head, view) with every pipeline (TMA ring, S/P buffers, O buffers, Q buffers) running ACROSS
// tile boundaries, so the next tile's loads and first QK^T products overlap the current tile's tail.  192 threads:
//   warp 0     TMA producer: Q tiles (double-buffered), K/V tiles (64 x 64 halves, 128B swizzle) through a 6-slot ring
//   warp 1     TMEM allocator + MMA issuer (tcgen05.mma cta_group::1, M = 128)
//   warps 2-9  softmax: two warps per TMEM lane quarter, each thread owns half (32 columns) of one query row's block
//
// Single pass, flash-style, KV blocks of 64:  S_j = Q K_j^T lands in one of TWO TMEM buffers so that the MMAs of block
// j+1 run while the softmax warps work on block j.  The softmax warps keep a running row maximum m and row sum l,
// write P_j = exp2(S_j*c - m*c) as packed fp16 over the S_j columns, and the MMA warp accumulates O += P_j V_j with P
// as the TMEM A operand and V as an MN-major smem B operand.  m is only raised (and O, l rescaled in TMEM) when the
// block maximum exceeds it by more than 2^8 in the exponent domain ("lazy rescale"): P then stays <= 256, exact in fp16
// range, and the rescale path is rare.  Final O / l -> fp16.
#include "attention.h"
#include "prof.h"
#include "ptx.cuh"
#include "tma_host.h"

namespace pg {

namespace {

constexpr int kHeadDim = 64;
constexpr int kBlockQ = 128;
constexpr int kBlockKV = 64;
constexpr int kSlots = 6;
constexpr int kQBytes = kBlockQ * kHeadDim * 2;       // 16 KB
constexpr int kTileBytes = kBlockKV * kHeadDim * 2;   // 8 KB
constexpr int kThreads = 320;                         // TMA warp + MMA warp + 8 softmax warps
constexpr int kTmemCols = 256;                        // S0 [0,64)  S1 [64,128)  O0 [128,192)  O1 [192,256)
constexpr int kOCol = 128;
constexpr int kXchgBytes = (2 * 4 * 64 + 4 * 64) * 4;  // half-row maxima (double-buffered) + row sums
constexpr int kSmemBytes = 2 * kQBytes + kSlots * kTileBytes + kXchgBytes + 1024 + 256;
constexpr float kRescaleThreshold = 8.0f;             // log2 domain

struct AttnArgs {
  int n_views, heads, q_tiles;
  int seq;      // tokens per view (577)
  int hidden;   // heads * 64
  __half* out;  // [n_views*seq, hidden]
  float scale_log2;  // (1/sqrt(64)) * log2(e)
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
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

// Work item `w` -> (view, head, q_tile); q_tile fastest so that co-scheduled CTAs share the K/V of one (view, head).
struct Tile {
  int view, head, qt;
};
__device__ __forceinline__ Tile decode_tile(int w, const AttnArgs& a) {
  Tile t;
  t.qt = w % a.q_tiles;
  const int vh = w / a.q_tiles;
  t.head = vh % a.heads;
  t.view = vh / a.heads;
  return t;
}

__global__ void __launch_bounds__(kThreads, 2)
attention_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;                       // 2 x 16 KB
  uint8_t* smem_kv = smem + 2 * kQBytes;        // kSlots x 8 KB
  float* xchg_max = reinterpret_cast<float*>(smem + 2 * kQBytes + kSlots * kTileBytes);  // [2][4][2][32]
  float* xchg_sum = xchg_max + 2 * 4 * 64;                                                // [4][2][32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kQBytes + kSlots * kTileBytes + kXchgBytes);
  uint64_t* full_bar = bars;             // [kSlots]  TMA -> MMA
  uint64_t* empty_bar = bars + kSlots;   // [kSlots]  MMA -> TMA
  uint64_t* q_full = bars + 2 * kSlots;  // [2] TMA -> MMA : Q tile landed
  uint64_t* q_empty = q_full + 2;        // [2] MMA -> TMA : every QK^T of the tile retired
  uint64_t* s_full = q_full + 4;         // [2] MMA -> softmax : S block complete in TMEM buffer b
  uint64_t* p_ready = q_full + 6;        // [2] softmax -> MMA : P written over buffer b (4 warps arrive)
  uint64_t* pv_done = q_full + 8;        // MMA -> softmax : one phase per P V product (O is quiescent after it)
  uint64_t* o_full = q_full + 9;         // [2] MMA -> softmax : last P V of the tile retired, O buffer complete
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(q_full + 11);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = args.seq;
  const int nb = (S + kBlockKV - 1) / kBlockKV;                 // KV blocks per tile (10 for S = 577)
  const int last_valid = S - (nb - 1) * kBlockKV;               // valid kv columns in the last block (1)
  const int last_n = (last_valid + 15) & ~15;                   // MMA N / K extent of the last block (16)
  const int total = args.n_views * args.heads * args.q_tiles;
  const int my_tiles = (total > (int)blockIdx.x) ? (total - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int G = my_tiles * nb;                                  // KV blocks this CTA will process, globally numbered

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    for (int s = 0; s < kSlots; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&q_full[b], 1);
      mbar_init(&q_empty[b], 1);
      mbar_init(&s_full[b], 1);
      mbar_init(&p_ready[b], 8);
      mbar_init(&o_full[b], 1);
    }
    mbar_init(pv_done, 1);
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
    // Load order = consumption order of the MMA warp over the GLOBAL block index g = tile_i * nb + j:
    //   K(0), K(1), then for g = 0, 1, ...: V(g), K(g+2);  the Q tile of a tile goes right before its K(., 0).
    if (lane == 0 && G > 0) {
      int slot = 0;
      uint32_t phase = 0;
      auto load_kv = [&](int g, bool is_v) {
        const int ti = g / nb, j = g - ti * nb;
        const Tile t = decode_tile(blockIdx.x + ti * gridDim.x, args);
        const int row0 = t.view * S;
        const int q_col = t.head * kHeadDim;
        if (!is_v && j == 0) {  // first block of a tile: its Q tile (two 64-row boxes) into Q buffer ti & 1
          mbar_wait(&q_empty[ti & 1], ((ti >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&q_full[ti & 1], kQBytes);
          uint8_t* qb = smem_q + (ti & 1) * kQBytes;
          tma_load_2d(qb, &tmap_qkv, &q_full[ti & 1], q_col, row0 + t.qt * kBlockQ);
          tma_load_2d(qb + kTileBytes, &tmap_qkv, &q_full[ti & 1], q_col, row0 + t.qt * kBlockQ + kBlockKV);
        }
        mbar_wait(&empty_bar[slot], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[slot], kTileBytes);
        tma_load_2d(smem_kv + slot * kTileBytes, &tmap_qkv, &full_bar[slot],
                    (is_v ? 2 * args.hidden : args.hidden) + q_col, row0 + j * kBlockKV);
        if (++slot == kSlots) { slot = 0; phase ^= 1; }
      };
      load_kv(0, false);
      if (G > 1) load_kv(1, false);
      for (int g = 0; g < G; ++g) {
        load_kv(g, true);
        if (g + 2 < G) load_kv(g + 2, false);
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0 && G > 0) {
      int slot = 0;
      uint32_t phase = 0;

      auto issue_s = [&](int g) {
        const int ti = g / nb, j = g - ti * nb;
        const int n = (j == nb - 1) ? last_n : kBlockKV;
        const uint32_t idesc = make_idesc_f16(kBlockQ, n, 0, 0);
        const uint32_t s_tmem = tmem_base + (g & 1) * kBlockKV;
        if (j == 0) {
          mbar_wait(&q_full[ti & 1], (ti >> 1) & 1);
          tc_fence_after();
        }
        const uint32_t q_addr = smem_u32(smem_q + (ti & 1) * kQBytes);
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
        tc_commit(&s_full[g & 1]);
        if (j == nb - 1) tc_commit(&q_empty[ti & 1]);  // the tile's last read of its Q buffer
        if (++slot == kSlots) { slot = 0; phase ^= 1; }
      };
      auto issue_pv = [&](int g) {
        const int ti = g / nb, j = g - ti * nb;
        const int kext = (j == nb - 1) ? last_n : kBlockKV;              // contraction extent = kv rows of this block
        const uint32_t idesc = make_idesc_f16(kBlockQ, kHeadDim, 0, 1);  // B (= V) is MN-major
        const uint32_t p_tmem = tmem_base + (g & 1) * kBlockKV;          // P aliases the S buffer, fp16 pairs
        const uint32_t o_tmem = tmem_base + kOCol + (ti & 1) * kHeadDim;
        mbar_wait(&full_bar[slot], phase);
        tc_fence_after();
        const uint32_t v_addr = smem_u32(smem_kv + slot * kTileBytes);
        for (int k = 0; k < kext / 16; ++k) {
          // V tile: row = kv index (128 B each, 64 halves of head_dim), 8-row groups 1024 B apart.
          const uint64_t b_desc = make_smem_desc(v_addr + k * 16 * 128, 1024, 1024, kLayoutSw128);
          umma_ts(o_tmem, p_tmem + k * 8, b_desc, idesc, (j | k) != 0);
        }
        tc_commit(&empty_bar[slot]);
        tc_commit(pv_done);
        if (j == nb - 1) tc_commit(&o_full[ti & 1]);
        if (++slot == kSlots) { slot = 0; phase ^= 1; }
      };

      issue_s(0);
      if (G > 1) issue_s(1);
      for (int g = 0; g < G; ++g) {
        mbar_wait(&p_ready[g & 1], (g >> 1) & 1);
        tc_fence_after();
        issue_pv(g);
        if (g + 2 < G) issue_s(g + 2);  // overwrites P_g only after P_g V_g (in-order tensor pipe)
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax warps (8 per CTA)
    // Two warps share each TMEM lane quarter (= 32 query rows): warp "half 0" owns S/P/O columns [0,32), "half 1"
    // owns [32,64).  Per block they exchange their half-row maxima through shared memory and a 64-thread named
    // barrier, so both take identical rescale decisions; row sums are combined once per tile.
    // Software pipeline over the global block index g: while the exponentials of block g are being evaluated the
    // TMEM load of block g+1 is already in flight, and its row maxima are reduced while the P store of block g drains.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const uint32_t lane_base = uint32_t(q * 32) << 16;
    const float c = args.scale_log2;
    const float2 c2 = make_float2(c, c);
    const int bar_id = 1 + q;
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory"); };

    // per-tile state
    int ti = 0, j = 0;
    Tile t = decode_tile(blockIdx.x, args);
    float m = -INFINITY;   // reference maximum currently used in the exponent (raw logit units)
    float2 l01 = make_float2(0.f, 0.f), l23 = make_float2(0.f, 0.f);
    float bm = -INFINITY;  // row maximum of the block about to be processed (both halves)

    // masked maximum of this thread's 32 columns of block index jj
    auto block_max = [&](const uint32_t (&r)[32], int jj) -> float {
      float b0 = -INFINITY, b1 = -INFINITY, b2 = -INFINITY, b3 = -INFINITY;
      if (jj == nb - 1 && last_valid < kBlockKV) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const int col = half * 32 + i;
          b0 = fmaxf(b0, col + 0 < last_valid ? __uint_as_float(r[i]) : -INFINITY);
          b1 = fmaxf(b1, col + 1 < last_valid ? __uint_as_float(r[i + 1]) : -INFINITY);
          b2 = fmaxf(b2, col + 2 < last_valid ? __uint_as_float(r[i + 2]) : -INFINITY);
          b3 = fmaxf(b3, col + 3 < last_valid ? __uint_as_float(r[i + 3]) : -INFINITY);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          b0 = fmaxf(b0, __uint_as_float(r[i]));
          b1 = fmaxf(b1, __uint_as_float(r[i + 1]));
          b2 = fmaxf(b2, __uint_as_float(r[i + 2]));
          b3 = fmaxf(b3, __uint_as_float(r[i + 3]));
        }
      }
      return fmaxf(fmaxf(b0, b1), fmaxf(b2, b3));
    };
    // asynchronous TMEM loads of this thread's 32 columns of block g, in two 16-column halves (the second half is issued
    // once the registers of the block being consumed are free); both complete at the next tmem_ld_wait()
    auto fetch_lo = [&](uint32_t (&r)[32], int g) {
      mbar_wait(&s_full[g & 1], (g >> 1) & 1);
      tc_fence_after();
      tmem_ld16(tmem_base + lane_base + (g & 1) * kBlockKV + half * 32, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
    };
    auto fetch_hi = [&](uint32_t (&r)[32], int g) {
      tmem_ld16(tmem_base + lane_base + (g & 1) * kBlockKV + half * 32 + 16, *reinterpret_cast<uint32_t(*)[16]>(&r[16]));
    };
    auto exchange_max = [&](float mine, int g) -> float {   // all 64 threads of the pair
      float* slot = xchg_max + ((g & 1) * 4 + q) * 64;
      slot[half * 32 + lane] = mine;
      pair_sync();
      return fmaxf(mine, slot[(half ^ 1) * 32 + lane]);
    };

    // one pipeline step: consume block g from `cur`, prefetch block g+1 into `nxt`
    auto step = [&](uint32_t (&cur)[32], uint32_t (&nxt)[32], int g) {
      const uint32_t p_tmem = tmem_base + lane_base + (g & 1) * kBlockKV + half * 16;
      const uint32_t o_tmem = tmem_base + lane_base + kOCol + (ti & 1) * kHeadDim + half * 32;
      const bool tail = (j == nb - 1) && (last_valid < kBlockKV);
      const bool more = (g + 1 < G);

      // running maximum with lazy rescale (identical decision in both warps of the pair)
      if (j == 0) {
        m = bm;
        l01 = make_float2(0.f, 0.f);
        l23 = make_float2(0.f, 0.f);
      } else {
        const float m_new = fmaxf(m, bm);
        const bool need = (m_new - m) * c > kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {
          // rare: raise m for every row of this warp and rescale its half of the O rows and l.  O is quiescent once
          // P_{g-1} V_{g-1} retired (S_g complete => every product before g-1 retired, so the parity wait below
          // cannot alias an older phase), and P_g V_g cannot be issued before this warp reports p_ready.
          mbar_wait(pv_done, (g - 1) & 1);
          tc_fence_after();
          const float alpha = ex2((m - m_new) * c);
          uint32_t o[32];
          tmem_ld32(o_tmem, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st32_(o_tmem, o);
          tmem_st_wait();
          l01.x *= alpha; l01.y *= alpha; l23.x *= alpha; l23.y *= alpha;
          m = m_new;
        }
      }
      const float mc = m * c;
      const float2 nmc2 = make_float2(-mc, -mc);

      // P is packed in place: columns i..i+3 are consumed before cur[i/2], cur[i/2+1] (<= i) are overwritten
      auto exps = [&](int i0) {   // 16 columns starting at i0
#pragma unroll
        for (int i = i0; i < i0 + 16; i += 4) {
          const float2 x01 = ffma2(make_float2(__uint_as_float(cur[i]), __uint_as_float(cur[i + 1])), c2, nmc2);
          const float2 x23 = ffma2(make_float2(__uint_as_float(cur[i + 2]), __uint_as_float(cur[i + 3])), c2, nmc2);
          float2 p01 = make_float2(ex2(x01.x), ex2(x01.y));
          float2 p23 = make_float2(ex2(x23.x), ex2(x23.y));
          if (tail) {
            const int col = half * 32 + i;
            if (col + 0 >= last_valid) p01.x = 0.f;
            if (col + 1 >= last_valid) p01.y = 0.f;
            if (col + 2 >= last_valid) p23.x = 0.f;
            if (col + 3 >= last_valid) p23.y = 0.f;
          }
          l01 = fadd2(l01, p01);
          l23 = fadd2(l23, p23);
          cur[i >> 1] = pack_half2(p01.x, p01.y);
          cur[(i >> 1) + 1] = pack_half2(p23.x, p23.y);
        }
      };
      exps(0);
      if (more) fetch_lo(nxt, g + 1);  // S_{g+1} is normally complete by now: its load overlaps the second half
      exps(16);
      if (more) fetch_hi(nxt, g + 1);
      tmem_st16(p_tmem, *reinterpret_cast<uint32_t(*)[16]>(&cur[0]));  // P over S columns both warps of the pair hold in registers
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[g & 1]);
      if (more) {
        tmem_ld_wait();
        // the exchange also orders both warps' S_{g+1} loads before any P_{g+1} store
        bm = exchange_max(block_max(nxt, (j + 1 == nb) ? 0 : j + 1), g + 1);
      }

      if (j == nb - 1) {
        // tile epilogue: O / l -> fp16 -> global.  The O buffer of tile ti is only rewritten by tile ti + 2, whose
        // first P V needs this warp's p_ready, i.e. comes after this read.
        const float l_own = (l01.x + l01.y) + (l23.x + l23.y);
        xchg_sum[q * 64 + half * 32 + lane] = l_own;
        pair_sync();
        const float l_tot = l_own + xchg_sum[q * 64 + (half ^ 1) * 32 + lane];
        mbar_wait(&o_full[ti & 1], (ti >> 1) & 1);
        tc_fence_after();
        const float inv_l = 1.0f / l_tot;
        const int q_row = t.qt * kBlockQ + q * 32 + lane;  // token index inside the view
        __half* orow = args.out + (size_t)(t.view * S + q_row) * args.hidden + t.head * kHeadDim + half * 32;
        uint32_t o[32];
        tmem_ld32(o_tmem, o);
        tmem_ld_wait();
        if (q_row < S) {
          uint4* o4 = reinterpret_cast<uint4*>(orow);
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
        tc_fence_before();  // order this tile's TMEM reads before the barrier arrivals of the next tile
        ++ti;
        j = 0;
        if (more) t = decode_tile(blockIdx.x + ti * gridDim.x, args);
      } else {
        ++j;
      }
    };

    if (G > 0) {
      uint32_t ra[32], rb[32];
      fetch_lo(ra, 0);
      fetch_hi(ra, 0);
      tmem_ld_wait();
      bm = exchange_max(block_max(ra, 0), 0);
      for (int g = 0; g < G; g += 2) {
        step(ra, rb, g);
        if (g + 1 < G) step(rb, ra, g + 1);
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

}  // namespace

int attention_f16(const void* qkv, void* out, int n_views, int seq, int heads, cudaStream_t stream) {
  if (n_views <= 0) return 0;
  const int hidden = heads * kHeadDim;
  CUtensorMap tm;
  if (make_tmap_f16_2d(&tm, qkv, (uint64_t)n_views * seq, 3 * hidden, 3 * hidden, kBlockKV, kHeadDim)) return 1;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) { set_last_error("attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 1; }
    attr_set = true;
  }
  AttnArgs a;
  a.n_views = n_views;
  a.heads = heads;
  a.q_tiles = (seq + kBlockQ - 1) / kBlockQ;
  a.seq = seq;
  a.hidden = hidden;
  a.out = reinterpret_cast<__half*>(out);
  a.scale_log2 = 0.125f * 1.4426950408889634f;
  const long total = (long)n_views * heads * a.q_tiles;
  int sms = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long grid = 2L * (sms > 0 ? sms : 148);   // persistent: two co-resident CTAs per SM
  if (grid > total) grid = total;
  ProfScope prof("attention", stream);
  attention_kernel<<<(unsigned)grid, kThreads, kSmemBytes, stream>>>(tm, a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_last_error("attention launch: %s", cudaGetErrorString(e)); return 1; }
  return 0;
}

}  // namespace pg
