// Persistent warp-specialised tcgen05 GEMM for sm_100a:  D[M,N] = A[M,K] * W[N,K]^T  (+ epilogue)
//
//   A : fp16 row-major [M, K]  (activations, K contiguous  -> "K-major" UMMA operand)
//   W : fp16 row-major [N, K]  (nn.Linear weight layout    -> "K-major" UMMA operand)
//   accumulate fp32 in TMEM, epilogue variants below.
//
// Replaces the cuBLAS sgemm calls behind HF CLIPAttention.{q,k,v,out}_proj / CLIPMLP.{fc1,fc2} and the
// cuDNN patch conv reached from reference models/clip_embedder.py:63 and models/super_guessr.py:395,
// and nn.Linear cell_layer at models/super_guessr.py:447.
//
// CTA = 256 threads, 1 CTA / SM, grid = min(#SM, #tiles):
//   warp 0      TMA producer   (one lane)   global -> smem ring, 128B swizzle
//   warp 1      MMA issuer     (one lane)   tcgen05.mma cta_group::1, M=128, N=BLOCK_N, K=16
//   warp 2      TMEM allocator
//   warps 4..7  epilogue: tcgen05.ld -> registers -> bias/activation/residual -> global
// Two TMEM accumulator stages (2 x BLOCK_N columns) so tile i's epilogue overlaps tile i+1's MMAs.
#include "gemm.h"
#include "gemm_epilogue.cuh"
#include "ptx.cuh"
#include "prof.h"
#include "tma_host.h"

#include <atomic>
#include <stdlib.h>

namespace pg {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 fp16 = 128 bytes = one swizzle atom row
constexpr int UMMA_K = 16;
constexpr int kNumThreads = 256;
constexpr int kEpiWarp0 = 4;

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int kStages = (BLOCK_N == 256) ? 4 : 6;
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BLOCK_N;  // 512 or 256: powers of two
  static constexpr int kSmemBytes = kStages * kStageBytes + 4 * kStageWarpBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_f16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const GemmArgs args) {
  using Cfg = GemmCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled tiles need 1024-byte alignment.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + Cfg::kStages * Cfg::kABytes;
  uint8_t* smem_stage = smem + Cfg::kStages * Cfg::kStageBytes;  // epilogue staging, 4 warps
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stage + 4 * kStageWarpBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::kStages;
  uint64_t* tmem_full_bar = bars + 2 * Cfg::kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_blocks = (args.M + BLOCK_M - 1) / BLOCK_M;
  const int n_blocks = (args.N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = m_blocks * n_blocks;
  const int num_kb = (args.K + BLOCK_K - 1) / BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 4);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_smem, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_blocks, n_blk = tile % n_blocks;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          tma_load_2d(smem_a + stage * Cfg::kABytes, &tmap_a, &full_bar[stage], kb * BLOCK_K, m_blk * BLOCK_M);
          tma_load_2d(smem_b + stage * Cfg::kBBytes, &tmap_b, &full_bar[stage], kb * BLOCK_K, n_blk * BLOCK_N);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(BLOCK_M, BLOCK_N, 0, 0) | args.idesc_fmt;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + stage * Cfg::kABytes);
          const uint32_t b_addr = smem_u32(smem_b + stage * Cfg::kBBytes);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // K-major, 128B swizzle: 8-row groups are 1024 B apart; a K step of 16 halves = +32 B.
            const uint64_t a_desc = make_smem_desc(a_addr + k * UMMA_K * 2, 16, 1024, kLayoutSw128);
            const uint64_t b_desc = make_smem_desc(b_addr + k * UMMA_K * 2, 16, 1024, kLayoutSw128);
            umma_ss(d_tmem, a_desc, b_desc, idesc, (kb | k) != 0);
          }
          tc_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tmem_full_bar[acc]);  // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ------------------------------------------------------------------ epilogue (gemm_epilogue.cuh)
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    uint8_t* stage = smem_stage + (warp - kEpiWarp0) * kStageWarpBytes;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / n_blocks, n_blk = tile % n_blocks;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + acc * BLOCK_N;
      epilogue_tile<EPI>(args, t_row, stage, m_blk * BLOCK_M + q * 32, n_blk * BLOCK_N, 0, BLOCK_N, lane);
      // release the accumulator stage back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int BLOCK_N, int EPI>
int launch(const GemmProblem& p, int num_sms, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N>;
  constexpr bool f16_out = EpiTraits<EPI>::kF16Out || EPI == EPI_BF16_DGELU;   // 2-byte output elements
  CUtensorMap ta, tb;
  if (make_tmap_f16_2d(&ta, p.a, p.M, p.K, p.lda, BLOCK_M, BLOCK_K)) return 1;
  if (make_tmap_f16_2d(&tb, p.w, p.N, p.K, p.ldw, BLOCK_N, BLOCK_K)) return 1;
  auto kern = gemm_f16_kernel<BLOCK_N, EPI>;
  static std::atomic<int> attr_set[64];   // per template instantiation and device (the attribute is per context)
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!attr_set[dev].load(std::memory_order_acquire)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) { set_last_error("gemm: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 1; }
    attr_set[dev].store(1, std::memory_order_release);
  }
  GemmArgs a;
  a.M = p.M; a.N = p.N; a.K = p.K; a.out = p.out; a.ldo = p.ldo; a.bias = p.bias;
  a.vec_ok = (p.ldo % (f16_out ? 8 : 4)) == 0;
  a.rowmap_div = p.rowmap_div > 0 ? p.rowmap_div : 1; a.rowmap_mul = p.rowmap_mul; a.rowmap_add = p.rowmap_add;
  // A and B of one type: tcgen05 kind::f16 traps (illegal instruction) on a bf16 x fp16 pair
  a.idesc_fmt = p.operand_bf16 ? ((1u << 7) | (1u << 10)) : 0u;
  a.resid = p.resid ? p.resid : reinterpret_cast<const float*>(p.out);
  a.aux = p.aux;
  a.mn_major = 0; a.mn_lbo = a.mn_sbo = 0;
  if ((EPI == EPI_BF16_DGELU || EPI == EPI_F16_BIAS_QGELU_SAVE || EPI == EPI_F32_BIAS_RESID_STATS) &&
      (!p.aux || (reinterpret_cast<uintptr_t>(p.aux) & 15))) {
    set_last_error("gemm: this epilogue needs a 16-byte aligned aux buffer"); return 1;
  }
  a.stats = p.stats; a.colsum = p.colsum; a.ln_eps = p.ln_eps; a.stats_slots = 0;
  if (EpiTraits<EPI>::kStats) {
    if (!p.stats || (p.N % 128) || !a.vec_ok || (p.ldo % 8)) { set_last_error("gemm: LayerNorm-producer epilogue needs stats, N %% 128 == 0 and ldo %% 8 == 0"); return 1; }
    a.stats_slots = p.N / 128;
  }
  if (EpiTraits<EPI>::kLn) {
    if (!p.stats || !p.colsum || !p.bias || (p.K % 128) || (p.N % 64) || (reinterpret_cast<uintptr_t>(p.colsum) & 15)) {
      set_last_error("gemm: LayerNorm-consumer epilogue needs stats, a bias, a 16-byte aligned colsum, K %% 128 == 0 and N %% 64 == 0"); return 1;
    }
    a.stats_slots = p.K / 128;
  }
  const int m_blocks = (p.M + BLOCK_M - 1) / BLOCK_M, n_blocks = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int tiles = m_blocks * n_blocks;
  const int grid = tiles < num_sms ? tiles : num_sms;
  static const char* const kNames[] = {"gemm_f16_bias", "gemm_f16_bias_qgelu", "gemm_f32_bias_resid", "gemm_f32_bias", "gemm_f32_rowmap",
                                       "gemm_bf16_dgelu", "gemm_f16_bias_qgelu_save", "gemm_f32_bias_resid_stats", "gemm_f16_ln_bias",
                                       "gemm_f16_ln_bias_qgelu"};
  ProfScope prof(kNames[EPI], stream);
  kern<<<grid, kNumThreads, Cfg::kSmemBytes, stream>>>(ta, tb, a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_last_error("gemm launch: %s", cudaGetErrorString(e)); return 1; }
  return 0;
}

}  // namespace

int gemm2_f16(const GemmProblem& p, int num_sms, cudaStream_t stream);  // gemm2_tcgen05.cu

static bool use_2cta() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PG_GEMM_1CTA");  // debugging / A-B switch: force the single-CTA kernel
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

int gemm_f16(const GemmProblem& p, int num_sms, cudaStream_t stream) {
  if (p.M <= 0 || p.N <= 0 || p.K <= 0) return 0;
  if ((p.lda % 8) || (p.ldw % 8)) { set_last_error("gemm: lda/ldw must be multiples of 8 halves (TMA 16B stride)"); return 1; }
  if ((reinterpret_cast<uintptr_t>(p.a) | reinterpret_cast<uintptr_t>(p.w) | reinterpret_cast<uintptr_t>(p.out) |
       reinterpret_cast<uintptr_t>(p.bias)) & 15) {
    set_last_error("gemm: pointers must be 16-byte aligned"); return 1;
  }
  // Wide tiles when N is a multiple of 256 (all ViT-L projections); 128-wide otherwise (head, small tests).
  const bool wide = (p.N % 256 == 0);
  if (p.mn_major) {
    if (!(wide && p.M > 128 && num_sms >= 2)) { set_last_error("gemm: MN-major operands need N %% 256 == 0 and M > 128"); return 1; }
    return gemm2_f16(p, num_sms, stream);
  }
  if (wide && p.M > 128 && num_sms >= 2 && use_2cta()) return gemm2_f16(p, num_sms, stream);
  switch (p.epi) {
    case EPI_F16_BIAS:       return wide ? launch<256, EPI_F16_BIAS>(p, num_sms, stream)       : launch<128, EPI_F16_BIAS>(p, num_sms, stream);
    case EPI_F16_BIAS_QGELU: return wide ? launch<256, EPI_F16_BIAS_QGELU>(p, num_sms, stream) : launch<128, EPI_F16_BIAS_QGELU>(p, num_sms, stream);
    case EPI_F32_BIAS_RESID: return wide ? launch<256, EPI_F32_BIAS_RESID>(p, num_sms, stream) : launch<128, EPI_F32_BIAS_RESID>(p, num_sms, stream);
    case EPI_F32_BIAS:       return wide ? launch<256, EPI_F32_BIAS>(p, num_sms, stream)       : launch<128, EPI_F32_BIAS>(p, num_sms, stream);
    case EPI_F32_ROWMAP:     return wide ? launch<256, EPI_F32_ROWMAP>(p, num_sms, stream)     : launch<128, EPI_F32_ROWMAP>(p, num_sms, stream);
    case EPI_BF16_DGELU:     return wide ? launch<256, EPI_BF16_DGELU>(p, num_sms, stream)     : launch<128, EPI_BF16_DGELU>(p, num_sms, stream);
    case EPI_F16_BIAS_QGELU_SAVE:
      return wide ? launch<256, EPI_F16_BIAS_QGELU_SAVE>(p, num_sms, stream) : launch<128, EPI_F16_BIAS_QGELU_SAVE>(p, num_sms, stream);
    case EPI_F32_BIAS_RESID_STATS:
      return wide ? launch<256, EPI_F32_BIAS_RESID_STATS>(p, num_sms, stream) : launch<128, EPI_F32_BIAS_RESID_STATS>(p, num_sms, stream);
    case EPI_F16_LN_BIAS:
      return wide ? launch<256, EPI_F16_LN_BIAS>(p, num_sms, stream) : launch<128, EPI_F16_LN_BIAS>(p, num_sms, stream);
    case EPI_F16_LN_BIAS_QGELU:
      return wide ? launch<256, EPI_F16_LN_BIAS_QGELU>(p, num_sms, stream) : launch<128, EPI_F16_LN_BIAS_QGELU>(p, num_sms, stream);
    default: set_last_error("gemm: unknown epilogue %d", p.epi); return 1;
  }
}

}  // namespace pg
