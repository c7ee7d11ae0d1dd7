// 2-CTA (cta_group::2) persistent tcgen05 GEMM for sm_100a:  D[M,N] = A[M,K] * W[N,K]^T (+ epilogue), N % 256 == 0.
//
// A CTA pair (cluster of 2 on one TPC) owns a 256 x 256 output tile: CTA r holds rows [256*m + 128*r, +128) of A and
// rows [256*n + 128*r, +128) of W in its shared memory; ONE thread of the leader CTA issues
// tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16) which makes both tensor cores multiply their own A half by
// BOTH W halves.  Versus the 1-CTA kernel each SM stages and reads half as much W: 8 KB instead of 12 KB of operand
// reads per 128-cycle MMA and 32 KB instead of 48 KB of TMA writes per k-block, and 6 pipeline stages fit instead of 4.
//
//   warp 0      TMA producer (both CTAs; completion bytes land on the LEADER's full barrier)
//   warp 1      MMA issuer (leader CTA only); tcgen05.commit multicasts to both CTAs' barriers
//   warp 2      TMEM allocator (both CTAs, cta_group::2)
//   warps 4..11 epilogue on the CTA's own 128 x 256 accumulator half: two warps per TMEM lane quarter, each taking
//               128 of the 256 columns (gemm_epilogue.cuh)
#include "gemm.h"
#include "gemm_epilogue.cuh"
#include "prof.h"
#include "ptx.cuh"
#include "tma_host.h"

#include <atomic>

namespace pg {

namespace {

constexpr int BLOCK_M_CTA = 128;
constexpr int BLOCK_N = 256;
constexpr int BLOCK_N_CTA = 128;
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int kNumThreads = 384;   // 4 role warps + 8 epilogue warps
constexpr int kNumEpiWarps = 8;
constexpr int kEpiWarp0 = 4;
constexpr int kStages = 6;
constexpr int kABytes = BLOCK_M_CTA * BLOCK_K * 2;  // 16 KB
constexpr int kBBytes = BLOCK_N_CTA * BLOCK_K * 2;  // 16 KB
constexpr int kStageBytes = kABytes + kBBytes;      // per CTA
constexpr int kTmemCols = 512;                      // two 256-column accumulator stages
constexpr int kSmemBytes = kStages * kStageBytes + kNumEpiWarps * kStageWarpBytes + 1024 + 256;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `smem_addr` (a shared::cta address) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default semantics (.release at CTA scope), as CUTLASS ClusterBarrier::arrive(cta_id): the data this arrive orders is
  // in TMEM / smem and is covered by tcgen05 fences and the async proxy; a cluster-scope release would add a heavy fence.
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr,
                                                int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
// same, with an L2 eviction policy for the lines the load touches
__device__ __forceinline__ void tma_load_2d_2sm_hint(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr,
                                                     int32_t c0, int32_t c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_ss_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at this smem offset in BOTH CTAs of the pair once the issued MMAs retired
__device__ __forceinline__ void tc_commit_2sm(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kNumThreads, 1)
gemm2_f16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const GemmArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * kABytes;
  uint8_t* smem_stage = smem + kStages * kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stage + kNumEpiWarps * kStageWarpBytes);
  uint64_t* full_bar = bars;                       // leader's copy is the live one
  uint64_t* empty_bar = bars + kStages;            // per CTA
  uint64_t* tmem_full_bar = bars + 2 * kStages;    // per CTA
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;    // leader's copy is the live one
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  const int m_blocks = (args.M + 2 * BLOCK_M_CTA - 1) / (2 * BLOCK_M_CTA);
  const int n_blocks = args.N / BLOCK_N;
  const int num_tiles = m_blocks * n_blocks;
  const int num_kb = (args.K + BLOCK_K - 1) / BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 2);   // leader's arrive.expect_tx + peer's remote arrive
      mbar_init(&empty_bar[s], 1);  // multicast tcgen05.commit
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);   // multicast tcgen05.commit
      mbar_init(&tmem_empty_bar[s], 2 * kNumEpiWarps);  // 8 epilogue warps x 2 CTAs
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_2sm(tmem_ptr_smem, kTmemCols);
  tc_fence_before();
  cluster_sync_all();  // peer barriers initialised and both TMEM allocations done before anything crosses CTAs
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // A is read by the n_blocks CTA pairs that share its 256 rows (they run concurrently, consecutive tile indices): keep its
      // lines in L2 over the streaming traffic of the epilogues
      const uint64_t pol_a = l2_policy_evict_last();
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int m_blk = tile / n_blocks, n_blk = tile % n_blocks;
        const int a_row = m_blk * 2 * BLOCK_M_CTA + rank * BLOCK_M_CTA;
        const int b_row = n_blk * BLOCK_N + rank * BLOCK_N_CTA;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          const uint32_t full_leader = mapa(smem_u32(&full_bar[stage]), 0);
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * kStageBytes);
          else        mbar_arrive_cluster(full_leader);
          if (!args.mn_major) {
            if (n_blocks > 1) tma_load_2d_2sm_hint(smem_a + stage * kABytes, &tmap_a, full_leader, kb * BLOCK_K, a_row, pol_a);
            else tma_load_2d_2sm(smem_a + stage * kABytes, &tmap_a, full_leader, kb * BLOCK_K, a_row);
            tma_load_2d_2sm(smem_b + stage * kBBytes, &tmap_b, full_leader, kb * BLOCK_K, b_row);
          } else {
            // operand tiles [64 contraction rows x 64 M/N columns] (128 B per row, swizzled): two per operand and CTA
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              tma_load_2d_2sm(smem_a + stage * kABytes + c * (kABytes / 2), &tmap_a, full_leader, a_row + c * 64, kb * BLOCK_K);
              tma_load_2d_2sm(smem_b + stage * kBBytes + c * (kBBytes / 2), &tmap_b, full_leader, b_row + c * 64, kb * BLOCK_K);
            }
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader && lane == 0) {
      const uint32_t mn = args.mn_major ? 1u : 0u;
      const uint32_t idesc = make_idesc_f16(2 * BLOCK_M_CTA, BLOCK_N, mn, mn) | args.idesc_fmt;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + stage * kABytes);
          const uint32_t b_addr = smem_u32(smem_b + stage * kBBytes);
          if (!mn) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              const uint64_t a_desc = make_smem_desc(a_addr + k * UMMA_K * 2, 16, 1024, kLayoutSw128);
              const uint64_t b_desc = make_smem_desc(b_addr + k * UMMA_K * 2, 16, 1024, kLayoutSw128);
              umma_ss_2sm(d_tmem, a_desc, b_desc, idesc, (kb | k) != 0);
            }
          } else {
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {   // 16 contraction rows of 128 B per step
              const uint64_t a_desc = make_smem_desc(a_addr + k * UMMA_K * 128, args.mn_lbo, args.mn_sbo, kLayoutSw128);
              const uint64_t b_desc = make_smem_desc(b_addr + k * UMMA_K * 128, args.mn_lbo, args.mn_sbo, kLayoutSw128);
              umma_ss_2sm(d_tmem, a_desc, b_desc, idesc, (kb | k) != 0);
            }
          }
          tc_commit_2sm(&empty_bar[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        tc_commit_2sm(&tmem_full_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ------------------------------------------------------------------ epilogue (both CTAs, own 128 rows)
    const int q = warp & 3;
    uint8_t* stage_buf = smem_stage + (warp - kEpiWarp0) * kStageWarpBytes;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int m_blk = tile / n_blocks, n_blk = tile % n_blocks;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + acc * BLOCK_N;
      // warps 4-7 take tile columns [0,128), warps 8-11 take [128,256) of the same lane quarters
      const int c_begin = (warp - kEpiWarp0) >= 4 ? BLOCK_N / 2 : 0;
      epilogue_tile<EPI>(args, t_row, stage_buf, m_blk * 2 * BLOCK_M_CTA + rank * BLOCK_M_CTA + q * 32, n_blk * BLOCK_N,
                         c_begin, c_begin + BLOCK_N / 2, lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa(smem_u32(&tmem_empty_bar[acc]), 0));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncwarp();        // role lanes rejoin their warp before the aligned cluster barrier
  cluster_sync_all();  // the pair's MMAs, commits and remote arrives are all done
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, kTmemCols);
  }
}

template <int EPI>
int launch2(const GemmProblem& p, int num_sms, cudaStream_t stream) {
  constexpr bool f16_out = EpiTraits<EPI>::kF16Out || EPI == EPI_BF16_DGELU;   // 2-byte output elements
  CUtensorMap ta, tb;
  if (p.mn_major) {   // tensors are [K, M] and [K, N]; box = 64 contraction rows x 64 columns
    if (p.M % 64) { set_last_error("gemm2: MN-major operands need M %% 64 == 0"); return 1; }
    if (make_tmap_f16_2d(&ta, p.a, p.K, p.M, p.lda, BLOCK_K, 64)) return 1;
    if (make_tmap_f16_2d(&tb, p.w, p.K, p.N, p.ldw, BLOCK_K, 64)) return 1;
  } else {
    if (make_tmap_f16_2d(&ta, p.a, p.M, p.K, p.lda, BLOCK_M_CTA, BLOCK_K)) return 1;
    if (make_tmap_f16_2d(&tb, p.w, p.N, p.K, p.ldw, BLOCK_N_CTA, BLOCK_K)) return 1;
  }
  auto kern = gemm2_f16_kernel<EPI>;
  static std::atomic<int> attr_set[64];   // per template instantiation and device (the attribute is per context)
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!attr_set[dev].load(std::memory_order_acquire)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) { set_last_error("gemm2: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 1; }
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
  a.mn_major = p.mn_major ? 1 : 0;
  a.mn_lbo = 8192; a.mn_sbo = 1024;   // chunk stride along M/N (64 columns = one swizzled 128-byte row set), 8-row group stride
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
  const int tiles = ((p.M + 255) / 256) * (p.N / BLOCK_N);
  int pairs = num_sms / 2;
  if (tiles < pairs) pairs = tiles;
  static const char* const kNames[] = {"gemm_f16_bias", "gemm_f16_bias_qgelu", "gemm_f32_bias_resid", "gemm_f32_bias",
                                       "gemm_f32_rowmap", "gemm_bf16_dgelu", "gemm_f16_bias_qgelu_save", "gemm_f32_bias_resid_stats", "gemm_f16_ln_bias",
                                       "gemm_f16_ln_bias_qgelu"};
  ProfScope prof(kNames[EPI], stream);
  kern<<<2 * pairs, kNumThreads, kSmemBytes, stream>>>(ta, tb, a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_last_error("gemm2 launch: %s", cudaGetErrorString(e)); return 1; }
  return 0;
}

}  // namespace

int gemm2_f16(const GemmProblem& p, int num_sms, cudaStream_t stream) {
  switch (p.epi) {
    case EPI_F16_BIAS:       return launch2<EPI_F16_BIAS>(p, num_sms, stream);
    case EPI_F16_BIAS_QGELU: return launch2<EPI_F16_BIAS_QGELU>(p, num_sms, stream);
    case EPI_F32_BIAS_RESID: return launch2<EPI_F32_BIAS_RESID>(p, num_sms, stream);
    case EPI_F32_BIAS:       return launch2<EPI_F32_BIAS>(p, num_sms, stream);
    case EPI_F32_ROWMAP:     return launch2<EPI_F32_ROWMAP>(p, num_sms, stream);
    case EPI_BF16_DGELU:     return launch2<EPI_BF16_DGELU>(p, num_sms, stream);
    case EPI_F16_BIAS_QGELU_SAVE: return launch2<EPI_F16_BIAS_QGELU_SAVE>(p, num_sms, stream);
    case EPI_F32_BIAS_RESID_STATS: return launch2<EPI_F32_BIAS_RESID_STATS>(p, num_sms, stream);
    case EPI_F16_LN_BIAS:          return launch2<EPI_F16_LN_BIAS>(p, num_sms, stream);
    case EPI_F16_LN_BIAS_QGELU:    return launch2<EPI_F16_LN_BIAS_QGELU>(p, num_sms, stream);
    default: set_last_error("gemm2: unknown epilogue %d", p.epi); return 1;
  }
}

}  // namespace pg
