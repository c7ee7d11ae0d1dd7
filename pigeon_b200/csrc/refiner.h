// Internal (C++) interface of the ProtoRefiner retrieval kernels; see refiner.cu.
#pragma once
#include <cuda_runtime.h>

namespace pg {

// CSR-packed prototype bank, all pointers device memory (built once from the reference's
// per-cell HF datasets: proto_refiner.py:257-313,359-384).
struct RefinerBank {
  int num_cells;               // C
  int dim;                     // D (multiple of 128)
  const long long* cell_off;   // [C+1]  prototype range of each geocell; empty range == "protos[cell] is None"
  const float* proto_emb;      // [P, D] prototype = mean of member embeddings
  const float* proto_lnglat;   // [P, 2] cluster (lng, lat), fp32 like the HF torch formatter yields
  const int* proto_count;      // [P]    members in the cluster
  const long long* member_off; // [P+1]  range into member_idx
  const long long* member_idx; // [...]  rows of data_emb / data_lnglat
  const float* data_emb;       // [Ntrain, D] training embeddings (4-view mean already applied)
  const float* data_lnglat;    // [Ntrain, 2] training labels (lng, lat) fp32
  const float* proto_sqnorm;   // [P] |prototype|^2 (refiner_bank_sqnorm), or nullptr: the tile scan needs it
};

int refiner_pool(const float* emb, float* q, long B, int V, int D, cudaStream_t stream);
int refiner_scan(const RefinerBank& bank, const float* q, const long long* cand, int cand_stride, long B, int topk,
                 float* best_logit, float* best_lnglat, int* best_proto, int num_sms, cudaStream_t stream);
// Cell-major variant of refiner_scan: pairs counting-sorted by geocell, each touched prototype segment read once.
size_t refiner_sort_workspace_bytes(int num_cells, long pairs);
int refiner_scan_cell_major(const RefinerBank& bank, const float* q, const long long* cand, int cand_stride, long B,
                            int topk, void* sort_ws, float* best_logit, float* best_lnglat, int* best_proto,
                            int num_sms, cudaStream_t stream);
// Tile scan (v4): same contract as refiner_scan_cell_major, needs bank.proto_sqnorm.  Persistent CTAs stream 16-prototype
// tiles through a shared-memory ring (cp.async.bulk) and score them against up to 32 staged queries of the geocell.
int refiner_scan_tiles(const RefinerBank& bank, const float* q, const long long* cand, int cand_stride, long B, int topk,
                       void* sort_ws, float* best_logit, float* best_lnglat, int* best_proto, int num_sms,
                       cudaStream_t stream);
// Slab scan (v5): 8 x 8 register tiles, TMA-streamed embedding chunks; for banks with hundreds of prototypes per geocell.
// num_protos = rows of bank.proto_emb (for the tensor map).
int refiner_scan_slabs(const RefinerBank& bank, long num_protos, const float* q, const long long* cand, int cand_stride,
                       long B, int topk, void* sort_ws, float* best_logit, float* best_lnglat, int* best_proto,
                       int num_sms, cudaStream_t stream);
// |p|^2 of every prototype row (once per bank)
int refiner_bank_sqnorm(const float* proto_emb, long P, int D, float* out, int num_sms, cudaStream_t stream);
// data_views [N, V, D] -> data_mean [N, D] (view mean) and proto_emb [P, D] (mean of member rows of data_mean)
int bank_build(const float* data_views, long N, int V, int D, const long long* member_off, const long long* member_idx,
               long P, float* data_mean, float* proto_emb, int num_sms, cudaStream_t stream);
int refiner_finalize(const float* best_logit, const float* best_lnglat, const long long* cand, const float* cand_prob,
                     int cand_stride, const double* init_lnglat, long B, int topk, float temperature,
                     double max_refinement, float* out_lnglat, long long* out_cell, int* out_choice,
                     cudaStream_t stream);

}  // namespace pg
