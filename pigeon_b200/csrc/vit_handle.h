// Shared by capi.cu and vit_train_capi.cu: the tower handle, workspace carving, SM-count cache.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <vector>

#include "../../include/pigeon_b200.h"
#include "tma_host.h"

struct pg_vit {
  pg_vit_config cfg;
  pg_vit_weights w;
  std::vector<pg_vit_layer> layers;
  int tokens;
  int grid_patches;
};

namespace pg {

int sm_count();   // cached cudaDevAttrMultiProcessorCount of the current device, -1 (and last_error) on failure

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct Carver {
  uint8_t* base;
  size_t off = 0;
  explicit Carver(void* p) : base(reinterpret_cast<uint8_t*>(p)) {}
  void* take(size_t bytes) {
    void* p = base ? base + off : nullptr;
    off += align_up(bytes, 1024);
    return p;
  }
};

}  // namespace pg
