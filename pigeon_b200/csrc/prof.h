// Optional per-launch device timing (CUDA events on the launch stream), used by bench.py to attribute a step's
// time to kernel families.  Disabled by default: a ProfScope then costs one branch.
#pragma once
#include <cuda_runtime.h>

namespace pg {

bool prof_enabled();
void prof_record(const char* name, cudaEvent_t start, cudaEvent_t stop);

struct ProfScope {
  const char* name;
  cudaStream_t stream;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  ProfScope(const char* n, cudaStream_t s) : name(n), stream(s) {
    if (prof_enabled()) {
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      cudaEventRecord(e0, stream);
    }
  }
  ~ProfScope() {
    if (e0) {
      cudaEventRecord(e1, stream);
      prof_record(name, e0, e1);
    }
  }
};

}  // namespace pg
