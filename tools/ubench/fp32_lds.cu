// Micro-benchmarks behind the refiner tile scan design: FP32 FMA issue rates (FFMA vs FFMA2, register vs shared operands) and
// the cost of LDS.128 / LDS.64 / LDS.32 under the broadcast patterns the kernel uses.  One CTA per SM, `warps` warps per CTA,
// clock64() around the loop, result = operations per clock per SM.   nvcc -arch=sm_100a -O3 fp32_lds.cu -o fp32_lds
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d)
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
        "l"(*reinterpret_cast<unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}

constexpr int ITERS = 2048;

__global__ void k_ffma(float* out, long long* clk, float x, float y) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x + i;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], x, y);
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
// FFMA2 with three full 64-bit register operands
__global__ void k_ffma2(float* out, long long* clk, float x, float y) {
  float2 a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = make_float2(threadIdx.x + i, i);
  float2 xx = make_float2(x, x * 1.0001f), yy = make_float2(y, y * 0.999f);
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = ffma2(a[i], xx, yy);
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 16; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
// FFMA2 in the kernel's form: acc = (p, p) * (qa, qb) + acc with 8 accumulators sharing 2 p scalars and 4 q pairs
__global__ void k_ffma2_tile(float* out, long long* clk, float x, float y) {
  float2 acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x + i, i);
  float p0 = x, p1 = y;
  float2 q[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = make_float2(x + i, y - i);
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        acc[c] = ffma2(make_float2(p0, p0), q[c], acc[c]);
        acc[4 + c] = ffma2(make_float2(p1, p1), q[c], acc[4 + c]);
      }
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

// LDS patterns. mode: 0 all lanes one address | 1 lane>>2 selects one of 8 rows (16 B each, conflict-free) | 2 lane&3 selects one
// of 4 rows | 3 32 distinct conflict-free chunks | 4 lane>>3 selects (quarter-warp uniform) | 5 lane&7 selects one of 8
template <int W> __device__ __forceinline__ float lds(uint32_t addr);
template <> __device__ __forceinline__ float lds<4>(uint32_t addr) {
  float a, b, c, d;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "r"(addr));
  return a;
}
template <> __device__ __forceinline__ float lds<2>(uint32_t addr) {
  float a, b;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a), "=f"(b) : "r"(addr));
  return a;
}
template <> __device__ __forceinline__ float lds<1>(uint32_t addr) {
  float a;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a) : "r"(addr));
  return a;
}
template <int W>
__global__ void k_lds(float* out, long long* clk, int mode) {
  extern __shared__ float4 sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_float4(i, 1, 2, 3);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  int idx;
  switch (mode) {
    case 0: idx = 0; break;
    case 1: idx = (lane >> 2) * 193; break;       // 193 * 16 B = 3088 B row stride
    case 2: idx = (lane & 3) * 385; break;        // 385 * 16 B = 6160 B
    case 3: idx = lane; break;
    case 4: idx = (lane >> 3) * 193; break;
    default: idx = (lane & 7) * 193; break;
  }
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm + idx);
  float acc = 0.f;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += lds<W>(base + ((it + i) & 15) * 16);
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out; long long* clk;
  cudaMalloc(&out, sizeof(float) * sms * 1024);
  cudaMalloc(&clk, sizeof(long long) * sms);
  long long* h = new long long[sms];
  auto report = [&](const char* name, int warps, double ops_per_thread_iter) {
    cudaDeviceSynchronize();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    cudaMemcpy(h, clk, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < sms; ++i) c += h[i]; c /= sms;
    printf("%-34s warps/SM %2d: %8.2f per clk per SM (%.0f clk)\n", name, warps, ops_per_thread_iter * ITERS * warps * 32 / c, c);
  };
  for (int warps : {4, 8, 16, 32}) {
    for (int rep = 0; rep < 2; ++rep) k_ffma<<<sms, warps * 32>>>(out, clk, 1.0001f, 0.5f);
    report("FFMA (FMA lanes/clk/SM)", warps, 16);
    for (int rep = 0; rep < 2; ++rep) k_ffma2<<<sms, warps * 32>>>(out, clk, 1.0001f, 0.5f);
    report("FFMA2 3x64-bit regs (FMA/clk/SM)", warps, 32);
    for (int rep = 0; rep < 2; ++rep) k_ffma2_tile<<<sms, warps * 32>>>(out, clk, 1.0001f, 0.5f);
    report("FFMA2 (p,p)*(qa,qb) tile form", warps, 32);
  }
  const char* names[6] = {"one address", "8 rows by lane>>2", "4 rows by lane&3", "32 distinct", "4 rows by lane>>3", "8 rows by lane&7"};
  for (int warps : {8, 16}) {
    for (int mode = 0; mode < 6; ++mode) {
      char nm[96];
      for (int rep = 0; rep < 2; ++rep) k_lds<4><<<sms, warps * 32, 65536>>>(out, clk, mode);
      snprintf(nm, sizeof nm, "LDS.128 %s (instr/clk/SM)", names[mode]); report(nm, warps, 8.0 / 32);
      for (int rep = 0; rep < 2; ++rep) k_lds<2><<<sms, warps * 32, 65536>>>(out, clk, mode);
      snprintf(nm, sizeof nm, "LDS.64  %s (instr/clk/SM)", names[mode]); report(nm, warps, 8.0 / 32);
      for (int rep = 0; rep < 2; ++rep) k_lds<1><<<sms, warps * 32, 65536>>>(out, clk, mode);
      snprintf(nm, sizeof nm, "LDS.32  %s (instr/clk/SM)", names[mode]); report(nm, warps, 8.0 / 32);
    }
  }
  return 0;
}
