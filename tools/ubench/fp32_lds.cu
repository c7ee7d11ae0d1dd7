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
  __shared__ float4 sm[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_float4(i, 1, 2, 3);
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

// The tile scan's inner step in isolation: 2 prototype LDS.128 (8 rows by lane >> 2) + 2 NCH query LDS.128 (4 rows by lane & 3)
// feeding 8 NCH FFMA2, software-pipelined one step ahead like the kernel.  Reports FMA lanes per clk per SM.
template <int NCH>
__global__ void k_mixed(float* out, long long* clk) {
  __shared__ float4 sm[3072];
  for (int i = threadIdx.x; i < 3072; i += blockDim.x) sm[i] = make_float4(i * 1e-3f, 1e-3f, 2e-3f, 3e-3f);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const float4* pa = sm + (lane >> 2) * 193;          // rows of 772 floats (3088 B)
  const float4* pb = pa + 8;
  const float4* qp = sm + 1024 + (lane & 3) * 97;     // 4 rows, 97 float4 apart (bank offset 4 like 6160 B)
  float2 acc[2][NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) { acc[0][c] = make_float2(0.f, 0.f); acc[1][c] = make_float2(0.f, 0.f); }
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int st = 0; st < 2; ++st) {
      const int k = (it * 2 + st) & 15;
      const float4 a = pa[k], b = pb[k];
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const float4 q0 = qp[c * 32 + 2 * k], q1 = qp[c * 32 + 2 * k + 1];
        acc[0][c] = ffma2(make_float2(a.x, a.x), make_float2(q0.x, q0.y), acc[0][c]);
        acc[1][c] = ffma2(make_float2(b.x, b.x), make_float2(q0.x, q0.y), acc[1][c]);
        acc[0][c] = ffma2(make_float2(a.y, a.y), make_float2(q0.z, q0.w), acc[0][c]);
        acc[1][c] = ffma2(make_float2(b.y, b.y), make_float2(q0.z, q0.w), acc[1][c]);
        acc[0][c] = ffma2(make_float2(a.z, a.z), make_float2(q1.x, q1.y), acc[0][c]);
        acc[1][c] = ffma2(make_float2(b.z, b.z), make_float2(q1.x, q1.y), acc[1][c]);
        acc[0][c] = ffma2(make_float2(a.w, a.w), make_float2(q1.z, q1.w), acc[0][c]);
        acc[1][c] = ffma2(make_float2(b.w, b.w), make_float2(q1.z, q1.w), acc[1][c]);
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int c = 0; c < NCH; ++c) s += acc[0][c].x + acc[0][c].y + acc[1][c].x + acc[1][c].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

// 8-prototype x 8-query register tile (lanes = 8 prototype groups x 4 query groups, warp tile 64 x 32): per 4 columns
// 8 + 8 LDS.128 feed 128 FFMA2 = 4 FMA per float delivered from shared memory.  P chunk [64 rows x 32 floats] with the
// 128-byte swizzle a TMA box would leave; query pairs interleaved, pair rows padded by 16 B.  Reports FMA per clk per SM.
__global__ void __launch_bounds__(256, 1) k_tile88(float* out, long long* clk) {
  __shared__ float4 sp[64 * 8];         // [64 rows][8 float4], read by every warp
  __shared__ float4 sq[16 * 17];        // 16 pair rows x (16 float4 + 1 pad)
  for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) sp[i] = make_float4(i * 1e-4f, 1e-3f, 2e-3f, 3e-3f);
  for (int i = threadIdx.x; i < 16 * 17; i += blockDim.x) sq[i] = make_float4(i * 1e-3f, 1e-3f, 2e-3f, 3e-3f);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, h = lane & 3;
  const float4* pw = sp + (warp & 0);
  float2 acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll 2
    for (int c = 0; c < 8; ++c) {
      float4 p[8], q0[4], q1[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = pw[(g + 8 * i) * 8 + (c ^ g)];
#pragma unroll
      for (int j = 0; j < 4; ++j) { q0[j] = sq[(h + 4 * j) * 17 + 2 * c]; q1[j] = sq[(h + 4 * j) * 17 + 2 * c + 1]; }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[i][j] = ffma2(make_float2(p[i].x, p[i].x), make_float2(q0[j].x, q0[j].y), acc[i][j]);
          acc[i][j] = ffma2(make_float2(p[i].y, p[i].y), make_float2(q0[j].z, q0[j].w), acc[i][j]);
          acc[i][j] = ffma2(make_float2(p[i].z, p[i].z), make_float2(q1[j].x, q1[j].y), acc[i][j]);
          acc[i][j] = ffma2(make_float2(p[i].w, p[i].w), make_float2(q1[j].z, q1[j].w), acc[i][j]);
        }
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j].x + acc[i][j].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
// the same with 4 prototypes x 8 queries per thread (warp tile 32 x 32): 4 + 8 LDS.128 per 64 FFMA2
__global__ void __launch_bounds__(256, 1) k_tile48(float* out, long long* clk) {
  __shared__ float4 sp[64 * 8];
  __shared__ float4 sq[16 * 17];
  for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) sp[i] = make_float4(i * 1e-4f, 1e-3f, 2e-3f, 3e-3f);
  for (int i = threadIdx.x; i < 16 * 17; i += blockDim.x) sq[i] = make_float4(i * 1e-3f, 1e-3f, 2e-3f, 3e-3f);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, h = lane & 3;
  const float4* pw = sp + (warp & 0);
  float2 acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll 2
    for (int c = 0; c < 8; ++c) {
      float4 p[4], q0[4], q1[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) p[i] = pw[(g + 8 * i) * 8 + (c ^ g)];
#pragma unroll
      for (int j = 0; j < 4; ++j) { q0[j] = sq[(h + 4 * j) * 17 + 2 * c]; q1[j] = sq[(h + 4 * j) * 17 + 2 * c + 1]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[i][j] = ffma2(make_float2(p[i].x, p[i].x), make_float2(q0[j].x, q0[j].y), acc[i][j]);
          acc[i][j] = ffma2(make_float2(p[i].y, p[i].y), make_float2(q0[j].z, q0[j].w), acc[i][j]);
          acc[i][j] = ffma2(make_float2(p[i].z, p[i].z), make_float2(q1[j].x, q1[j].y), acc[i][j]);
          acc[i][j] = ffma2(make_float2(p[i].w, p[i].w), make_float2(q1[j].z, q1[j].w), acc[i][j]);
        }
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j].x + acc[i][j].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
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
  for (int rep = 0; rep < 2; ++rep) k_tile88<<<sms, 256>>>(out, clk);
  report("8x8 register tile (FMA/clk/SM)", 8, 8 * 128 * 2.0 / 4);
  for (int rep = 0; rep < 2; ++rep) k_tile48<<<sms, 256>>>(out, clk);
  report("4x8 register tile (FMA/clk/SM)", 8, 8 * 64 * 2.0 / 4);
  for (int warps : {8, 16}) {
    for (int rep = 0; rep < 2; ++rep) k_mixed<2><<<sms, warps * 32>>>(out, clk);
    report("mixed step NCH=2 (FMA/clk/SM)", warps, 2 * 16 * 2.0);
    for (int rep = 0; rep < 2; ++rep) k_mixed<3><<<sms, warps * 32>>>(out, clk);
    report("mixed step NCH=3 (FMA/clk/SM)", warps, 2 * 24 * 2.0);
    for (int rep = 0; rep < 2; ++rep) k_mixed<4><<<sms, warps * 32>>>(out, clk);
    report("mixed step NCH=4 (FMA/clk/SM)", warps, 2 * 32 * 2.0);
  }
  const char* names[6] = {"one address", "8 rows by lane>>2", "4 rows by lane&3", "32 distinct", "4 rows by lane>>3", "8 rows by lane&7"};
  for (int warps : {8, 16, 32}) {
    for (int mode = 0; mode < 6; ++mode) {
      char nm[96];
      for (int rep = 0; rep < 2; ++rep) k_lds<4><<<sms, warps * 32>>>(out, clk, mode);
      snprintf(nm, sizeof nm, "LDS.128 %s (instr/clk/SM)", names[mode]); report(nm, warps, 8.0 / 32);
      for (int rep = 0; rep < 2; ++rep) k_lds<2><<<sms, warps * 32>>>(out, clk, mode);
      snprintf(nm, sizeof nm, "LDS.64  %s (instr/clk/SM)", names[mode]); report(nm, warps, 8.0 / 32);
      for (int rep = 0; rep < 2; ++rep) k_lds<1><<<sms, warps * 32>>>(out, clk, mode);
      snprintf(nm, sizeof nm, "LDS.32  %s (instr/clk/SM)", names[mode]); report(nm, warps, 8.0 / 32);
    }
  }
  return 0;
}
