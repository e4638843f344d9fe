// Microbenchmark (GPU box): does the FLAVOUR of the load change the rate at which an SM gathers 128-byte rows that miss its
// L1 and hit L2?  (gather_l2.cu measured 2.0 clk per row per SM with ld.global.nc.v4 = LDG.E.128.CONSTANT; an L1 hit costs
// 1.05.)  Variants, all reading the same pseudo-random rows of a 27 MB (L2-resident) table, 8 loads in flight per lane:
//   nc128      ld.global.nc.v4.f32                      8 lanes per row, 4 rows per warp instruction (the operator's load)
//   nc128_na   ld.global.nc.L1::no_allocate.v4.f32      same, rows are not installed in L1
//   cg128      ld.global.cg.v4.f32                      cache at L2 only
//   cv128      ld.volatile.global.v4.f32
//   nc32       ld.global.nc.f32                         32 lanes per row, 1 row per warp instruction
//   nc64       ld.global.nc.v2.f32                      16 lanes per row, 2 rows per warp instruction
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_l2_paths gather_l2_paths.cu && ./gather_l2_paths
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

constexpr int UNROLL = 8;
enum { NC128, NC128_NA, CG128, CV128, NC32, NC64 };

template <int K>
__device__ __forceinline__ float load_sum(const float* p) {
  float a, b, c, d;
  if (K == NC128) asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "l"(p));
  else if (K == NC128_NA) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "l"(p));
  else if (K == CG128) asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "l"(p));
  else if (K == CV128) asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "l"(p));
  else if (K == NC64) { asm volatile("ld.global.nc.v2.f32 {%0,%1}, [%2];" : "=f"(a), "=f"(b) : "l"(p)); c = d = 0.f; }
  else { asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(a) : "l"(p)); b = c = d = 0.f; }
  return (a + b) + (c + d);
}

template <int K>
__global__ void __launch_bounds__(1024, 1) k(const float* __restrict__ table, uint32_t rows, int iters, float* __restrict__ sink) {
  constexpr int LPR = K == NC32 ? 32 : (K == NC64 ? 16 : 8);  // lanes per row
  constexpr int G = 32 / LPR;                                  // rows per warp instruction
  const int lane = threadIdx.x & 31, g = lane / LPR, cl = lane % LPR;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    float v[UNROLL];
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) {
      const uint32_t row = hash32((warp * 131071u + it) * 32u + j * G + g) % rows;
      v[j] = load_sum<K>(table + (size_t)row * 32 + cl * (32 / LPR));
    }
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) acc += v[j];
  }
  if (acc == 123.456f) sink[threadIdx.x] = acc;
}

template <int K>
void run(const char* name, const float* table, size_t mb, float* sink, int sms, int khz) {
  constexpr int G = K == NC32 ? 1 : (K == NC64 ? 2 : 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const uint32_t rows = (uint32_t)((mb << 20) / 128);
  for (int threads : {512, 1024}) {
    const int iters = 256;
    k<K><<<sms, threads>>>(table, rows, iters, sink);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
      cudaEventRecord(e0);
      k<K><<<sms, threads>>>(table, rows, iters, sink);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    const double rows_per_sm = (double)(threads / 32) * iters * UNROLL * G;
    const double clk = best * 1e-3 * khz * 1e3;
    printf("{\"load\": \"%s\", \"table_mb\": %zu, \"threads_per_sm\": %d, \"us\": %.2f, \"clk_per_row_per_sm\": %.3f, \"chip_TBps\": %.2f}\n", name, mb,
           threads, best * 1e3, clk / rows_per_sm, rows_per_sm * sms * 128 / (best * 1e-3) / 1e12);
  }
}

int main() {
  int dev = 0, sms = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const size_t max_bytes = (size_t)96 << 20;
  float* table; cudaMalloc(&table, max_bytes); cudaMemset(table, 0, max_bytes);
  float* sink; cudaMalloc(&sink, 4096);
  for (size_t mb : {27, 96}) {
    run<NC128>("nc128", table, mb, sink, sms, khz);
    run<NC128_NA>("nc128 L1::no_allocate", table, mb, sink, sms, khz);
    run<CG128>("cg128", table, mb, sink, sms, khz);
    run<CV128>("volatile128", table, mb, sink, sms, khz);
    run<NC64>("nc64", table, mb, sink, sms, khz);
    run<NC32>("nc32", table, mb, sink, sms, khz);
  }
  printf("{\"sms\": %d, \"clock_khz\": %d, \"err\": \"%s\"}\n", sms, khz, cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
