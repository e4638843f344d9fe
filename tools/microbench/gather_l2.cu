// Microbenchmark (GPU box): rate at which an SM can gather 128-byte rows that MISS its L1 and hit L2 (the common case of
// the operator's taps: L1 hit rate 22 % on an encoder call).  Each lane group (8 lanes x 16 B) reads pseudo-random rows
// of a table shared by all SMs; table sizes: 1 MB (L1-missing but tiny), 27 MB (one C2 value tensor, L2-resident),
// 96 MB (still L2), 700 MB (HBM).  8 independent LDG.128 in flight per lane.  Output: clk per row per SM.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_l2 gather_l2.cu && ./gather_l2
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

constexpr int UNROLL = 8;

__global__ void __launch_bounds__(1024, 1) k(const float* __restrict__ table, uint32_t rows, int iters, float* __restrict__ sink) {
  const int lane = threadIdx.x & 31, g = lane >> 3, cl = lane & 7;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    float4 v[UNROLL];
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) {
      const uint32_t row = hash32((warp * 131071u + it) * 32u + j * 4u + g) % rows;
      v[j] = __ldg(reinterpret_cast<const float4*>(table + (size_t)row * 32 + cl * 4));
    }
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) acc += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  }
  if (acc == 123.456f) sink[threadIdx.x] = acc;
}

int main() {
  int dev = 0, sms = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const size_t max_bytes = (size_t)700 << 20;
  float* table; cudaMalloc(&table, max_bytes); cudaMemset(table, 0, max_bytes);
  float* sink; cudaMalloc(&sink, 4096);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (size_t mb : {1, 27, 96, 700}) {
    const uint32_t rows = (uint32_t)((mb << 20) / 128);
    for (int threads : {256, 512, 1024}) {
      const int iters = 256;
      k<<<sms, threads>>>(table, rows, iters, sink);
      cudaDeviceSynchronize();
      float best = 1e30f;
      for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        k<<<sms, threads>>>(table, rows, iters, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
      }
      const double rows_per_sm = (double)(threads / 32) * iters * UNROLL * 4;
      const double clk = best * 1e-3 * khz * 1e3;
      printf("{\"table_mb\": %zu, \"threads_per_sm\": %d, \"us\": %.2f, \"clk_per_row_per_sm\": %.3f, \"chip_TBps\": %.2f}\n", mb, threads,
             best * 1e3, clk / rows_per_sm, rows_per_sm * sms * 128 / (best * 1e-3) / 1e12);
    }
  }
  printf("{\"sms\": %d, \"clock_khz\": %d, \"err\": \"%s\"}\n", sms, khz, cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
