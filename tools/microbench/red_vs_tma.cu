// Microbenchmark (GPU box): throughput of scattering 128-byte fp32 rows into global memory with
//   (a) red.global.add.v4.f32        -- 8 lanes x 16 B per row (what msda_bwd_sg_kernel does)
//   (b) red.global.add.f32           -- 32 lanes x 4 B per row
//   (c) cp.reduce.async.bulk (TMA)   -- one bulk reduce-add of a 128-byte shared-memory row per row
//   (d) like (c) but 512-byte rows   -- to see the per-op vs per-byte cost of the TMA path
// Rows are chosen pseudo-randomly in a 27 MB (L2-resident) or 700 MB (HBM) target, like grad_value.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o red_vs_tma red_vs_tma.cu && ./red_vs_tma
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

__global__ void k_red_v4(float* dst, uint32_t nrows, int rows_per_warp) {
  const int lane = threadIdx.x & 31, g = lane >> 3, cl = lane & 7;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  for (int i = 0; i < rows_per_warp; i += 4) {
    const uint32_t row = hash32(warp * 4099u + i + g) % nrows;
    float* p = dst + (size_t)row * 32 + cl * 4;
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(1.f), "f"(2.f), "f"(3.f), "f"(4.f) : "memory");
  }
}

__global__ void k_red_s(float* dst, uint32_t nrows, int rows_per_warp) {
  const int lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  for (int i = 0; i < rows_per_warp; ++i) {
    const uint32_t row = hash32(warp * 4099u + i) % nrows;
    atomicAdd(dst + (size_t)row * 32 + lane, 1.f);
  }
}

template <int ROW_FLOATS>
__global__ void k_tma(float* dst, uint32_t nrows, int rows_per_warp) {
  extern __shared__ __align__(128) float smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* my = smem + w * ROW_FLOATS * 4;  // 4 staging rows per warp
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  for (int i = 0; i < rows_per_warp; i += 4) {
    // wait until the previous 4 bulk ops have finished READING the staging rows
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
    for (int j = lane; j < ROW_FLOATS * 4; j += 32) my[j] = 1.f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane < 4) {
      const uint32_t row = hash32(warp * 4099u + i + lane) % (nrows / (ROW_FLOATS / 32));
      float* p = dst + (size_t)row * ROW_FLOATS;
      const uint32_t s = (uint32_t)__cvta_generic_to_shared(my + lane * ROW_FLOATS);
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(p), "r"(s), "n"(ROW_FLOATS * 4) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (lane < 4) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <typename F>
float time_it(F launch, int reps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < reps; ++r) launch();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / reps * 1e3f;
}

int main() {
  for (size_t mb : {27, 700}) {
    const size_t bytes = mb << 20;
    float* dst; cudaMalloc(&dst, bytes); cudaMemset(dst, 0, bytes);
    const uint32_t nrows = (uint32_t)(bytes / 128);
    for (int total_rows : {307200, 4915200}) {
      const int rows_per_warp = 64;
      const int warps = total_rows / rows_per_warp;
      const int wpb = 4;
      const int blocks = warps / wpb;
      float t1 = time_it([&] { k_red_v4<<<blocks, wpb * 32>>>(dst, nrows, rows_per_warp); }, 20);
      float t2 = time_it([&] { k_red_s<<<blocks, wpb * 32>>>(dst, nrows, rows_per_warp); }, 20);
      float t3 = time_it([&] { k_tma<32><<<blocks, wpb * 32, wpb * 32 * 4 * 4>>>(dst, nrows, rows_per_warp); }, 20);
      float t4 = time_it([&] { k_tma<128><<<blocks / 4, wpb * 32, wpb * 128 * 4 * 4>>>(dst, nrows, rows_per_warp); }, 20);
      cudaError_t e = cudaDeviceSynchronize();
      printf("{\"target_mb\": %zu, \"rows128\": %d, \"red_v4_us\": %.2f, \"red_scalar_us\": %.2f, \"tma_reduce_128B_us\": %.2f, \"tma_reduce_512B_us\": %.2f, \"err\": \"%s\"}\n",
             mb, total_rows, t1, t2, t3, t4, cudaGetErrorString(e));
    }
    cudaFree(dst);
  }
  return 0;
}
