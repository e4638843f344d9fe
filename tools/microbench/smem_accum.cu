// Microbenchmark (GPU box): how fast can a CTA accumulate 128-byte fp32 rows into a PRIVATE shared-memory tile, compared
// with scattering them to global memory with red.global.add.v4.f32 (the L2-atomic path of msda_bwd_sg_kernel)?
//
// sm_100a has no native fp32 add on shared memory: atomicAdd(float*) on a __shared__ address compiles to a
// LDS / FADD / ATOMS.CAST.SPIN / BRA loop (checked with cuobjdump).  This measures what that loop costs per row when
// 8..32 warps of a CTA hit a tile of R rows (R = 169 / 794 / 1323: the coarsest one / two levels of the COCO pyramids).
//   (a) smem_cas    : lane = channel, one row per warp instruction (32 scalar CAS-adds)
//   (b) smem_int    : same addressing with the NATIVE integer ATOMS.ADD (lower bound of the shared-memory atomic path)
//   (c) global_red  : red.global.add.v4.f32, 4 rows per warp instruction, random rows of a 27 MB target
//   (d) mixed       : half of the rows (a), half (c), interleaved -- do the two paths overlap?
// Output: JSON lines, clk per row per SM.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o smem_accum smem_accum.cu && ./smem_accum
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

template <int MODE>  // 0 = smem cas, 1 = smem int, 2 = global red.v4, 3 = mixed
__global__ void k(float* gdst, uint32_t grows, int R, int rows_per_warp, float* sink) {
  extern __shared__ __align__(16) float tile[];
  const int lane = threadIdx.x & 31, g = lane >> 3, cl = lane & 7;
  for (int i = threadIdx.x; i < R * 32; i += blockDim.x) tile[i] = 0.f;
  __syncthreads();
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  for (int i = 0; i < rows_per_warp; i += 8) {
    if (MODE == 0 || MODE == 3) {
      const int n = MODE == 3 ? 4 : 8;
#pragma unroll
      for (int j = 0; j < n; ++j) {
        const uint32_t row = hash32(warp * 4099u + i + j) % (uint32_t)R;
        atomicAdd(&tile[row * 32 + lane], 1.0f + j);
      }
    }
    if (MODE == 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t row = hash32(warp * 4099u + i + j) % (uint32_t)R;
        atomicAdd(reinterpret_cast<int*>(&tile[row * 32 + lane]), 1 + j);
      }
    }
    if (MODE == 2 || MODE == 3) {
      const int n = MODE == 3 ? 1 : 2;
#pragma unroll
      for (int j = 0; j < n; ++j) {
        const uint32_t row = hash32(warp * 8191u + i + 4 * j + g) % grows;
        float* p = gdst + (size_t)row * 32 + cl * 4;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(1.f), "f"(2.f), "f"(3.f), "f"(4.f) : "memory");
      }
    }
  }
  __syncthreads();
  float s = 0.f;
  for (int i = threadIdx.x; i < R * 32; i += blockDim.x) s += tile[i];
  if (s == -1.f) sink[0] = s;
}

template <int MODE>
void run(const char* name, float* gdst, uint32_t grows, int R, int warps, float* sink, int sms, double clk_khz) {
  const int rows_per_warp = 4096;
  const size_t smem = (size_t)R * 128;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<sms, warps * 32, smem>>>(gdst, grows, R, rows_per_warp, sink);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0);
    k<MODE><<<sms, warps * 32, smem>>>(gdst, grows, R, rows_per_warp, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double rows_per_sm = (double)rows_per_warp * warps;
  const double clk = best * 1e-3 * clk_khz * 1e3;
  printf("{\"path\": \"%s\", \"tile_rows\": %d, \"warps_per_sm\": %d, \"us\": %.2f, \"clk_per_row_per_sm\": %.3f, \"err\": \"%s\"}\n",
         name, R, warps, best * 1e3, clk / rows_per_sm, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const int sms = p.multiProcessorCount;
  const uint32_t grows = 27u * 1024 * 1024 / 128;
  float* gdst; cudaMalloc(&gdst, (size_t)grows * 128); cudaMemset(gdst, 0, (size_t)grows * 128);
  float* sink; cudaMalloc(&sink, 4);
  const int Rs[3] = {169, 794, 1323};
  const int Ws[3] = {8, 16, 32};
  for (int r : Rs)
    for (int w : Ws) {
      run<0>("smem_cas", gdst, grows, r, w, sink, sms, clk_khz);
      run<1>("smem_int", gdst, grows, r, w, sink, sms, clk_khz);
    }
  for (int w : Ws) {
    run<2>("global_red_v4", gdst, grows, 169, w, sink, sms, clk_khz);
    run<3>("mixed_half_half", gdst, grows, 794, w, sink, sms, clk_khz);
  }
  // is the red.v4 rate an SM-side or an L2-side limit?  Same per-SM load on 1/2 and 1/4 of the SMs.
  run<2>("global_red_v4 (74 CTAs)", gdst, grows, 169, 32, sink, sms / 2, clk_khz);
  run<2>("global_red_v4 (37 CTAs)", gdst, grows, 169, 32, sink, sms / 4, clk_khz);
  run<2>("global_red_v4 (18 CTAs)", gdst, grows, 169, 32, sink, sms / 8, clk_khz);
  printf("{\"sms\": %d, \"clock_khz\": %d}\n", sms, clk_khz);
  return 0;
}
