// Microbenchmark (GPU box): pure load-instruction rate of the L1 / shared-memory pipe for 128-byte-row gathers,
// with (almost) no address arithmetic: each lane group starts at a pseudo-random row and walks UNROLL rows at a
// compile-time stride (immediate offsets), so a row costs one load + one FADD.
//   W = bytes per lane (4, 8, 16) -> rows per warp instruction = W / 4 (1, 2, 4)
// Complements gather_paths.cu (which is partly issue-bound by its index arithmetic).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_rate gather_rate.cu && ./gather_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int UNROLL = 16;
constexpr int ROW_STRIDE = 7;  // rows between consecutive loads of a lane group (distinct 128-B lines)

template <int W, bool SMEM>
__global__ void __launch_bounds__(1024, 1) k_rate(const float* __restrict__ table, int rows, int iters, float* __restrict__ sink) {
  extern __shared__ __align__(128) float smem[];
  const int lane = threadIdx.x & 31;
  const float* base = table + (size_t)blockIdx.x * rows * 32;
  if (SMEM) {
    for (int i = threadIdx.x; i < rows * 32; i += blockDim.x) smem[i] = base[i];
    __syncthreads();
  }
  constexpr int LPR = 128 / W;  // lanes per row
  const int g = lane / LPR, cl = lane % LPR;
  const int span = UNROLL * ROW_STRIDE;
  uint32_t r = ((threadIdx.x >> 5) * 37u + g * 11u + blockIdx.x) % (uint32_t)(rows - span);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const float* src = SMEM ? smem : base;
  for (int it = 0; it < iters; ++it) {
    const float* p = src + r * 32 + cl * (W / 4);
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) {
      const float* q = p + j * ROW_STRIDE * 32;
      if (W == 16) {
        const float4 v = SMEM ? *reinterpret_cast<const float4*>(q) : __ldg(reinterpret_cast<const float4*>(q));
        acc[j & 3] += (v.x + v.y) + (v.z + v.w);
      } else if (W == 8) {
        const float2 v = SMEM ? *reinterpret_cast<const float2*>(q) : __ldg(reinterpret_cast<const float2*>(q));
        acc[j & 3] += v.x + v.y;
      } else {
        acc[j & 3] += SMEM ? *q : __ldg(q);
      }
    }
    r += 29u;
    if (r >= (uint32_t)(rows - span)) r -= (uint32_t)(rows - span);
  }
  const float s = acc[0] + acc[1] + acc[2] + acc[3];
  if (s == 123.456f) sink[threadIdx.x] = s;
}

template <typename F>
float time_us(F launch, int reps) {
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

template <int W, bool SMEM>
void run(const char* name, const float* table, int rows, float* sink, int sms, int khz) {
  const int iters = 128;
  const int smem_bytes = SMEM ? rows * 128 : 0;
  if (SMEM) cudaFuncSetAttribute(k_rate<W, SMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  for (int threads : {256, 512, 1024}) {
    const float t = time_us([&] { k_rate<W, SMEM><<<sms, threads, smem_bytes>>>(table, rows, iters, sink); }, 10);
    const double instr = (double)(threads / 32) * iters * UNROLL;  // load instructions per SM
    const double rows_moved = instr * (W / 4);
    const double clk = t * 1e-6 * khz * 1e3;
    printf("{\"path\": \"%s\", \"threads_per_sm\": %d, \"us\": %.2f, \"clk_per_load_instr\": %.2f, \"clk_per_row\": %.2f, \"bytes_per_clk_per_sm\": %.1f}\n",
           name, threads, t, clk / instr, clk / rows_moved, rows_moved * 128 / clk);
  }
}

int main() {
  int dev = 0, sms = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const int rows = 512;  // 64 KB per CTA
  float* table; cudaMalloc(&table, (size_t)sms * rows * 128); cudaMemset(table, 0, (size_t)sms * rows * 128);
  float* sink; cudaMalloc(&sink, 4096);
  run<4, false>("ldg32 (1 row/instr, L1)", table, rows, sink, sms, khz);
  run<8, false>("ldg64 (2 rows/instr, L1)", table, rows, sink, sms, khz);
  run<16, false>("ldg128 (4 rows/instr, L1)", table, rows, sink, sms, khz);
  run<4, true>("lds32 (1 row/instr, smem)", table, rows, sink, sms, khz);
  run<8, true>("lds64 (2 rows/instr, smem)", table, rows, sink, sms, khz);
  run<16, true>("lds128 (4 rows/instr, smem)", table, rows, sink, sms, khz);
  printf("{\"sms\": %d, \"clock_khz\": %d, \"err\": \"%s\"}\n", sms, khz, cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
