/* Plain-C consumer of the C ABI (include/msda_b200.h): no Python, no torch.
 *
 * What a non-torch host (the reference's TensorRT plugin is one: alonet/torch2trt/plugins/ms_deform_im2col/sources/
 * ms_deform_im2col_plugin.cpp:99-110) does with the library: cudaMalloc the tensors, call msda_forward / msda_backward /
 * msda_forward_host / msda_im2col_inference on a stream, read the results back.  Results are checked against the C oracle
 * (oracle/msda_oracle.c, test infrastructure) on the same inputs: fp32, rtol 1e-4.
 *
 *   gcc -std=c11 -O1 -Iinclude -I/usr/local/cuda/include tests/c_abi/c_abi_smoke.c -o c_abi_smoke \
 *       -Laloception_oss_b200 -lmsda_b200 -Loracle -lmsda_oracle -L/usr/local/cuda/lib64 -lcudart -lm
 */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "msda_b200.h"

void msda_oracle_forward_f32(const float* value, const int32_t* shapes, const int32_t* start, const float* loc, const float* attn,
                             float* out, int N, int S, int M, int D, int L, int Lq, int P);
void msda_oracle_backward_f32(const float* grad_out, const float* value, const int32_t* shapes, const int32_t* start, const float* loc,
                              const float* attn, float* grad_value, float* grad_loc, float* grad_attn, int N, int S, int M, int D,
                              int L, int Lq, int P);

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)

static uint32_t rng = 12345u;
static float frand(void) { rng = rng * 1664525u + 1013904223u; return (float)(rng >> 8) / 16777216.0f; }

static int close_enough(const char* what, const float* got, const float* want, size_t n) {
  double scale = 0;
  for (size_t i = 0; i < n; ++i) scale += (double)want[i] * want[i];
  scale = sqrt(scale / (n ? n : 1));
  for (size_t i = 0; i < n; ++i)
    if (fabs((double)got[i] - want[i]) > 1e-4 * fabs(want[i]) + 1e-4 * scale) {
      fprintf(stderr, "%s[%zu]: got %g want %g\n", what, i, got[i], want[i]);
      return 0;
    }
  printf("%-12s ok (%zu values, rms %.3e)\n", what, n, scale);
  return 1;
}

int main(void) {
  enum { N = 2, M = 8, D = 32, L = 3, Lq = 37, P = 4 };
  const int32_t shapes[L * 2] = {9, 11, 5, 6, 3, 3};
  int32_t start[L];
  int S = 0;
  for (int l = 0; l < L; ++l) { start[l] = S; S += shapes[2 * l] * shapes[2 * l + 1]; }
  const size_t n_value = (size_t)N * S * M * D, n_attn = (size_t)N * Lq * M * L * P, n_loc = 2 * n_attn, n_out = (size_t)N * Lq * M * D;
  float *value = malloc(4 * n_value), *loc = malloc(4 * n_loc), *attn = malloc(4 * n_attn), *go = malloc(4 * n_out);
  for (size_t i = 0; i < n_value; ++i) value[i] = frand() * 0.01f;
  for (size_t i = 0; i < n_loc; ++i) loc[i] = frand() * 1.4f - 0.2f; /* some samples fall outside the levels */
  for (size_t i = 0; i < n_attn; ++i) attn[i] = frand() / (L * P);
  for (size_t i = 0; i < n_out; ++i) go[i] = frand() - 0.5f;

  float *w_out = malloc(4 * n_out), *w_gv = malloc(4 * n_value), *w_gl = malloc(4 * n_loc), *w_ga = malloc(4 * n_attn);
  msda_oracle_forward_f32(value, shapes, start, loc, attn, w_out, N, S, M, D, L, Lq, P);
  msda_oracle_backward_f32(go, value, shapes, start, loc, attn, w_gv, w_gl, w_ga, N, S, M, D, L, Lq, P);

  if (msda_version() != MSDA_ABI_VERSION) { fprintf(stderr, "ABI version mismatch\n"); return 1; }
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  float *d_value, *d_loc, *d_attn, *d_go, *d_out, *d_gv, *d_gl, *d_ga;
  int32_t *d_shapes, *d_start;
  CK(cudaMalloc((void**)&d_value, 4 * n_value)); CK(cudaMalloc((void**)&d_loc, 4 * n_loc)); CK(cudaMalloc((void**)&d_attn, 4 * n_attn));
  CK(cudaMalloc((void**)&d_go, 4 * n_out)); CK(cudaMalloc((void**)&d_out, 4 * n_out)); CK(cudaMalloc((void**)&d_gv, 4 * n_value));
  CK(cudaMalloc((void**)&d_gl, 4 * n_loc)); CK(cudaMalloc((void**)&d_ga, 4 * n_attn));
  CK(cudaMalloc((void**)&d_shapes, sizeof shapes)); CK(cudaMalloc((void**)&d_start, sizeof start));
  CK(cudaMemcpyAsync(d_value, value, 4 * n_value, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_loc, loc, 4 * n_loc, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_attn, attn, 4 * n_attn, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_go, go, 4 * n_out, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_shapes, shapes, sizeof shapes, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_start, start, sizeof start, cudaMemcpyHostToDevice, st));

  const msda_dims dims = {N, S, M, D, L, Lq, P};
  float *g_out = malloc(4 * n_out), *g_gv = malloc(4 * n_value), *g_gl = malloc(4 * n_loc), *g_ga = malloc(4 * n_attn);
  int ok = 1;

  if (msda_forward(d_value, d_shapes, d_start, d_loc, d_attn, d_out, &dims, MSDA_F32, st)) { fprintf(stderr, "%s\n", msda_last_error_string()); return 1; }
  CK(cudaMemcpyAsync(g_out, d_out, 4 * n_out, cudaMemcpyDeviceToHost, st));
  if (msda_backward(d_go, d_value, d_shapes, d_start, d_loc, d_attn, d_gv, d_gl, d_ga, NULL, msda_backward_workspace_bytes(&dims, MSDA_F32),
                    &dims, MSDA_F32, 0, st)) { fprintf(stderr, "%s\n", msda_last_error_string()); return 1; }
  CK(cudaMemcpyAsync(g_gv, d_gv, 4 * n_value, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(g_gl, d_gl, 4 * n_loc, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(g_ga, d_ga, 4 * n_attn, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  ok &= close_enough("forward", g_out, w_out, n_out);
  ok &= close_enough("grad_value", g_gv, w_gv, n_value);
  ok &= close_enough("grad_attn", g_ga, w_ga, n_attn);
  /* grad_loc jumps where a pixel coordinate crosses an integer; with 7 104 random samples none sits within 1e-5 of one */
  ok &= close_enough("grad_loc", g_gl, w_gl, n_loc);

  /* host-buffer entry point */
  memset(g_out, 0, 4 * n_out);
  if (msda_forward_host(value, shapes, start, loc, attn, g_out, &dims, MSDA_F32, st)) { fprintf(stderr, "%s\n", msda_last_error_string()); return 1; }
  ok &= close_enough("forward_host", g_out, w_out, n_out);

  /* TensorRT-plugin twin: same arguments as the reference plugin's kernel wrapper, nvinfer1::DataType::kFLOAT = 0 */
  CK(cudaMemsetAsync(d_out, 0, 4 * n_out, st));
  if (msda_im2col_inference(st, d_value, d_shapes, d_start, d_loc, d_attn, N, S, M, D, L, Lq, P, d_out, 0)) { fprintf(stderr, "%s\n", msda_last_error_string()); return 1; }
  CK(cudaMemcpyAsync(g_out, d_out, 4 * n_out, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  ok &= close_enough("plugin_twin", g_out, w_out, n_out);

  /* deterministic backward: 64-bit fixed-point accumulation in caller-provided scratch, twice -> the same bits */
  {
    const size_t ws_bytes = msda_backward_workspace_bytes_ex(&dims, MSDA_F32, MSDA_BWD_DETERMINISTIC);
    void* d_ws = NULL;
    float* g_gv2 = malloc(4 * n_value);
    CK(cudaMalloc(&d_ws, ws_bytes));
    for (int rep = 0; rep < 2; ++rep) {
      if (msda_backward(d_go, d_value, d_shapes, d_start, d_loc, d_attn, d_gv, d_gl, d_ga, d_ws, ws_bytes, &dims, MSDA_F32,
                        MSDA_BWD_DETERMINISTIC, st)) { fprintf(stderr, "%s\n", msda_last_error_string()); return 1; }
      CK(cudaMemcpyAsync(rep ? g_gv2 : g_gv, d_gv, 4 * n_value, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    ok &= close_enough("grad_value (deterministic)", g_gv, w_gv, n_value);
    if (memcmp(g_gv, g_gv2, 4 * n_value)) { fprintf(stderr, "deterministic backward is not bit-reproducible\n"); ok = 0; }
    CK(cudaFree(d_ws));
    free(g_gv2);
  }

  /* forward with caller-provided scratch: no schedule needs it by default (0 bytes); the opt-in SM-affine paired forward does */
  {
    if (msda_forward_workspace_bytes(&dims, MSDA_F32) != 0) { fprintf(stderr, "default forward asks for scratch\n"); ok = 0; }
    msda_set_tuning("fwd_pair_mode", 3);
    const size_t ws_bytes = msda_forward_workspace_bytes(&dims, MSDA_F32);
    void* d_ws = NULL;
    if (ws_bytes == 0) { fprintf(stderr, "fwd_pair_mode = 3 asks for no scratch\n"); ok = 0; }
    CK(cudaMalloc(&d_ws, ws_bytes ? ws_bytes : 8));
    CK(cudaMemsetAsync(d_out, 0xff, 4 * n_out, st));
    if (msda_forward_ws(d_value, d_shapes, d_start, d_loc, d_attn, d_out, d_ws, ws_bytes, &dims, MSDA_F32, st)) { fprintf(stderr, "%s\n", msda_last_error_string()); return 1; }
    CK(cudaMemcpyAsync(g_out, d_out, 4 * n_out, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    ok &= close_enough("forward_ws (paired, SM-affine)", g_out, w_out, n_out);
    msda_set_tuning("fwd_pair_mode", 0);
    CK(cudaFree(d_ws));
  }

  /* mixed precision: bf16 value / output next to fp32 locations and weights (MSDA_LOC_F32 | MSDA_ATTN_F32) */
  {
    uint16_t* h16 = malloc(2 * n_value);
    float* vr = malloc(4 * n_value);
    for (size_t i = 0; i < n_value; ++i) {  /* round to nearest even, keep the rounded value for the oracle */
      uint32_t u; memcpy(&u, &value[i], 4);
      u += 0x7fffu + ((u >> 16) & 1u);
      h16[i] = (uint16_t)(u >> 16);
      u &= 0xffff0000u; memcpy(&vr[i], &u, 4);
    }
    float* w16 = malloc(4 * n_out);
    msda_oracle_forward_f32(vr, shapes, start, loc, attn, w16, N, S, M, D, L, Lq, P);
    uint16_t *d_v16, *d_o16, *o16 = malloc(2 * n_out);
    CK(cudaMalloc((void**)&d_v16, 2 * n_value)); CK(cudaMalloc((void**)&d_o16, 2 * n_out));
    CK(cudaMemcpyAsync(d_v16, h16, 2 * n_value, cudaMemcpyHostToDevice, st));
    if (msda_forward(d_v16, d_shapes, d_start, d_loc, d_attn, d_o16, &dims, MSDA_BF16 | MSDA_LOC_F32 | MSDA_ATTN_F32, st)) { fprintf(stderr, "%s\n", msda_last_error_string()); return 1; }
    CK(cudaMemcpyAsync(o16, d_o16, 2 * n_out, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    double worst = 0.0, scale = 0.0;
    for (size_t i = 0; i < n_out; ++i) {
      uint32_t u = (uint32_t)o16[i] << 16; float f; memcpy(&f, &u, 4);
      const double e = fabs((double)f - w16[i]);
      if (e > worst) worst = e;
      if (fabs(w16[i]) > scale) scale = fabs(w16[i]);
    }
    if (!(worst <= 1e-2 * scale)) { fprintf(stderr, "mixed-precision forward: max error %g vs scale %g\n", worst, scale); ok = 0; }
    else printf("forward (bf16 value, fp32 loc / attn): max |err| %.3g (output scale %.3g)\n", worst, scale);
    if (msda_forward(d_v16, d_shapes, d_start, d_loc, d_attn, d_o16, &dims, MSDA_F64 | MSDA_LOC_F32, st) == 0) { fprintf(stderr, "MSDA_F64 | MSDA_LOC_F32 was not rejected\n"); ok = 0; }
    CK(cudaFree(d_v16)); CK(cudaFree(d_o16));
    free(h16); free(vr); free(w16); free(o16);
  }

  /* error path: bad dtype is reported, not crashed on */
  if (msda_forward(d_value, d_shapes, d_start, d_loc, d_attn, d_out, &dims, 42, st) == 0 || !strstr(msda_last_error_string(), "dtype")) {
    fprintf(stderr, "bad dtype was not rejected\n");
    ok = 0;
  }
  printf("kernel launches from the library: %llu\n", (unsigned long long)msda_kernel_launch_count());
  printf(ok ? "C ABI smoke: OK\n" : "C ABI smoke: FAILED\n");
  return ok ? 0 : 1;
}
