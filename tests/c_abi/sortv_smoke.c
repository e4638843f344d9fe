/*
 * Plain-C consumer of include/sortv_b200.h: cudaMalloc'ed tensors, one call of sortv_sort_vertices on a side stream, result
 * compared index for index with the C oracle (oracle/sortv_oracle.c -- test infrastructure, linked here as the checker).
 * Inputs in the spirit of the reference's demo (aloscene/utils/rotated_iou/cuda_op/cuda_ext.py:33-41): random vertices around
 * the origin, random mask; coordinates quantised so that ties occur; at most 8 valid candidates per polygon.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sortv_b200.h"

void sortv_oracle(const float* vertices, const uint8_t* mask, const int32_t* num_valid, int32_t* idx, int b, int n, int m);

static uint32_t rng_state = 12345u;
static float frand(void) {
  rng_state = rng_state * 1664525u + 1013904223u;
  return (float)(rng_state >> 8) * (1.0f / 16777216.0f);
}

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return 2;                                                                    \
    }                                                                              \
  } while (0)

int main(void) {
  const int b = 3, n = 1000, m = 24; /* 3000 polygons: 23 full tiles + a ragged one */
  const size_t np = (size_t)b * n;
  float* v = (float*)malloc(np * m * 2 * sizeof(float));
  uint8_t* mk = (uint8_t*)malloc(np * m);
  int32_t* nv = (int32_t*)malloc(np * sizeof(int32_t));
  int32_t* want = (int32_t*)malloc(np * 9 * sizeof(int32_t));
  int32_t* got = (int32_t*)malloc(np * 9 * sizeof(int32_t));
  for (size_t p = 0; p < np; ++p) {
    int cnt = 0;
    for (int k = 0; k < m; ++k) {
      /* quantised coordinates: ties, duplicates and y == 0 occur */
      v[(p * m + k) * 2 + 0] = (float)((int)(frand() * 8.f)) * 0.125f;
      v[(p * m + k) * 2 + 1] = (float)((int)(frand() * 8.f)) * 0.125f;
      mk[p * m + k] = (uint8_t)(frand() < 0.25f && cnt < 8);
      cnt += mk[p * m + k];
    }
    nv[p] = cnt;
    for (int k = 0; k < m; ++k) { /* centre of the unit square instead of the mean: keeps the quantisation exact */
      v[(p * m + k) * 2 + 0] -= 0.5f;
      v[(p * m + k) * 2 + 1] -= 0.5f;
    }
  }
  sortv_oracle(v, mk, nv, want, b, n, m);

  if (sortv_version() != SORTV_ABI_VERSION) {
    printf("ABI version mismatch\n");
    return 1;
  }
  float* dv;
  uint8_t* dm;
  int32_t *dn, *di;
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  CK(cudaMalloc((void**)&dv, np * m * 2 * sizeof(float)));
  CK(cudaMalloc((void**)&dm, np * m));
  CK(cudaMalloc((void**)&dn, np * sizeof(int32_t)));
  CK(cudaMalloc((void**)&di, np * 9 * sizeof(int32_t)));
  CK(cudaMemcpyAsync(dv, v, np * m * 2 * sizeof(float), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dm, mk, np * m, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dn, nv, np * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  for (int variant = 0; variant <= 4; ++variant) {
    if (sortv_set_variant(variant) != 0) {
      printf("sortv_set_variant(%d): %s\n", variant, sortv_last_error_string());
      return 1;
    }
    CK(cudaMemsetAsync(di, 0xff, np * 9 * sizeof(int32_t), st));
    const uint64_t l0 = sortv_kernel_launch_count();
    if (sortv_sort_vertices(dv, dm, dn, di, b, n, m, (void*)st) != 0) {
      printf("sortv_sort_vertices: %s\n", sortv_last_error_string());
      return 1;
    }
    if (sortv_kernel_launch_count() != l0 + 1) {
      printf("launch counter did not advance\n");
      return 1;
    }
    CK(cudaMemcpyAsync(got, di, np * 9 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    size_t bad = 0;
    for (size_t i = 0; i < np * 9; ++i) bad += got[i] != want[i];
    printf("variant %d     %s (%zu of %zu indices differ)\n", variant, bad ? "MISMATCH" : "ok", bad, np * 9);
    if (bad) return 1;
  }
  sortv_set_variant(0);
  /* error path: fewer than 9 candidates is a malformed call; the message is retrievable */
  if (sortv_sort_vertices(dv, dm, dn, di, b, n, 8, (void*)st) == 0 || strstr(sortv_last_error_string(), "candidates") == NULL) {
    printf("error path: expected a failure mentioning the candidates\n");
    return 1;
  }
  printf("error path    ok\n");
  cudaFree(dv); cudaFree(dm); cudaFree(dn); cudaFree(di);
  cudaStreamDestroy(st);
  free(v); free(mk); free(nv); free(want); free(got);
  printf("sortv C ABI smoke: OK\n");
  return 0;
}
