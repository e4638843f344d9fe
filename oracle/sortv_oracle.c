/*
 * sortv_oracle.c -- CPU restatement of the reference `sort_vertices` kernel (rotated-IoU intersection polygon ordering).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may build, load or call this file.  The shipped operator
 * (aloception_oss_b200.rotated_iou -> libsortv_b200.so) never falls back to it.
 *
 * Parity status: PINNED, two ways.
 *   1. Known answers of the reference's own tests, run through the reference's unmodified pure-PyTorch pipeline
 *      (box_intersection_2d.py: box_intersection_th / box_in_box_th / build_vertices / calculate_area) with this oracle in
 *      the place of the CUDA op: aloscene/utils/rotated_iou/_test_box_intersection_2d.py:9-21 (areas 0.5 and 0.25),
 *      _test_corner_cases.py:12-27 (identical / touching / contained boxes) and unittest/test_oriented_boxes_2d.py
 *      (IoU / GIoU values) -- tests/test_sortv_oracle.py, fixtures from oracle/make_golden_sortv.py.
 *   2. Index-for-index against the reference's OWN CUDA kernel compiled for sm_100a (oracle/build_ref_sortv.py ->
 *      oracle/_ref/sort_vertices_ref.so) on the GPU box: tests/test_sortv_gpu.py, and the committed golden indices
 *      under tests/golden_sortv/ that run produced.
 *
 * What it follows: aloscene/utils/rotated_iou/cuda_op/sort_vert_kernel.cu
 *   compare_vertices      :16-40   (float differences against the DOUBLE constant 1e-8, squared norm = float fma + double
 *                                   add rounded to float, IEEE float division)
 *   sort_vertices_kernel  :42-134  (pad index :56-62, < 3 vertices :63-68, selection scan :69-101, closing index :104,
 *                                   padding :107-109, identical-boxes corner case :111-131)
 * The loop structure is the reference's (idx re-read for the previously selected vertex, idx written on every k), so the
 * restatement stays independent of the register-resident CUDA kernel it checks.
 *
 * Undefined corners of the reference, pinned as in include/sortv_b200.h: compare_vertices falling off its end returns
 * false (what the reference's sm_100a binary does); pad = m - 1 when no intersection candidate is invalid; num_valid > 8
 * is clamped to 8 selected vertices.
 *
 * Build: gcc -O2 -ffp-contract=off (the one fused multiply-add the GPU performs is written out with fmaf).
 */
#include <math.h>
#include <stdint.h>

#define MAX_IDX 9
#define INTER_OFF 8
#define EPS 1e-8

/* sort_vert_kernel.cu:16-40 */
static int compare_vertices(float x1, float y1, float x2, float y2) {
  if ((double)fabsf(x1 - x2) < EPS && (double)fabsf(y2 - y1) < EPS) return 0;
  if (y1 > 0 && y2 < 0) return 1;
  if (y1 < 0 && y2 > 0) return 0;
  /* nvcc contracts x*x + y*y into fma(x, x, y*y); the + EPSILON is a double addition rounded back to float */
  float n1 = (float)((double)fmaf(x1, x1, y1 * y1) + EPS);
  float n2 = (float)((double)fmaf(x2, x2, y2 * y2) + EPS);
  float a = (fabsf(x1) * x1) / n1;
  float b = (fabsf(x2) * x2) / n2;
  float d = a - b;
  if (y1 > 0 && y2 > 0) return (double)d > EPS;
  if (y1 < 0 && y2 < 0) return (double)d < EPS;
  return 0; /* no return statement in the reference; its sm_100a build yields false */
}

/* sort_vert_kernel.cu:42-134, polygon by polygon */
void sortv_oracle(const float* vertices, const uint8_t* mask, const int32_t* num_valid, int32_t* idx, int b, int n, int m) {
  const long long total = (long long)b * n;
  for (long long i = 0; i < total; ++i) {
    const float* v = vertices + i * m * 2;
    const uint8_t* mk = mask + i * m;
    int32_t* out = idx + i * MAX_IDX;
    int pad = m - 1;
    for (int j = INTER_OFF; j < m; ++j) {
      if (!mk[j]) {
        pad = j;
        break;
      }
    }
    if (num_valid[i] < 3) {
      for (int j = 0; j < MAX_IDX; ++j) out[j] = pad;
      continue;
    }
    const int nv = num_valid[i] > MAX_IDX - 1 ? MAX_IDX - 1 : num_valid[i];
    for (int j = 0; j < nv; ++j) {
      float x_min = 1;
      float y_min = (float)(-EPS);
      int i_take = 0;
      for (int k = 0; k < m; ++k) {
        float x = v[k * 2 + 0];
        float y = v[k * 2 + 1];
        if (j == 0) {
          if (mk[k] && compare_vertices(x, y, x_min, y_min)) {
            x_min = x;
            y_min = y;
            i_take = k;
          }
        } else {
          int i2 = out[j - 1];
          float x2 = v[i2 * 2 + 0];
          float y2 = v[i2 * 2 + 1];
          if (mk[k] && compare_vertices(x, y, x_min, y_min) && compare_vertices(x2, y2, x, y)) {
            x_min = x;
            y_min = y;
            i_take = k;
          }
        }
        out[j] = i_take;
      }
    }
    out[nv] = out[0];
    for (int j = nv + 1; j < MAX_IDX; ++j) out[j] = pad;
    if (num_valid[i] == 8) {
      int counter = 0;
      for (int j = 0; j < 4; ++j) {
        int check = out[j];
        for (int k = 4; k < INTER_OFF; ++k)
          if (out[k] == check) counter++;
      }
      if (counter == 4) {
        out[4] = out[0];
        for (int j = 5; j < MAX_IDX; ++j) out[j] = pad;
      }
    }
  }
}
