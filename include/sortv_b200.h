/*
 * sortv_b200.h -- C ABI of the B200-native `sort_vertices` operator.
 *
 * Drop-in boundary for the reference's pybind extension `sort_vertices`
 * (aloscene/utils/rotated_iou/cuda_op/sort_vert.cpp:6-29 -> sort_vertices_wrapper, sort_vert_kernel.cu:136-140 ->
 * sort_vertices_kernel, sort_vert_kernel.cu:42-134), which orders the vertices of the intersection polygon of two rotated
 * boxes counter-clockwise for the shoelace formula of the rotated-IoU loss (box_intersection_2d.py:132-154).
 *
 * Plain C, raw device pointers, no torch types; enqueues on `stream` and never synchronises; 0 = success, else the text is in
 * sortv_last_error_string().
 *
 *   vertices  (b, n, m, 2) float32  polygon candidates AROUND THEIR MEAN (the caller normalises, box_intersection_2d.py:150-153);
 *                                   the first 8 are box corners, the rest edge intersections
 *   mask      (b, n, m)    bool (1 byte)   candidate is a vertex of the intersection polygon
 *   num_valid (b, n)       int32           = sum(mask) per polygon
 *   idx       (b, n, 9)    int32  OUT      indices of the valid vertices in counter-clockwise order starting from the smallest
 *                                          angle, then the first one again, then padding with the index of an invalid
 *                                          intersection candidate (value 0, gradient 0): (A, B, C, ..., A, X, X, X)
 *
 * Semantics follow the reference kernel operation by operation (same float / double mix in the comparator, same scan order, so
 * the indices are identical), with its undefined corners pinned down:
 *   - compare_vertices() falls off its end when one of the two y coordinates is exactly 0 and the signs are not opposite
 *     (sort_vert_kernel.cu:16-40); the reference's sm_100a build (nvcc 12.9) returns false there, and so does this library;
 *   - `pad` is uninitialised in the reference when every intersection candidate is valid (sort_vert_kernel.cu:56-62, cannot
 *     happen for two rectangles); here pad = m - 1;
 *   - num_valid > 8 makes the reference write past idx[i][8] (sort_vert_kernel.cu:74-101, cannot happen for two rectangles);
 *     here the first 8 vertices of the order are written and idx[i][8] = idx[i][0].
 * m <= 32 runs in registers (m = 24 is the reference's only shape); larger m takes a local-memory path; m < 9 is rejected
 * like a malformed call (there must be intersection candidates to pad with).
 */
#ifndef SORTV_B200_H_
#define SORTV_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SORTV_ABI_VERSION 1
#define SORTV_MAX_NUM_VERT_IDX 9 /* sort_vert_kernel.cu:6 */

int sortv_version(void);
const char* sortv_last_error_string(void);
/* Kernels this library has launched in this process (evidence that the CUDA path ran; never reset). */
uint64_t sortv_kernel_launch_count(void);
/* Test / measurement knob, process-wide: 0 = default schedule (TMA-staged tile kernel for m = 24 and 16-byte aligned tensors),
 * 1 = register-resident kernels only (m = 16 / 24 / 32), 2 = generic kernel only, 3 = tile kernel without the in-CTA load
 * balancing, 4 = tile kernel that always runs the reference's selection rounds (no sorted-order fast path).  All variants
 * return identical indices. */
int sortv_set_variant(int variant);

/* Replaces sort_vertices_wrapper (sort_vert_kernel.cu:136-140; torch entry sort_vert.cpp:6-29). */
int sortv_sort_vertices(const float* vertices, const uint8_t* mask, const int32_t* num_valid, int32_t* idx, int b, int n, int m,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SORTV_B200_H_ */
