/*
 * msda_b200.h -- C ABI of the B200-native multi-scale deformable attention operator.
 *
 * This is the drop-in boundary for the reference's native op library
 * `MultiScaleDeformableAttention.so` (torch namespace `alonet_custom`).  Every entry point names
 * the reference interface it replaces; paths are relative to the reference repository root.
 *
 * Conventions
 *   - plain C, no torch / ATen types; all tensor arguments are raw pointers into memory owned by the
 *     caller, contiguous row-major; the library never allocates or frees caller tensors;
 *   - `stream` is a `cudaStream_t` passed as `void*` (NULL = legacy default stream); every device entry
 *     point only ENQUEUES work on it and never synchronises the host (the reference launches on
 *     at::cuda::getCurrentCUDAStream(): alonet/deformable_detr/ops/src/cuda/ms_deform_attn_cuda.cu:65,135);
 *   - return value 0 = success; non-zero = error, text available from msda_last_error_string()
 *     (the reference only printf()s launch failures: ms_deform_im2col_cuda.cuh:948-952,1321-1325);
 *   - thread-safe and re-entrant; no global state besides read-mostly tuning knobs.
 *
 * Tensor layouts (identical to the reference op, ms_deform_attn_cuda.cu:40-48):
 *   value             (N, S, M, D)        S = sum_l H_l * W_l
 *   spatial_shapes    (L, 2)  int32, DEVICE memory, (H_l, W_l)      [ms_deform_attn_cuda.cu:67]
 *   level_start_index (L,)    int32, DEVICE memory                  [ms_deform_attn_cuda.cu:68]
 *   sampling_loc      (N, Lq, M, L, P, 2) (x, y) normalised to [0,1]
 *   attn_weight       (N, Lq, M, L, P)
 *   output            (N, Lq, M*D)
 */
#ifndef MSDA_B200_H_
#define MSDA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSDA_ABI_VERSION 1

/* Element type of value / sampling_loc / attn_weight / output / gradients.
 * The reference op dispatches float and double only (AT_DISPATCH_FLOATING_TYPES,
 * ms_deform_attn_cuda.cu:64,134); bf16 / f16 storage with fp32 arithmetic is an extension. */
enum msda_dtype { MSDA_F32 = 0, MSDA_BF16 = 1, MSDA_F16 = 2, MSDA_F64 = 3 };
/* Mixed precision (msda_forward / msda_backward, OR-ed into `dtype` next to MSDA_BF16 / MSDA_F16): sampling_loc (and
 * grad_sampling_loc) / attn_weight (and grad_attn_weight) are fp32 while value, output and grad_output are 16-bit.  This is
 * what torch.autocast hands the operator -- the Linear layers emit bf16, the location arithmetic with the fp32 reference
 * points stays fp32 -- and it matters: a bf16 location is quantised to 1/256 of the image (0.65 px on a 167-px level).
 * Every kernel family serves it (the deterministic backward too); ignored for MSDA_F32, an error for MSDA_F64. */
#define MSDA_LOC_F32 0x100
#define MSDA_ATTN_F32 0x200

/* Problem sizes, in the reference's naming (ms_deform_attn_cuda.cu:40-48). */
typedef struct msda_dims {
  int batch;        /* N  = value.size(0)            */
  int spatial_size; /* S  = value.size(1)            */
  int num_heads;    /* M  = value.size(2)            */
  int channels;     /* D  = value.size(3)            */
  int num_levels;   /* L  = spatial_shapes.size(0)   */
  int num_query;    /* Lq = sampling_loc.size(1)     */
  int num_point;    /* P  = sampling_loc.size(4)     */
} msda_dims;

/* ABI version of the loaded library (== MSDA_ABI_VERSION it was built with). */
int msda_version(void);

/* Text of the last error raised on the calling thread ("" if none). */
const char* msda_last_error_string(void);

/*
 * Forward.  Replaces `ms_deform_attn_cuda_forward` (ms_deform_attn_cuda.cu:20-80) together with its
 * launcher `ms_deformable_im2col_cuda` and kernel `ms_deformable_im2col_gpu_kernel`
 * (ms_deform_im2col_cuda.cuh:923-954, 237-299), i.e. the body of torch op
 * `alonet_custom::ms_deform_attn_forward` (src/vision.cpp:22).
 * `output` need not be initialised (every element is written).  `im2col_step` of the reference is a
 * batching knob of its launcher only and has no equivalent here.
 */
int msda_forward(const void* value, const int32_t* spatial_shapes, const int32_t* level_start_index,
                 const void* sampling_loc, const void* attn_weight, void* output, const msda_dims* dims,
                 int dtype, void* stream);

/*
 * Forward with caller-provided scratch.  Same operation and results (bit for bit) as msda_forward; `workspace` -- device
 * memory of at least msda_forward_workspace_bytes() bytes, 8-byte aligned, owned by this call until it completes on `stream`,
 * contents irrelevant -- lets the library pick schedules that need a few scheduling words on the device: the SM-affine paired
 * forward for encoder-sized calls (queries = the pixels of the pyramid, Lq == S).  The library never
 * allocates device memory itself (allocation is not capturable in a CUDA graph and not stream-ordered); with workspace == NULL
 * (or too small) the call is msda_forward.  msda_forward_workspace_bytes is a pure host function: 0 when no such schedule would
 * be chosen for this problem (every decoder-sized call), a few KB otherwise.
 */
size_t msda_forward_workspace_bytes(const msda_dims* dims, int dtype);
int msda_forward_ws(const void* value, const int32_t* spatial_shapes, const int32_t* level_start_index,
                    const void* sampling_loc, const void* attn_weight, void* output, void* workspace,
                    size_t workspace_bytes, const msda_dims* dims, int dtype, void* stream);

/*
 * Bytes of scratch `msda_backward` needs for this problem (0 for MSDA_F32 / MSDA_F64; an fp32
 * accumulation image of grad_value for the 16-bit types).
 */
size_t msda_backward_workspace_bytes(const msda_dims* dims, int dtype);
/* Same, for a call with `flags` (MSDA_BWD_DETERMINISTIC needs a 64-bit fixed-point image of grad_value + a 256-byte header). */
size_t msda_backward_workspace_bytes_ex(const msda_dims* dims, int dtype, int flags);

/*
 * Backward.  Replaces `ms_deform_attn_cuda_backward` (ms_deform_attn_cuda.cu:83-153), the dispatcher
 * `ms_deformable_col2im_cuda` and its six kernels (ms_deform_im2col_cuda.cuh:956-1327, 301-920), i.e. the
 * body of torch op `alonet_custom::ms_deform_attn_backward` (src/vision.cpp:23).
 * None of the three gradient tensors needs initialising: grad_value is zero-filled by the library,
 * grad_sampling_loc and grad_attn_weight are written in full.
 * `workspace` : device scratch of at least msda_backward_workspace_bytes() bytes (may be NULL when 0).
 * `flags`     : 0, or MSDA_BWD_PREZEROED: the caller has already zero-filled the accumulation buffer (grad_value for
 *               MSDA_F32 / MSDA_F64, `workspace` for the 16-bit types) in stream order before this call -- e.g. with
 *               msda_zero_fill() on a side stream while the forward pass ran; the library then skips its own fill.
 *               MSDA_BWD_DETERMINISTIC: bit-reproducible grad_value.  The reference scatters with fp32 atomicAdd
 *               (ms_deform_im2col_cuda.cuh:125-152) and so does the default path here (vectorised reds): the result depends
 *               on the order in which the additions reach L2.  With this flag every contribution is converted to 64-bit
 *               fixed point (scale derived from max|grad_output| and max|attn_weight| of the call) and accumulated with
 *               integer reds -- associative, hence order-independent -- then converted back once per element.  Needs
 *               msda_backward_workspace_bytes_ex(dims, dtype, flags) bytes of workspace; f32 / bf16 / f16, D in
 *               {16, 32, 64, 128}, finite inputs; slower than the default (8-byte scalar reds).  grad_sampling_loc and
 *               grad_attn_weight never depend on atomics and are bit-reproducible in either mode.
 */
#define MSDA_BWD_PREZEROED 1
#define MSDA_BWD_DETERMINISTIC 2
int msda_backward(const void* grad_output, const void* value, const int32_t* spatial_shapes,
                  const int32_t* level_start_index, const void* sampling_loc, const void* attn_weight,
                  void* grad_value, void* grad_sampling_loc, void* grad_attn_weight, void* workspace,
                  size_t workspace_bytes, const msda_dims* dims, int dtype, int flags, void* stream);

/*
 * The library's zero-fill (the first of the two kernels of msda_backward) as a call of its own, so that a caller can
 * enqueue it early on another stream (see MSDA_BWD_PREZEROED).  The reference zero-fills inside the op
 * (`at::zeros_like`, ms_deform_attn_cuda.cu:121).
 */
int msda_zero_fill(void* ptr, size_t bytes, void* stream);

/*
 * Host-buffer forward for callers without a device allocator (the TensorRT-plugin-style consumer,
 * alonet/torch2trt/plugins/ms_deform_im2col/sources/ms_deform_im2col_kernel.cu:261-327, hands the kernel
 * raw pointers in the same way).  All seven pointers are HOST memory (pinned memory makes the copies
 * asynchronous); the call stages them through stream-ordered device allocations, runs the same kernel as
 * msda_forward and returns after `output` is complete on the host.
 */
int msda_forward_host(const void* value, const int32_t* spatial_shapes, const int32_t* level_start_index,
                      const void* sampling_loc, const void* attn_weight, void* output, const msda_dims* dims,
                      int dtype, void* stream);

/*
 * Fused operator: the elementwise part of `MSDeformAttn.forward` done inside the kernels
 * (alonet/deformable_detr/ops/modules/ms_deform_attn.py:121-137 -- softmax of the attention logits over L*P, and
 * sampling_locations = reference_points + offsets / (W_l, H_l)            for reference_points (N, Lq, L, 2), or
 *                     = ref_xy + offsets / P * ref_wh * 0.5               for reference_points (N, Lq, L, 4)),
 * followed by the same sampling as msda_forward.  Inputs are the RAW outputs of the two Linear layers:
 *   sampling_offsets (N, Lq, M, L, P, 2), attn_logits (N, Lq, M, L*P); reference_points (N, Lq, L, ref_dim).
 * This is SURVEY.md section 8(f) row 1 (the reference materialises sampling_locations / attention_weights with ~5
 * elementwise kernels per call).  Served by the vector kernels only: msda_fused_supported() tells whether a
 * problem qualifies (L*P <= 32, D in {16,32,64,128}, f32/bf16/f16); otherwise the entry points return 2 and the
 * caller composes the unfused operator.
 * msda_fused_backward: grad_offsets / grad_logits are gradients w.r.t. the raw inputs; grad_reference_points
 * (fp32, (N, Lq, L, ref_dim), ZEROED BY THE CALLER, may be NULL) is accumulated with reds.
 * MSDA_FUSED_REF_F32 (flag of msda_fused_forward_ex / msda_fused_backward): reference_points are fp32 although `dtype` is a
 * 16-bit type -- what torch.autocast produces (the Linear layers emit bf16, the reference points stay fp32).  Rounding a
 * reference point to bf16 moves the sample by up to 1/256 of the image (0.4-0.8 px on a 100-200 px level); with the flag
 * the location arithmetic starts from the exact fp32 point.  Ignored for MSDA_F32.
 */
#define MSDA_FUSED_REF_F32 4
int msda_fused_supported(const msda_dims* dims, int dtype, int ref_dim);
int msda_fused_forward(const void* value, const int32_t* spatial_shapes, const int32_t* level_start_index,
                       const void* reference_points, int ref_dim, const void* sampling_offsets, const void* attn_logits,
                       void* output, const msda_dims* dims, int dtype, void* stream);
int msda_fused_forward_ex(const void* value, const int32_t* spatial_shapes, const int32_t* level_start_index,
                          const void* reference_points, int ref_dim, const void* sampling_offsets, const void* attn_logits,
                          void* output, const msda_dims* dims, int dtype, int flags, void* stream);
int msda_fused_backward(const void* grad_output, const void* value, const int32_t* spatial_shapes,
                        const int32_t* level_start_index, const void* reference_points, int ref_dim,
                        const void* sampling_offsets, const void* attn_logits, void* grad_value, void* grad_offsets,
                        void* grad_logits, float* grad_reference_points, void* workspace, size_t workspace_bytes,
                        const msda_dims* dims, int dtype, int flags, void* stream);

/*
 * TensorRT-plugin twin.  Same signature, argument order and return convention as the
 * kernel wrapper the reference's `MsDeformIm2ColTRT` plugin calls from `IPluginV2IOExt::enqueue`
 * (alonet/torch2trt/plugins/ms_deform_im2col/sources/ms_deform_im2col_kernel.h:10-26, ..._kernel.cu:261-327,
 * called at ..._plugin.cpp:99-110): plugin inputs carry no batch dimension, `batch_size` comes from enqueue();
 * `data_type` takes the values of nvinfer1::DataType (0 = kFLOAT, 1 = kHALF); anything else returns -1 like the
 * reference; 0 = success.  All pointers are device memory, the launch goes on `stream`.
 *   data_value (batch, spatial_size, num_heads, channels); data_sampling_loc (batch, num_query, num_heads, num_levels,
 *   num_point, 2); data_attn_weight (batch, num_query, num_heads, num_levels, num_point);
 *   data_col (batch, num_query, num_heads * channels).
 * kHALF: storage is fp16, arithmetic fp32, and the pixel mapping is the operator's `loc * size - 0.5` -- the
 * reference's half kernel maps `loc * (size - 1)` instead (..._kernel.cu:245-246, its `- 0.5` is commented out), which
 * disagrees with its own float kernel and with `ms_deform_attn_core_pytorch`; the twin follows the float semantics.
 * A maintainer swaps the call in `MsDeformIm2Col::enqueue` for this symbol.
 */
int msda_im2col_inference(void* stream, const void* data_value, const void* data_spatial_shapes,
                          const void* data_level_start_index, const void* data_sampling_loc,
                          const void* data_attn_weight, int batch_size, int spatial_size, int num_heads, int channels,
                          int num_levels, int num_query, int num_point, void* data_col, int data_type);

/*
 * Tuning knobs (benchmark / test use).  Unknown names return non-zero.  Names:
 *   "force_generic"    0|1   route every call through the shape-generic kernels
 *   "fwd_unroll"       0=auto, 1, 2 or 4 samples in flight per lane group in the vector forward kernel
 *   "bwd_unroll"       0=auto, 1, 2 or 4 likewise for the vector backward kernel
 *   "warps_per_block"  0=auto, 1..4 (vector kernels), 1..8 (generic kernels)
 *   "no_pdl"           0|1   launch the backward kernel without programmatic dependent launch
 *   "head_major"       0=auto, 1=unit-major CTAs, 2=head-major CTAs (one head x consecutive queries per CTA)
 *   "smem_records"     0=auto, 1=per-sample records broadcast with warp shuffles, 2=through shared memory (forward)
 *   "patch_mode"       0=auto, 1=unit-ordered forward, 2=patch-ordered persistent forward (pixel-aligned queries)
 *   "patch_px/py"      0=default, else patch width in queries / patch height (= warps per CTA, <= 16)
 *   "patch_ctas"       0=auto, else persistent CTAs per SM of the patch-ordered forward
 *   "staged_mode"      0=auto, 1=off, 2=TMA-staged persistent forward (coarse levels of one (image, head) in shared memory)
 *   "staged_kb"        0=all the shared memory there is, else tile budget in KB;  "staged_warps" 0=32, else warps per CTA
 *   "staged_variant"   0=sample rounds unrolled when L*P == 16, 1=run-time loop
 *   "spec_mode"        0=auto (on), 1=flagged forward gather (zero-line address for invalid taps), 2=speculative regular-window
 *                      gather (zero weight on a clamped address; non-finite sums are redone on the flagged path)
 *   "zero_mode"        0=auto, 1=128-bit store kernel, 2=TMA bulk-store kernel (zero fill of grad_value)
 *   "zero_ctas"        0=default, else zero-fill CTAs per SM;  "zero_threads" threads per zero-fill CTA;
 *   "zero_chunk_kb"    bytes per TMA bulk store of the zero fill, in KB
 *   "bwd_tile_mode"    0=auto (off), 1=unit-ordered backward, 2=tile-binned backward (msda_bwd_tile.cuh: query-tiled
 *                      privatised accumulation of grad_value -- records counting-sorted by destination in shared memory,
 *                      summed in registers, ONE red per destination and tile; fp32, D = 32, P = 4, L <= 16, S <= 2^19,
 *                      16-byte aligned tensors; other problems keep the unit-ordered kernel)
 *   "bwd_tile_ctas"    0=2, else persistent CTAs per SM of the tile-binned backward (1 or 2)
 *   "bwd_chunk_mb"     0=auto (64), -1=off, else MB: backward calls whose accumulation image exceeds 96 MB are issued as groups of
 *                      images of at most this size (zero-fill + scatter per group), so that every grad_value line meets DRAM once
 *   "bwd_two_pass"     0=auto, 1=unit-ordered backward scatters each round right behind its gather, 2=all gather rounds, then
 *                      the fence behind the zero-fill, then all scatter rounds
 *   "fwd_pair_mode"    0=auto (= off), 1=off, 2=paired forward: one warp serves two heads of a query (L*P <= 16, D = 32; 24 % fewer
 *                      warp instructions), 3=paired + SM-affine patch order for pixel-aligned queries (needs msda_forward_ws'
 *                      workspace; else 2).  Bit-identical results; measured slower than the unit-ordered forward (profiles/)
 *   "fwd_win_mode"     0=auto (= off), 1=off, 2=windowed forward for pixel-aligned queries (fp32, D = 32, L*P <= 16): a CTA stages the
 *                      per-level boxes of `value` its 8x8 query tile samples in shared memory and gathers from there; levels whose
 *                      box exceeds the budget stay on the global path.  Bit-identical; measured slower (profiles/)
 *   "fwd_pair_px/py"   log2 of the SM-affine patch size in queries (default 3, 3 = 8 x 8); "fwd_pair_ctas": its CTAs per SM (5)
 */
int msda_set_tuning(const char* name, int value);
int msda_get_tuning(const char* name, int* value);

/* Number of kernels this library has launched since load (all threads); for launch accounting. */
uint64_t msda_kernel_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MSDA_B200_H_ */
