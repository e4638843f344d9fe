// msda_torch_shim.cpp -- C++ registration of torch.ops.alonet_custom.ms_deform_attn_{forward,backward} on the C ABI.
//
// Replaces TORCH_LIBRARY(alonet_custom, m) of the reference (alonet/deformable_detr/ops/src/vision.cpp:21-24) and the
// launchers behind it (ops/src/ms_deform_attn.h:20-62, ops/src/cuda/ms_deform_attn_cuda.cu:20-153): same two schemas, same
// argument checks and messages, CUDA tensors only ("Not implemented on the CPU"), plus a Meta kernel.  The compute is
// include/msda_b200.h (libmsda_b200.so); this file only takes the Python interpreter and ctypes off the call path -- a
// decoder-sized call (5 us of kernel) is host-bound when issued eagerly (profiles/README.md "Host side").  The Python
// registration in torch_ops.py stays as the fallback when this library has not been built; both end in the same C calls.
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>

#include <cstdlib>
#include <vector>

#include "../../include/msda_b200.h"

namespace {

int dtype_of(const at::Tensor& t) {
  switch (t.scalar_type()) {
    case at::kFloat: return MSDA_F32;
    case at::kBFloat16: return MSDA_BF16;
    case at::kHalf: return MSDA_F16;
    case at::kDouble: return MSDA_F64;
    default: TORCH_CHECK(false, "ms_deform_attn: unsupported dtype ", t.scalar_type(), " (float32, float64, bfloat16, float16)");
  }
}

at::Tensor meta_i32(const at::Tensor& t, const char* name) {
  if (t.scalar_type() == at::kInt) return t;
  // upstream Deformable-DETR passes int64; this fork int32 (ms_deform_attn_cuda.cu:67-68)
  TORCH_CHECK(t.scalar_type() == at::kLong, name, " must be an int32 (or int64) tensor, got ", t.scalar_type());
  return t.to(at::kInt);
}

msda_dims check_inputs(const at::Tensor& value, const at::Tensor& shapes, const at::Tensor& start, const at::Tensor& loc,
                       const at::Tensor& attn, int64_t im2col_step, const at::Tensor* grad_output) {
  // ms_deform_attn_cuda.cu:28-38 -- same order, same messages
  TORCH_CHECK(value.is_contiguous(), "value tensor has to be contiguous");
  TORCH_CHECK(shapes.is_contiguous(), "spatial_shapes tensor has to be contiguous");
  TORCH_CHECK(start.is_contiguous(), "level_start_index tensor has to be contiguous");
  TORCH_CHECK(loc.is_contiguous(), "sampling_loc tensor has to be contiguous");
  TORCH_CHECK(attn.is_contiguous(), "attn_weight tensor has to be contiguous");
  if (grad_output) TORCH_CHECK(grad_output->is_contiguous(), "grad_output tensor has to be contiguous");
  TORCH_CHECK(value.is_cuda(), "value must be a CUDA tensor");
  TORCH_CHECK(shapes.is_cuda(), "spatial_shapes must be a CUDA tensor");
  TORCH_CHECK(start.is_cuda(), "level_start_index must be a CUDA tensor");
  TORCH_CHECK(loc.is_cuda(), "sampling_loc must be a CUDA tensor");
  TORCH_CHECK(attn.is_cuda(), "attn_weight must be a CUDA tensor");
  if (grad_output) TORCH_CHECK(grad_output->is_cuda(), "grad_output must be a CUDA tensor");
  // mixed precision: sampling_loc / attn_weight may stay float32 next to 16-bit value (MSDA_LOC_F32 / MSDA_ATTN_F32)
  const bool half = value.scalar_type() == at::kBFloat16 || value.scalar_type() == at::kHalf;
  TORCH_CHECK(loc.scalar_type() == value.scalar_type() || (half && loc.scalar_type() == at::kFloat), "sampling_loc has dtype ",
              loc.scalar_type(), ", expected ", value.scalar_type(), " (same as value)");
  TORCH_CHECK(attn.scalar_type() == value.scalar_type() || (half && attn.scalar_type() == at::kFloat), "attn_weight has dtype ",
              attn.scalar_type(), ", expected ", value.scalar_type(), " (same as value)");
  TORCH_CHECK(loc.device() == value.device() && attn.device() == value.device(), "sampling_loc / attn_weight are not on value's device");
  TORCH_CHECK(value.dim() == 4, "value must be (N, S, M, D), got ", value.sizes());
  TORCH_CHECK(loc.dim() == 6 && loc.size(5) == 2, "sampling_loc must be (N, Lq, M, L, P, 2), got ", loc.sizes());
  TORCH_CHECK(shapes.dim() == 2 && shapes.size(1) == 2, "spatial_shapes must be (L, 2), got ", shapes.sizes());
  const int64_t N = value.size(0), S = value.size(1), M = value.size(2), D = value.size(3);
  const int64_t L = shapes.size(0), Lq = loc.size(1), P = loc.size(4);
  TORCH_CHECK(loc.size(0) == N && loc.size(2) == M && loc.size(3) == L, "sampling_loc ", loc.sizes(), " inconsistent with value ",
              value.sizes(), " and ", L, " levels");
  TORCH_CHECK(attn.dim() == 5 && attn.size(0) == N && attn.size(1) == Lq && attn.size(2) == M && attn.size(3) == L && attn.size(4) == P,
              "attn_weight must be (", N, ", ", Lq, ", ", M, ", ", L, ", ", P, "), got ", attn.sizes());
  TORCH_CHECK(start.numel() == L, "level_start_index must have one entry per level");
  const int64_t step = std::min<int64_t>(N, im2col_step);
  // ms_deform_attn_cuda.cu:50-52 -- kept for API fidelity; the kernels do not batch by im2col_step
  TORCH_CHECK(N == 0 || (step > 0 && N % step == 0), "batch(", N, ") must divide im2col_step(", step, ")");
  if (grad_output) {
    TORCH_CHECK(grad_output->scalar_type() == value.scalar_type() && grad_output->device() == value.device(),
                "grad_output must have value's dtype and device");
    TORCH_CHECK(grad_output->numel() == N * Lq * M * D, "grad_output ", grad_output->sizes(), " does not match the forward output");
  }
  msda_dims d;
  d.batch = (int)N; d.spatial_size = (int)S; d.num_heads = (int)M; d.channels = (int)D;
  d.num_levels = (int)L; d.num_query = (int)Lq; d.num_point = (int)P;
  return d;
}

const void* ptr(const at::Tensor& t) { return t.numel() ? t.data_ptr() : nullptr; }

// the C ABI's dtype word: value's type plus the mixed-precision bits
int io_dtype(const at::Tensor& value, const at::Tensor& loc, const at::Tensor& attn) {
  int dt = dtype_of(value);
  if (loc.scalar_type() != value.scalar_type()) dt |= MSDA_LOC_F32;
  if (attn.scalar_type() != value.scalar_type()) dt |= MSDA_ATTN_F32;
  return dt;
}

at::Tensor forward_cuda(const at::Tensor& value, const at::Tensor& spatial_shapes, const at::Tensor& level_start_index,
                        const at::Tensor& sampling_loc, const at::Tensor& attn_weight, int64_t im2col_step) {
  const msda_dims d = check_inputs(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step, nullptr);
  const at::Tensor shapes = meta_i32(spatial_shapes, "spatial_shapes"), start = meta_i32(level_start_index, "level_start_index");
  const c10::cuda::CUDAGuard guard(value.device());
  at::Tensor out = at::empty({d.batch, d.num_query, (int64_t)d.num_heads * d.channels}, value.options());
  const int dt = io_dtype(value, sampling_loc, attn_weight);
  // scheduling words for the schedules that need them (msda_forward_ws): 0 bytes -- and no allocation -- for every default path
  const size_t ws_bytes = msda_forward_workspace_bytes(&d, dt);
  at::Tensor ws;
  if (ws_bytes) ws = at::empty({(int64_t)ws_bytes}, value.options().dtype(at::kByte));
  const int rc = msda_forward_ws(ptr(value), (const int32_t*)ptr(shapes), (const int32_t*)ptr(start), ptr(sampling_loc), ptr(attn_weight),
                                 const_cast<void*>(ptr(out)), ws_bytes ? ws.data_ptr() : nullptr, ws_bytes, &d, dt,
                                 (void*)at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(rc == 0, "msda_forward failed: ", msda_last_error_string());
  return out;
}

bool deterministic_requested() {
  static const bool env = [] { const char* e = std::getenv("MSDA_DETERMINISTIC"); return e && e[0] == '1'; }();
  return env || at::globalContext().deterministicAlgorithms();
}

std::vector<at::Tensor> backward_cuda(const at::Tensor& value, const at::Tensor& spatial_shapes, const at::Tensor& level_start_index,
                                      const at::Tensor& sampling_loc, const at::Tensor& attn_weight, const at::Tensor& grad_output,
                                      int64_t im2col_step) {
  const msda_dims d = check_inputs(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step, &grad_output);
  const at::Tensor shapes = meta_i32(spatial_shapes, "spatial_shapes"), start = meta_i32(level_start_index, "level_start_index");
  const c10::cuda::CUDAGuard guard(value.device());
  const int dt = io_dtype(value, sampling_loc, attn_weight);
  int flags = 0;
  // an implicit request (torch.use_deterministic_algorithms / MSDA_DETERMINISTIC) applies where the mode is served
  if (deterministic_requested() && dt != MSDA_F64 && (d.channels == 16 || d.channels == 32 || d.channels == 64 || d.channels == 128))
    flags = MSDA_BWD_DETERMINISTIC;
  at::Tensor grad_value = at::empty_like(value), grad_loc = at::empty_like(sampling_loc), grad_attn = at::empty_like(attn_weight);
  const size_t ws_bytes = msda_backward_workspace_bytes_ex(&d, dt, flags);
  at::Tensor ws;
  if (ws_bytes) ws = at::empty({(int64_t)ws_bytes}, value.options().dtype(at::kByte));
  const int rc = msda_backward(ptr(grad_output), ptr(value), (const int32_t*)ptr(shapes), (const int32_t*)ptr(start), ptr(sampling_loc),
                               ptr(attn_weight), const_cast<void*>(ptr(grad_value)), const_cast<void*>(ptr(grad_loc)),
                               const_cast<void*>(ptr(grad_attn)), ws_bytes ? ws.data_ptr() : nullptr, ws_bytes, &d, dt, flags,
                               (void*)at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(rc == 0, "msda_backward failed: ", msda_last_error_string());
  return {grad_value, grad_loc, grad_attn};
}

// ops/src/ms_deform_attn.h:38,60
at::Tensor forward_cpu(const at::Tensor&, const at::Tensor&, const at::Tensor&, const at::Tensor&, const at::Tensor&, int64_t) {
  TORCH_CHECK(false, "Not implemented on the CPU");
}
std::vector<at::Tensor> backward_cpu(const at::Tensor&, const at::Tensor&, const at::Tensor&, const at::Tensor&, const at::Tensor&,
                                     const at::Tensor&, int64_t) {
  TORCH_CHECK(false, "Not implemented on the CPU");
}

at::Tensor forward_meta(const at::Tensor& value, const at::Tensor&, const at::Tensor&, const at::Tensor& sampling_loc, const at::Tensor&,
                        int64_t) {
  return at::empty_symint({value.sym_size(0), sampling_loc.sym_size(1), value.sym_size(2) * value.sym_size(3)}, value.options());
}
std::vector<at::Tensor> backward_meta(const at::Tensor& value, const at::Tensor&, const at::Tensor&, const at::Tensor& sampling_loc,
                                      const at::Tensor& attn_weight, const at::Tensor&, int64_t) {
  return {at::empty_like(value), at::empty_like(sampling_loc), at::empty_like(attn_weight)};
}

}  // namespace

TORCH_LIBRARY(alonet_custom, m) {
  m.def("ms_deform_attn_forward(Tensor value, Tensor spatial_shapes, Tensor level_start_index, Tensor sampling_loc, "
        "Tensor attn_weight, int im2col_step) -> Tensor");
  m.def("ms_deform_attn_backward(Tensor value, Tensor spatial_shapes, Tensor level_start_index, Tensor sampling_loc, "
        "Tensor attn_weight, Tensor grad_output, int im2col_step) -> Tensor[]");
}
TORCH_LIBRARY_IMPL(alonet_custom, CUDA, m) {
  m.impl("ms_deform_attn_forward", forward_cuda);
  m.impl("ms_deform_attn_backward", backward_cuda);
}
TORCH_LIBRARY_IMPL(alonet_custom, CPU, m) {
  m.impl("ms_deform_attn_forward", forward_cpu);
  m.impl("ms_deform_attn_backward", backward_cpu);
}
TORCH_LIBRARY_IMPL(alonet_custom, Meta, m) {
  m.impl("ms_deform_attn_forward", forward_meta);
  m.impl("ms_deform_attn_backward", backward_meta);
}

extern "C" int msda_torch_shim_version(void) { return MSDA_ABI_VERSION; }
