// msda_capi.cu -- C ABI (include/msda_b200.h) over the sm_100a kernels in msda_kernels.cuh.
//
// Host-side responsibilities mirror the reference launchers
// (alonet/deformable_detr/ops/src/cuda/ms_deform_attn_cuda.cu:20-153 and
//  ms_deform_im2col_cuda.cuh:923-954, 956-1327): validate, pick a kernel for (dtype, D), launch on the
// caller's stream, report errors.  Unlike the reference there is no im2col_step batching loop (one launch
// covers the whole batch with 64-bit indexing), no output memset in forward, and errors are returned
// instead of printed.
#include "../../include/msda_b200.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <type_traits>

#include "msda_kernels.cuh"
#include "msda_bwd_tile.cuh"
#include "msda_fwd_win.cuh"

namespace {

// set for the duration of one msda_forward / msda_backward call: MSDA_LOC_F32 / MSDA_ATTN_F32 of its `dtype` argument
// (fp32 sampling locations / attention weights -- and their gradients -- next to 16-bit value, output and grad_output)
thread_local int t_io32 = 0;
struct Io32Scope {
  explicit Io32Scope(int v) { t_io32 = v; }
  ~Io32Scope() { t_io32 = 0; }
};
size_t loc_elt(size_t e) { return (t_io32 & MSDA_LOC_F32) ? sizeof(float) : e; }
size_t attn_elt(size_t e) { return (t_io32 & MSDA_ATTN_F32) ? sizeof(float) : e; }

// scheduling words of the SM-affine paired forward for the duration of one msda_forward call (caller's workspace), or NULL
thread_local unsigned long long* t_fwd_sched = nullptr;

thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};
std::atomic<int> g_force_generic{0}, g_fwd_unroll{0}, g_bwd_unroll{0}, g_warps_per_block{0}, g_no_pdl{0}, g_head_major{0}, g_smem_records{0}, g_patch_mode{0}, g_patch_px{0}, g_patch_py{0}, g_patch_ctas{0}, g_staged_mode{0}, g_staged_kb{0}, g_staged_warps{0}, g_staged_variant{0}, g_zero_ctas{0}, g_zero_threads{0}, g_zero_mode{0}, g_zero_chunk_kb{0}, g_spec_mode{0}, g_bwd_tile_mode{0}, g_bwd_tile_ctas{0}, g_bwd_two_pass{0}, g_fwd_pair_mode{0}, g_fwd_pair_ctas{0}, g_fwd_pair_px{0}, g_fwd_pair_py{0}, g_fwd_win_mode{0}, g_fwd_win_ctas{0};

int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  const cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();  // clear the (non-sticky) launch error
    return fail("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
  }
  return 0;
}

size_t elt_size(int dtype) {
  switch (dtype) {
    case MSDA_F32: return 4;
    case MSDA_BF16: case MSDA_F16: return 2;
    case MSDA_F64: return 8;
    default: return 0;
  }
}

int validate_dims(const msda_dims* d, int dtype) {
  if (!d) return fail("dims is NULL");
  if (elt_size(dtype) == 0) return fail("unknown dtype %d (0=f32, 1=bf16, 2=f16, 3=f64)", dtype);
  if (d->batch < 0 || d->spatial_size < 0 || d->num_heads < 0 || d->channels < 0 || d->num_levels < 0 ||
      d->num_query < 0 || d->num_point < 0)
    return fail("negative dimension in msda_dims");
  if ((long long)d->num_levels * d->num_point > (1 << 20)) return fail("num_levels * num_point too large");
  return 0;
}

// cudaLaunchKernelEx with the programmatic-stream-serialization attribute (see pdl_wait / pdl_trigger in the kernels)
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

int check_pdl_launch(cudaError_t e, const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
  }
  return 0;
}

bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

struct Launch {
  dim3 grid, block;
  int head_major = 0;
};

// knob: 0 = auto, 1 = force unit-major, 2 = force head-major
int head_major_for(const msda_dims& d);

// one warp per (b, q, m) unit
Launch unit_launch(long long units, int default_warps) {
  int wpb = g_warps_per_block.load(std::memory_order_relaxed);
  if (wpb <= 0 || wpb > 8) wpb = default_warps;
  Launch l;
  l.block = dim3(32 * wpb);
  l.grid = dim3((unsigned)((units + wpb - 1) / wpb));
  return l;
}

// one warp per unit, grid.y = image (the "sg" kernels use 32-bit in-image indexing)
Launch image_launch(const msda_dims& d, int default_warps) {
  int wpb = g_warps_per_block.load(std::memory_order_relaxed);
  const long long qm = (long long)d.num_query * d.num_heads;
  if (wpb <= 0 || wpb > MSDA_MAX_THREADS / 32) {
    // measured (profiles/r1_sweep_tuning.jsonl): calls that fit one wave (<= 36 warps on each of the 148 SMs) start
    // faster with fewer, larger CTAs; multi-wave calls balance better with 2-warp CTAs
    wpb = (qm * d.batch <= 148LL * 36) ? 4 : default_warps;
  }
  Launch l;
  l.block = dim3(32 * wpb);
  l.head_major = head_major_for(d);
  if (l.head_major)  // CTA = one head x wpb consecutive queries
    l.grid = dim3((unsigned)(((long long)d.num_query + wpb - 1) / wpb * d.num_heads), (unsigned)d.batch);
  else
    l.grid = dim3((unsigned)((qm + wpb - 1) / wpb), (unsigned)d.batch);
  return l;
}

// preconditions of the vector kernels: 32-bit in-image offsets with 4 flag bits, grid.y = batch
bool vec_shape_ok(const msda_dims& d) {
  return !g_force_generic.load(std::memory_order_relaxed) && d.batch <= 65535 && d.num_point < (1 << 15) &&
         (long long)d.spatial_size * d.num_heads * d.channels <= (1LL << 27) &&
         (long long)d.num_query * d.num_heads < (1LL << 31) - 64;
}

int head_major_for(const msda_dims& d) {
  (void)d;  // no shape-dependent rule: head-major ordering measured +-1 % (profiles/r1_sweep_head_major.jsonl)
  const int k = g_head_major.load(std::memory_order_relaxed);
  if (k == 1) return 0;
  if (k == 2) return 1;
  return 0;
}

// forward kernel: sample records through shared memory instead of shuffles?
// Measured (profiles/r1_sweep_smem_records.jsonl): fp32 gains 1-5 % everywhere (encoder shapes most: SHFL shares the
// L1 data pipe with the returning tap rows); 16-bit rows gain on one-wave decoder calls (C2 bf16 5.2 -> 4.7 us) and lose
// ~4 % on multi-wave encoder calls, where the two extra CTA-resident kilobytes cost more L1 than the shuffles did.
bool smem_records_auto(const msda_dims& d, size_t elt) {
  if (elt == 4) return true;
  return (long long)d.batch * d.num_query * d.num_heads <= 148LL * 36;
}

// forward: patch-ordered persistent kernel?  Only pays when the queries are the pixels of the pyramid (Lq == S).
bool patch_mode_auto(const msda_dims& d) {
  (void)d;
  return false;
}

// forward: TMA-staged coarse levels (persistent CTAs, one per SM)?  knob "staged_mode": 0 = auto, 1 = off, 2 = on
bool staged_mode_auto(const msda_dims& d) {
  (void)d;
  return false;
}

int sm_count() {  // of the current device
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  std::atomic<int>& slot = cache[dev & 63];
  int n = slot.load(std::memory_order_relaxed);
  if (n == 0) {
    int v = 0;
    n = (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) ? v : 148;
    slot.store(n, std::memory_order_relaxed);
  }
  return n;
}

int pick_unroll(int knob, int fallback) {
  const int u = knob;
  return (u == 1 || u == 2 || u == 4) ? u : fallback;
}

// ---------------------------------------------------------------------------------------------
// forward dispatch
// ---------------------------------------------------------------------------------------------
// SM-affine paired forward (msda_fwd_pair_kernel<.., AFFINE>): the schedule for pixel-aligned queries (encoder self-attention)
bool fwd_affine_shape_ok(const msda_dims& d) {
  return vec_shape_ok(d) && d.channels == 32 && d.num_levels * d.num_point <= 16 && d.num_heads >= 2 &&
         (long long)d.spatial_size * d.num_heads * d.channels * 4 <= (1LL << 29);
}
bool fwd_affine_wanted(const msda_dims& d) {
  const int pkm = g_fwd_pair_mode.load(std::memory_order_relaxed);
  if (!fwd_affine_shape_ok(d) || g_spec_mode.load(std::memory_order_relaxed) == 1) return false;
  if (pkm == 3) return true;
  return false;
}
int zero_sched(unsigned long long* p, cudaStream_t st) {
  const cudaError_t e = cudaMemsetAsync(p, 0, sizeof(unsigned long long) * MSDA_SCHED_WORDS, st);
  return e == cudaSuccess ? 0 : fail("msda_forward(paired, SM-affine): cudaMemsetAsync: %s", cudaGetErrorString(e));
}
template <typename T, int D, int MC, bool FUSED = false>
int launch_fwd_vec(const void* value, const int32_t* shapes, const int32_t* start, const void* loc, const void* attn,
                   void* out, const msda_dims& d, cudaStream_t st, const void* ref = nullptr, int ref_dim = 0) {
  int U = pick_unroll(g_fwd_unroll.load(std::memory_order_relaxed), 1);
  // the rounds of a pass address samples k0 + j*G + g < 32 (one per lane): G * U must not exceed the warp.  16-bit rows of
  // D = 16 channels are covered by 2 lanes (G = 16 lane groups), so the "fwd_unroll = 4" knob is clamped to 2 there
  constexpr int G = 32 / (D / msda::Vec16<T>::N);
  while (G * U > 32) U >>= 1;
  const Launch l = image_launch(d, 2);
  const float inv_p = 1.0f / (float)(d.num_point > 0 ? d.num_point : 1);
  const bool pdl = false;  // forward kernels launch normally (see pdl_wait / pdl_trigger in msda_kernels.cuh)
  // speculative regular-window gather (knob "spec_mode": 0 = auto = on, 1 = off, 2 = on)
  // Measured (profiles/r1_sweep_speculative_gather.jsonl): fp32 +3 % on decoder shapes, +11 % on encoder shapes; 16-bit +2..6 %
  // except multi-wave decoder calls (C4DEC bf16: 47.5 -> 51.5 us, the clamped taps fetch real rows instead of the zero line).
  const int spk = g_spec_mode.load(std::memory_order_relaxed);
  const bool spec_auto = sizeof(T) == 4 || d.num_query >= 2048 || (long long)d.batch * d.num_query * d.num_heads <= 148LL * 36;
  const int io = sizeof(T) == 4 || FUSED ? 0 : t_io32;
  const int spec_on = (spk == 1 ? 0 : (spk == 2 ? 1 : (spec_auto ? 1 : 0))) | ((io & MSDA_LOC_F32) ? 2 : 0) | ((io & MSDA_ATTN_F32) ? 4 : 0);
  // TMA-staged coarse levels, persistent CTAs, one per SM (knob "staged_mode")
  const int smk = g_staged_mode.load(std::memory_order_relaxed);
  // (experimental schedules are instantiated for D = 32 only: every shipped model has 32 channels per head)
  if constexpr (D == 32) if (io == 0 && d.num_levels * d.num_point <= 32 && (smk == 2 || (smk == 0 && staged_mode_auto(d)))) {
    int warps = g_staged_warps.load(std::memory_order_relaxed);
    if (warps <= 0 || warps > 32) warps = 32;
    const int threads = 32 * warps;
    const size_t rowb = (size_t)D * sizeof(T);
    const size_t rec_bytes = ((size_t)threads * 24 + 127) & ~(size_t)127;
    size_t tile_bytes = (size_t)(227 * 1024) - rec_bytes - rowb - 16;
    const int kb = g_staged_kb.load(std::memory_order_relaxed);
    if (kb > 0 && (size_t)kb * 1024 < tile_bytes) tile_bytes = (size_t)kb * 1024;
    long long tile_rows = (long long)(tile_bytes / rowb);
    if (tile_rows > d.spatial_size) tile_rows = d.spatial_size;
    const size_t smem = rec_bytes + (size_t)tile_rows * rowb + rowb + 16;
    const long long items = (long long)d.batch * d.num_heads * ((d.num_query + warps - 1) / warps);
    if (items > 0x7fffffffLL) return fail("msda_forward(staged): too many work items");
    long long ctas = (long long)sm_count();
    if (ctas > items) ctas = items;
    cudaError_t e;
#define MSDA_FWDS(LPC)                                                                                                 \
  do {                                                                                                                 \
    auto kern = msda::msda_fwd_staged_kernel<T, D, MC, FUSED, 1024, 1, LPC>;                                           \
    /* per instantiation AND device (function attributes are per context): raise the dynamic shared-memory limit */    \
    static std::atomic<size_t> smem_set[64];                                                                           \
    int dev_ = 0;                                                                                                      \
    cudaGetDevice(&dev_);                                                                                              \
    std::atomic<size_t>& slot = smem_set[dev_ & 63];                                                                   \
    if (slot.load(std::memory_order_relaxed) < smem) {                                                                 \
      const cudaError_t ae = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);       \
      if (ae != cudaSuccess) return fail("msda_forward(staged): cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(ae)); \
      slot.store(smem, std::memory_order_relaxed);                                                                     \
    }                                                                                                                  \
    e = launch_pdl(kern, dim3((unsigned)ctas), dim3((unsigned)threads), smem, st, pdl, (const T*)value, shapes, start,  \
                   (const T*)loc, (const T*)attn, (T*)out, d.batch, d.spatial_size, d.num_heads, d.num_levels,          \
                   d.num_point, inv_p, d.num_query * d.num_heads, (const T*)ref, ref_dim, (int)tile_rows, spec_on);    \
  } while (0)
    if (d.num_levels * d.num_point == 16 && g_staged_variant.load(std::memory_order_relaxed) != 1) MSDA_FWDS(16); else MSDA_FWDS(0);
#undef MSDA_FWDS
    return check_pdl_launch(e, FUSED ? "msda_fused_forward(staged)" : "msda_forward(staged)");
  }
  // patch-ordered persistent kernel for pixel-aligned queries (knob "patch_mode": 0 = auto, 1 = off, 2 = on)
  const int pmk = g_patch_mode.load(std::memory_order_relaxed);
  if constexpr (D == 32) if (io == 0 && d.num_levels * d.num_point <= 32 && (pmk == 2 || (pmk == 0 && patch_mode_auto(d)))) {  // one sample per lane
    int py = g_patch_py.load(std::memory_order_relaxed), px = g_patch_px.load(std::memory_order_relaxed);
    if (py <= 0 || py > MSDA_PATCH_MAX_THREADS / 32) py = 16;
    if (px <= 0) px = 8;
    int ctas = g_patch_ctas.load(std::memory_order_relaxed);
    if (ctas <= 0) ctas = 1024 / (32 * py);  // 32 warps per SM (the kernel is compiled for <= 64 registers)
    const int srk2 = g_smem_records.load(std::memory_order_relaxed);
    const bool sr2 = srk2 == 2 || (srk2 == 0 && sizeof(T) == 4);
    const dim3 grid((unsigned)(148 * ctas)), block((unsigned)(32 * py));
    cudaError_t e;
#define MSDA_FWDP(SR, MINB)                                                                                    \
  e = launch_pdl(msda::msda_fwd_patch_kernel<T, D, MC, FUSED, SR, MINB>, grid, block, SR ? 24 * block.x : 0, st, pdl,  \
                 (const T*)value, shapes, start, (const T*)loc, (const T*)attn, (T*)out, d.batch, d.spatial_size, \
                 d.num_heads, d.num_levels, d.num_point, inv_p, d.num_query * d.num_heads, (const T*)ref, ref_dim, px, spec_on & 1)
    // (a 40-register instantiation for 48 warps per SM spills and measured 139 us vs 113 us on the encoder shape: dropped)
    if (sr2) MSDA_FWDP(true, 2); else MSDA_FWDP(false, 2);
#undef MSDA_FWDP
    return check_pdl_launch(e, FUSED ? "msda_fused_forward(patch)" : "msda_forward(patch)");
  }
  // records through shared memory (24 B / thread) or through shuffles: knob "smem_records" 0 = auto, 1 = shuffles, 2 = smem
  const int srk = g_smem_records.load(std::memory_order_relaxed);
  const bool sr = U == 1 && (srk == 2 || (srk == 0 && smem_records_auto(d, sizeof(T))));
  cudaError_t e;
  // windowed forward for pixel-aligned queries (msda_fwd_win.cuh; knob "fwd_win_mode": 0 = auto, 1 = off, 2 = on)
  const int wkm = g_fwd_win_mode.load(std::memory_order_relaxed);
  if constexpr (!FUSED && D == 32 && sizeof(T) == 4) if (wkm == 2 && io == 0 && d.num_levels * d.num_point <= 16 && d.num_levels <= MSDA_WIN_LEVELS &&
                                                         (long long)d.spatial_size * d.num_heads * d.channels * 4 <= (1LL << 29)) {
    using Cfg = msda::FwdWinCfg<32, 3, 3, 704>;
    auto kern = msda::msda_fwd_win_kernel<Cfg, MC>;
    static std::atomic<int> smem_set[64];
    int dev_ = 0;
    cudaGetDevice(&dev_);
    std::atomic<int>& slot = smem_set[dev_ & 63];
    if (!slot.load(std::memory_order_relaxed)) {
      const cudaError_t ae = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
      if (ae != cudaSuccess) return fail("msda_forward(windowed): cudaFuncSetAttribute(%zu B): %s", (size_t)Cfg::SMEM, cudaGetErrorString(ae));
      slot.store(1, std::memory_order_relaxed);
    }
    int ctas = g_fwd_win_ctas.load(std::memory_order_relaxed);
    if (ctas <= 0 || ctas > 4) ctas = 2;
    e = launch_pdl(kern, dim3((unsigned)(sm_count() * ctas)), dim3(Cfg::THREADS), Cfg::SMEM, st, false, (const float*)value, shapes, start,
                   (const float*)loc, (const float*)attn, (float*)out, d.batch, d.spatial_size, d.num_heads, d.num_levels, d.num_point,
                   inv_p, d.num_query * d.num_heads);
    return check_pdl_launch(e, "msda_forward(windowed)");
  }
  // paired forward (two heads of a query per warp; knob "fwd_pair_mode": 0 = auto, 1 = off, 2 = static order, 3 = SM-affine
  // patch order -- needs the scheduling words of t_fwd_sched)
  const int pkm = g_fwd_pair_mode.load(std::memory_order_relaxed);
  // Measured and NOT adopted (profiles/r2_sweep_paired_forward_rejected.jsonl): 24 % fewer warp instructions and, SM-affine,
  // 53 % L1 hits instead of 22 %, but 99 / 122 us against 95 us on the encoder shape -- kept behind the knob, bit-identical.
  if constexpr (!FUSED && D == 32) if (U == 1 && (spec_on & 1) && d.num_levels * d.num_point <= 16 && d.num_heads >= 2 && (pkm == 2 || pkm == 3)) {
    const int hp = (d.num_heads + 1) / 2;
    const long long items = (long long)d.num_query * hp;
    const int io_bits = spec_on & 6;
    if (pkm == 3 && t_fwd_sched != nullptr) {
      int pxs = g_fwd_pair_px.load(std::memory_order_relaxed), pys = g_fwd_pair_py.load(std::memory_order_relaxed);
      if (pxs <= 0 || pxs > 6) pxs = 3;
      if (pys <= 0 || pys > 6) pys = 3;
      int ctas = g_fwd_pair_ctas.load(std::memory_order_relaxed);
      if (ctas <= 0 || ctas > 8) ctas = 5;
      const dim3 grid((unsigned)(sm_count() * ctas)), block(256);
      if (int rc = zero_sched(t_fwd_sched, st)) return rc;
#define MSDA_FWDA(SR)                                                                                                      \
  e = launch_pdl(msda::msda_fwd_pair_kernel<T, D, MC, SR, true>, grid, block, SR ? 24 * block.x : 0, st, false,             \
                 (const T*)value, shapes, start, (const T*)loc, (const T*)attn, (T*)out, d.batch, d.spatial_size,           \
                 d.num_heads, d.num_levels, d.num_point, inv_p, d.num_query * d.num_heads, io_bits, t_fwd_sched, pxs, pys)
      if (sr) MSDA_FWDA(true); else MSDA_FWDA(false);
#undef MSDA_FWDA
      return check_pdl_launch(e, "msda_forward(paired, SM-affine)");
    }
    const int wpb = 2;
    const dim3 grid((unsigned)((items + wpb - 1) / wpb), (unsigned)d.batch), block(32 * wpb);
#define MSDA_FWDQ(SR)                                                                                                      \
  e = launch_pdl(msda::msda_fwd_pair_kernel<T, D, MC, SR, false>, grid, block, SR ? 24 * block.x : 0, st, false,            \
                 (const T*)value, shapes, start, (const T*)loc, (const T*)attn, (T*)out, d.batch, d.spatial_size,           \
                 d.num_heads, d.num_levels, d.num_point, inv_p, d.num_query * d.num_heads, io_bits,                         \
                 (unsigned long long*)nullptr, 0, 0)
    if (sr) MSDA_FWDQ(true); else MSDA_FWDQ(false);
#undef MSDA_FWDQ
    return check_pdl_launch(e, "msda_forward(paired)");
  }
#define MSDA_FWD(UU, SR)                                                                                      \
  e = launch_pdl(msda::msda_fwd_sg_kernel<T, D, MC, UU, FUSED, SR>, l.grid, l.block, SR ? 24 * l.block.x : 0, st, pdl, \
                 (const T*)value, shapes, start, (const T*)loc, (const T*)attn, (T*)out, d.spatial_size, d.num_heads,   \
                 d.num_levels, d.num_point, inv_p, d.num_query * d.num_heads, (const T*)ref, ref_dim, l.head_major, spec_on)
  if (U == 1) { if (sr) MSDA_FWD(1, true); else MSDA_FWD(1, false); }
  else if (U == 2) MSDA_FWD(2, false); else MSDA_FWD(4, false);
#undef MSDA_FWD
  return check_pdl_launch(e, FUSED ? "msda_fused_forward" : "msda_forward(vector)");
}

template <typename T>
int launch_fwd_generic(const void* value, const int32_t* shapes, const int32_t* start, const void* loc,
                       const void* attn, void* out, const msda_dims& d, long long units, cudaStream_t st) {
  const Launch l = unit_launch(units, 8);
#define MSDA_FWD_GENERIC(TL, TA)                                                                                    \
  msda::msda_fwd_generic_kernel<T, TL, TA><<<l.grid, l.block, 0, st>>>(                                              \
      (const T*)value, shapes, start, (const TL*)loc, (const TA*)attn, (T*)out, d.spatial_size, d.num_heads,         \
      d.channels, d.num_levels, d.num_query, d.num_point, units)
  if constexpr (sizeof(T) == 2) {  // mixed precision: fp32 locations / weights next to 16-bit value
    const bool l32 = (t_io32 & MSDA_LOC_F32) != 0, a32 = (t_io32 & MSDA_ATTN_F32) != 0;
    if (l32 && a32) MSDA_FWD_GENERIC(float, float);
    else if (l32) MSDA_FWD_GENERIC(float, T);
    else if (a32) MSDA_FWD_GENERIC(T, float);
    else MSDA_FWD_GENERIC(T, T);
  } else {
    MSDA_FWD_GENERIC(T, T);
  }
#undef MSDA_FWD_GENERIC
  return check_launch("msda_forward(generic)");
}

template <typename T>
int forward_typed(const void* value, const int32_t* shapes, const int32_t* start, const void* loc, const void* attn,
                  void* out, const msda_dims& d, long long units, cudaStream_t st) {
  const bool vec_ok = vec_shape_ok(d) && aligned(value, 16) && aligned(out, 16) && aligned(loc, 2 * loc_elt(sizeof(T))) &&
                      aligned(attn, attn_elt(sizeof(T)));
  if (vec_ok) {
#define MSDA_CASE(DD)                                                                                         \
  case DD:                                                                                                    \
    return d.num_heads == 8 ? launch_fwd_vec<T, DD, 8>(value, shapes, start, loc, attn, out, d, st)           \
                            : launch_fwd_vec<T, DD, 0>(value, shapes, start, loc, attn, out, d, st);
    switch (d.channels) {
#ifndef MSDA_DEV_FAST
      MSDA_CASE(16)
      MSDA_CASE(64)
      MSDA_CASE(128)
#endif
      MSDA_CASE(32)
      default: break;
    }
#undef MSDA_CASE
  }
  return launch_fwd_generic<T>(value, shapes, start, loc, attn, out, d, units, st);
}

// ---------------------------------------------------------------------------------------------
// backward dispatch
// ---------------------------------------------------------------------------------------------
// set for the duration of one msda_backward call: the caller zeroed the accumulation buffer itself
// (MSDA_BWD_PREZEROED), so the fill is skipped and the scatter kernel is launched without the programmatic dependency
thread_local bool t_prezeroed = false;

int zero_fill(void* p, size_t bytes, cudaStream_t st) {
  if (bytes == 0) return 0;
  if (!aligned(p, 16)) {  // odd views: let the driver do it
    const cudaError_t e = cudaMemsetAsync(p, 0, bytes, st);
    return e == cudaSuccess ? 0 : fail("cudaMemsetAsync: %s", cudaGetErrorString(e));
  }
  const long long n16 = (long long)(bytes / 16);
  const int ntail = (int)(bytes % 16);
  // every CTA must be resident at once: the dependent backward kernel is released when all of them have started
  int zc = g_zero_ctas.load(std::memory_order_relaxed), zt = g_zero_threads.load(std::memory_order_relaxed);
  // knob "zero_mode": 0 = auto, 1 = 128-bit store kernel, 2 = TMA bulk stores from a zeroed shared-memory buffer.
  // Measured (profiles/r1_sweep_zero_fill.jsonl): the TMA fill loses 0.7-1.7 us on L2-sized fills that the backward
  // kernel overlaps (C2 13.15 -> 13.9 us) and gains 2.5 % on HBM-sized ones (C4DEC, 728 MB: 291.6 -> 284.5 us).
  const int zmode = g_zero_mode.load(std::memory_order_relaxed);
  if (zmode == 2 || (zmode == 0 && bytes > ((size_t)256 << 20))) {
    int ck = g_zero_chunk_kb.load(std::memory_order_relaxed);
    if (ck <= 0 || ck > 200) ck = 32;
    const int chunk = ck * 1024;
    if (zc <= 0 || zc > 16) zc = 1;
    long long blocks = ((long long)bytes + chunk - 1) / chunk;
    if (blocks > 148LL * zc) blocks = 148LL * zc;
    if (chunk > 48 * 1024) {  // beyond the default dynamic shared-memory limit: opt in, per device
      static std::atomic<int> smem_set[64];
      int dev_ = 0;
      cudaGetDevice(&dev_);
      std::atomic<int>& slot = smem_set[dev_ & 63];
      if (slot.load(std::memory_order_relaxed) < chunk) {
        const cudaError_t ae = cudaFuncSetAttribute(msda::msda_zero_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, chunk);
        if (ae != cudaSuccess) return fail("zero fill: cudaFuncSetAttribute: %s", cudaGetErrorString(ae));
        slot.store(chunk, std::memory_order_relaxed);
      }
    }
    const cudaError_t e = launch_pdl(msda::msda_zero_tma_kernel, dim3((unsigned)blocks), dim3(128), (size_t)chunk, st, false,
                                     (unsigned char*)p, n16, ntail, chunk);
    return check_pdl_launch(e, "msda_backward(zero grad_value, TMA)");
  }
  if (zc <= 0 || zc > 16) zc = 4;
  if (zt <= 0 || zt > 256 || (zt & 31)) zt = 256;
  long long blocks = (n16 + zt * 8 - 1) / (zt * 8);
  if (blocks < 1) blocks = 1;
  if (blocks > 148LL * zc) blocks = 148LL * zc;
  const cudaError_t e = launch_pdl(msda::msda_zero_kernel, dim3((unsigned)blocks), dim3((unsigned)zt), 0, st, false, (uint4*)p, n16,
                                   (unsigned char*)p + n16 * 16, ntail);
  return check_pdl_launch(e, "msda_backward(zero grad_value)");
}

// backward: all gather rounds before the fence?  Pays when the call is one wave behind a programmatic zero-fill (the gathers'
// latency chains then hide behind the fill); measured in profiles/r2_sweep_bwd_two_pass.jsonl
bool bwd_two_pass_auto(const msda_dims& d, bool pdl) {
  (void)d; (void)pdl;
  return false;
}

template <typename T, int D, int MC, bool FUSED = false, bool DET = false>
int launch_bwd_vec(const void* go, const void* value, const int32_t* shapes, const int32_t* start, const void* loc,
                   const void* attn, float* gv, void* gloc, void* gattn, const msda_dims& d, cudaStream_t st,
                   const void* ref = nullptr, int ref_dim = 0, float* gref = nullptr, const unsigned* det_hdr = nullptr) {
  const int U = pick_unroll(g_bwd_unroll.load(std::memory_order_relaxed), 1);
  const Launch l = image_launch(d, 2);
  const float inv_p = 1.0f / (float)(d.num_point > 0 ? d.num_point : 1);
  // Programmatic dependent launch: pass 1 of the kernel overlaps the zero-fill that precedes it in the stream;
  // `griddepcontrol.wait` inside the kernel orders pass 2 (the scatter) after the fill.
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = l.grid;
  cfg.blockDim = l.block;
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  // Only when the fill is short (grad_value fits in L2): behind a long fill the early-launched CTAs would just sit on
  // the SMs the fill needs (C4DEC, 728 MB: 298 us with PDL vs 290 us without).
  const size_t fill_bytes = sizeof(float) * (size_t)d.batch * d.spatial_size * d.num_heads * d.channels;
  const bool pdl = !DET && !g_no_pdl.load(std::memory_order_relaxed) && fill_bytes <= (96u << 20) && !t_prezeroed;
  attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const T* go_ = (const T*)go; const T* value_ = (const T*)value; const T* loc_ = (const T*)loc; const T* attn_ = (const T*)attn;
  T* gloc_ = (T*)gloc; T* gattn_ = (T*)gattn; const T* ref_ = (const T*)ref;
  const int S = d.spatial_size, M = d.num_heads, L = d.num_levels, P = d.num_point, QM = d.num_query * d.num_heads;
  // knob "bwd_two_pass": 0 = auto, 1 = single pass (scatter of a round right behind its gather), 2 = gather pass / fence / scatter pass
  const int tpk = g_bwd_two_pass.load(std::memory_order_relaxed);
  const bool two_pass = tpk == 2 || (tpk == 0 && bwd_two_pass_auto(d, pdl));
  const int io = sizeof(T) == 4 || FUSED ? 0 : t_io32;
  const int hm = l.head_major | (two_pass ? 2 : 0) | ((io & MSDA_LOC_F32) ? 4 : 0) | ((io & MSDA_ATTN_F32) ? 8 : 0);
  cudaError_t e;
#define MSDA_BWD(UU)                                                                                          \
  e = cudaLaunchKernelEx(&cfg, msda::msda_bwd_sg_kernel<T, D, MC, UU, FUSED, DET>, go_, value_, shapes, start, loc_, attn_, \
                         gv, gloc_, gattn_, S, M, L, P, inv_p, QM, ref_, ref_dim, gref, hm, det_hdr)
  if constexpr (DET) { MSDA_BWD(1); }
  else { if (U == 1) MSDA_BWD(1); else if (U == 2) MSDA_BWD(2); else MSDA_BWD(4); }
#undef MSDA_BWD
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail("msda_backward(vector): CUDA launch failed: %s", cudaGetErrorString(e));
  }
  return 0;
}

// Tile-binned backward (msda_bwd_tile.cuh): fp32, D = 32, P = 4.  knob "bwd_tile_mode": 0 = auto, 1 = off, 2 = on.
// Auto is OFF: measured on the B200 (profiles/r2_bwd_tile_*.jsonl, r2_ncu_bwd_tile.md) the kernel issues 5 x fewer reds and
// fetches each window row once per tile, but it is instruction-bound (11.4 warp instructions per tap, as many as the
// unit-ordered kernel, which is bound by the red rate instead, at a lower issue rate): ENC 309 us vs 265 us, C5ENC 473 vs 451,
// C4ENC 906 vs 898; random locations 518 vs 292.  It stays as a tested schedule of the same function (any sampling locations,
// any Lq) for callers that want fewer atomics.
constexpr int kTileTPQ = 2, kTileMaxB = 2048;

bool bwd_tile_shape_ok(const msda_dims& d) {
  return d.channels == 32 && d.num_point == 4 && d.num_levels >= 1 && d.num_levels <= 16 && d.spatial_size <= (1 << 19) &&
         (long long)d.spatial_size * d.num_heads * d.channels < (1LL << 29) &&
         d.num_query >= 1 && (long long)d.num_query * d.num_heads * d.num_levels * d.num_point < (1LL << 31) &&
         d.batch <= 65535;
}

bool bwd_tile_auto(const msda_dims& d) {
  (void)d;
  return false;
}

int launch_bwd_tile(const void* go, const void* value, const int32_t* shapes, const int32_t* start, const void* loc,
                    const void* attn, float* gv, void* gloc, void* gattn, const msda_dims& d, cudaStream_t st) {
  using Cfg = msda::BwdTileCfg<32, 4, kTileTPQ, kTileMaxB>;
  auto kern = msda::msda_bwd_tile_kernel<32, 4, kTileTPQ, kTileMaxB>;
  static std::atomic<int> smem_set[64];
  int dev_ = 0;
  cudaGetDevice(&dev_);
  std::atomic<int>& slot = smem_set[dev_ & 63];
  if (!slot.load(std::memory_order_relaxed)) {
    const cudaError_t ae = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES);
    if (ae != cudaSuccess) return fail("msda_backward(tile): cudaFuncSetAttribute(%zu B): %s", (size_t)Cfg::SMEM_BYTES, cudaGetErrorString(ae));
    slot.store(1, std::memory_order_relaxed);
  }
  int per_sm = g_bwd_tile_ctas.load(std::memory_order_relaxed);
  if (per_sm <= 0 || per_sm > 2) per_sm = 2;
  // upper bound of the work items (the level shapes live in device memory): ceil-tiles of every level + tail tiles
  const long long max_tiles = (long long)(d.num_query + 255) / 256 + (long long)d.spatial_size / 256 + 2LL * d.num_levels * 64 + 4;
  long long ctas = (long long)sm_count() * per_sm;
  const long long max_items = max_tiles * d.batch * d.num_heads;
  if (ctas > max_items) ctas = max_items;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)ctas);
  cfg.blockDim = dim3((unsigned)Cfg::THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  const size_t fill_bytes = sizeof(float) * (size_t)d.batch * d.spatial_size * d.num_heads * d.channels;
  const bool pdl = !g_no_pdl.load(std::memory_order_relaxed) && fill_bytes <= (96u << 20) && !t_prezeroed;
  attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, (const float*)go, (const float*)value, shapes, start, (const float*)loc,
                                           (const float*)attn, gv, (float*)gloc, (float*)gattn, d.batch, d.spatial_size,
                                           d.num_heads, d.num_levels, d.num_query);
  return check_pdl_launch(e, "msda_backward(tile)");
}

// Deterministic backward (MSDA_BWD_DETERMINISTIC): int64 fixed-point accumulation of grad_value (msda_kernels.cuh).
// workspace = [256-byte header: max|grad_out|, max|attn| bit patterns][int64 image of grad_value].
constexpr size_t kDetHeader = 256;

template <typename T>
int backward_deterministic(const void* go, const void* value, const int32_t* shapes, const int32_t* start, const void* loc,
                           const void* attn, void* grad_value, void* gloc, void* gattn, void* workspace, const msda_dims& d,
                           long long units, cudaStream_t st) {
  const size_t n_value = (size_t)d.batch * d.spatial_size * d.num_heads * d.channels;
  unsigned* hdr = (unsigned*)workspace;
  long long* img = (long long*)((char*)workspace + kDetHeader);
  if (n_value == 0) return 0;
  cudaError_t ce = cudaMemsetAsync(hdr, 0, kDetHeader, st);
  if (ce != cudaSuccess) return fail("msda_backward(deterministic): cudaMemsetAsync: %s", cudaGetErrorString(ce));
  const bool has_samples = units > 0 && d.num_levels * d.num_point > 0;
  if (has_samples) {
    const long long n_go = units * d.channels, n_attn = units * d.num_levels * d.num_point;
    long long blocks = (n_go + 256 * 8 - 1) / (256 * 8);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (sizeof(T) != 4 && (t_io32 & MSDA_ATTN_F32)) {  // fp32 weights next to 16-bit grad_output: two passes of the same kernel
      msda::msda_absmax_kernel<T><<<(unsigned)blocks, 256, 0, st>>>((const T*)go, n_go, (const T*)nullptr, 0, hdr);
      if (int rc = check_launch("msda_backward(deterministic: absmax)")) return rc;
      msda::msda_absmax_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((const float*)nullptr, 0, (const float*)attn, n_attn, hdr);
    } else {
      msda::msda_absmax_kernel<T><<<(unsigned)blocks, 256, 0, st>>>((const T*)go, n_go, (const T*)attn, n_attn, hdr);
    }
    if (int rc = check_launch("msda_backward(deterministic: absmax)")) return rc;
  }
  if (int rc = zero_fill(img, n_value * sizeof(long long), st)) return rc;
  if (has_samples) {
    int rc = -1;
#define MSDA_CASE(DD)                                                                                                   \
  case DD:                                                                                                              \
    rc = d.num_heads == 8 ? launch_bwd_vec<T, DD, 8, false, true>(go, value, shapes, start, loc, attn, (float*)img, gloc,  \
                                                                  gattn, d, st, nullptr, 0, nullptr, hdr)                \
                          : launch_bwd_vec<T, DD, 0, false, true>(go, value, shapes, start, loc, attn, (float*)img, gloc,  \
                                                                  gattn, d, st, nullptr, 0, nullptr, hdr);               \
    break;
    switch (d.channels) {
#ifndef MSDA_DEV_FAST
      MSDA_CASE(16)
      MSDA_CASE(64)
      MSDA_CASE(128)
#endif
      MSDA_CASE(32)
      default: break;
    }
#undef MSDA_CASE
    if (rc != 0) return rc > 0 ? rc : fail("msda_backward(deterministic): unsupported channel count %d", d.channels);
  }
  long long blocks = (long long)((n_value + 256 * 8 - 1) / (256 * 8));
  if (blocks > 148 * 16) blocks = 148 * 16;
  msda::msda_det_cvt_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(img, (T*)grad_value, (long long)n_value, hdr);
  return check_launch("msda_backward(deterministic: convert grad_value)");
}

bool det_shape_ok(const msda_dims& d, int dtype) {
  return vec_shape_ok(d) && dtype != MSDA_F64 && (d.channels == 16 || d.channels == 32 || d.channels == 64 || d.channels == 128) &&
         (long long)d.num_query * d.num_levels * d.num_point <= (1LL << 25);
}

template <typename T>
int backward_typed(const void* go, const void* value, const int32_t* shapes, const int32_t* start, const void* loc,
                   const void* attn, void* grad_value, void* gloc, void* gattn, void* workspace, const msda_dims& d,
                   long long units, cudaStream_t st) {
  using A = typename msda::AccOf<T>::type;
  const size_t n_value = (size_t)d.batch * d.spatial_size * d.num_heads * d.channels;
  constexpr bool needs_ws = sizeof(T) != sizeof(A);
  A* acc = needs_ws ? (A*)workspace : (A*)grad_value;
  if (!t_prezeroed)
    if (int rc = zero_fill(acc, n_value * sizeof(A), st)) return rc;
  if (units > 0 && d.channels > 0 && d.num_levels * d.num_point > 0) {
    bool done = false;
    if constexpr (std::is_same<T, float>::value) {
      const int tm = g_bwd_tile_mode.load(std::memory_order_relaxed);
      if (tm != 1 && bwd_tile_shape_ok(d) && (tm == 2 || bwd_tile_auto(d)) && !g_force_generic.load(std::memory_order_relaxed) &&
          aligned(value, 16) && aligned(go, 16) && aligned(acc, 16) && aligned(loc, 16) && aligned(attn, 16) &&
          aligned(gloc, 16) && aligned(gattn, 16)) {
        if (int rc = launch_bwd_tile(go, value, shapes, start, loc, attn, (float*)acc, gloc, gattn, d, st)) return rc;
        done = true;
      }
    }
    if constexpr (!std::is_same<T, double>::value) if (!done) {
      const bool vec_ok = vec_shape_ok(d) && aligned(value, 16) && aligned(go, 16) && aligned(acc, 16) &&
                          aligned(loc, 2 * loc_elt(sizeof(T))) && aligned(gloc, 2 * loc_elt(sizeof(T))) &&
                          aligned(attn, attn_elt(sizeof(T))) && aligned(gattn, attn_elt(sizeof(T)));
      if (vec_ok) {
        int rc = -1;
#define MSDA_CASE(DD)                                                                                         \
  case DD:                                                                                                    \
    rc = d.num_heads == 8                                                                                     \
             ? launch_bwd_vec<T, DD, 8>(go, value, shapes, start, loc, attn, (float*)acc, gloc, gattn, d, st)  \
             : launch_bwd_vec<T, DD, 0>(go, value, shapes, start, loc, attn, (float*)acc, gloc, gattn, d, st); \
    break;
        switch (d.channels) {
#ifndef MSDA_DEV_FAST
          MSDA_CASE(16)
          MSDA_CASE(64)
          MSDA_CASE(128)
#endif
          MSDA_CASE(32)
          default: break;
        }
#undef MSDA_CASE
        if (rc > 0) return rc;
        done = rc == 0;
      }
    }
    if (!done) {
      const Launch l = unit_launch(units, 8);
#define MSDA_BWD_GENERIC(TL, TA)                                                                                          \
  msda::msda_bwd_generic_kernel<T, TL, TA><<<l.grid, l.block, 0, st>>>(                                                    \
      (const T*)go, (const T*)value, shapes, start, (const TL*)loc, (const TA*)attn, acc, (TL*)gloc, (TA*)gattn,           \
      d.spatial_size, d.num_heads, d.channels, d.num_levels, d.num_query, d.num_point, units)
      if constexpr (sizeof(T) == 2) {  // mixed precision: fp32 locations / weights (and their gradients) next to 16-bit value
        const bool l32 = (t_io32 & MSDA_LOC_F32) != 0, a32 = (t_io32 & MSDA_ATTN_F32) != 0;
        if (l32 && a32) MSDA_BWD_GENERIC(float, float);
        else if (l32) MSDA_BWD_GENERIC(float, T);
        else if (a32) MSDA_BWD_GENERIC(T, float);
        else MSDA_BWD_GENERIC(T, T);
      } else {
        MSDA_BWD_GENERIC(T, T);
      }
#undef MSDA_BWD_GENERIC
      if (int rc = check_launch("msda_backward(generic)")) return rc;
    }
  }
  if constexpr (needs_ws) {
    if (n_value > 0) {
      long long blocks = (long long)((n_value + 256 * 8 - 1) / (256 * 8));
      if (blocks > 148 * 16) blocks = 148 * 16;
      msda::msda_cvt_kernel<T><<<(unsigned)blocks, 256, 0, st>>>((const float*)acc, (T*)grad_value, (long long)n_value);
      if (int rc = check_launch("msda_backward(convert grad_value)")) return rc;
    }
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// fused operator (softmax + location arithmetic in the kernel); vector kernels only -> rc 2 = "unsupported here"
// ---------------------------------------------------------------------------------------------
constexpr int kUnsupported = 2;

bool fused_shape_ok(const msda_dims& d, int dtype, int ref_dim) {
  return vec_shape_ok(d) && dtype != MSDA_F64 && (ref_dim == 2 || ref_dim == 4) && d.num_levels * d.num_point >= 1 &&
         d.num_levels * d.num_point <= 32 &&
         (d.channels == 16 || d.channels == 32 || d.channels == 64 || d.channels == 128);
}

// `ref_dim` carries the MSDA_FUSED_REF_F32 request in bit 8 on its way to the kernels (fp32 reference points next to 16-bit
// value / offsets / logits; for T = float the bit is dropped: nothing differs)
constexpr int kRef32Bit = 0x100;

template <typename T>
int fused_forward_typed(const void* value, const int32_t* shapes, const int32_t* start, const void* ref, int ref_dim,
                        const void* off, const void* logits, void* out, const msda_dims& d, cudaStream_t st) {
  if (sizeof(T) == 4) ref_dim &= ~kRef32Bit;
  if (!(aligned(value, 16) && aligned(out, 16) && aligned(off, 2 * sizeof(T)) && aligned(logits, sizeof(T)) &&
        aligned(ref, (ref_dim & kRef32Bit) ? sizeof(float) : sizeof(T))))
    return kUnsupported;
#define MSDA_CASE(DD)                                                                                         \
  case DD:                                                                                                    \
    return d.num_heads == 8                                                                                   \
               ? launch_fwd_vec<T, DD, 8, true>(value, shapes, start, off, logits, out, d, st, ref, ref_dim)   \
               : launch_fwd_vec<T, DD, 0, true>(value, shapes, start, off, logits, out, d, st, ref, ref_dim);
  switch (d.channels) {
#ifndef MSDA_DEV_FAST
    MSDA_CASE(16)
    MSDA_CASE(64)
    MSDA_CASE(128)
#endif
    MSDA_CASE(32)
    default: break;
  }
#undef MSDA_CASE
  return kUnsupported;
}

template <typename T>
int fused_backward_typed(const void* go, const void* value, const int32_t* shapes, const int32_t* start, const void* ref,
                         int ref_dim, const void* off, const void* logits, void* grad_value, void* goff, void* glogits,
                         float* gref, void* workspace, const msda_dims& d, cudaStream_t st) {
  constexpr bool needs_ws = sizeof(T) != sizeof(float);
  if (sizeof(T) == 4) ref_dim &= ~kRef32Bit;
  const size_t n_value = (size_t)d.batch * d.spatial_size * d.num_heads * d.channels;
  float* acc = needs_ws ? (float*)workspace : (float*)grad_value;
  if (!(aligned(value, 16) && aligned(go, 16) && aligned(acc, 16) && aligned(off, 2 * sizeof(T)) &&
        aligned(goff, 2 * sizeof(T)) && aligned(logits, sizeof(T)) && aligned(ref, (ref_dim & kRef32Bit) ? sizeof(float) : sizeof(T))))
    return kUnsupported;
  if (int rc = zero_fill(acc, n_value * sizeof(float), st)) return rc;
  int rc = kUnsupported;
#define MSDA_CASE(DD)                                                                                                  \
  case DD:                                                                                                             \
    rc = d.num_heads == 8 ? launch_bwd_vec<T, DD, 8, true>(go, value, shapes, start, off, logits, acc, goff, glogits, d, \
                                                           st, ref, ref_dim, gref)                                      \
                          : launch_bwd_vec<T, DD, 0, true>(go, value, shapes, start, off, logits, acc, goff, glogits, d, \
                                                           st, ref, ref_dim, gref);                                     \
    break;
  switch (d.channels) {
#ifndef MSDA_DEV_FAST
    MSDA_CASE(16)
    MSDA_CASE(64)
    MSDA_CASE(128)
#endif
    MSDA_CASE(32)
    default: break;
  }
#undef MSDA_CASE
  if (rc != 0) return rc;
  if constexpr (needs_ws) {
    if (n_value > 0) {
      long long blocks = (long long)((n_value + 256 * 8 - 1) / (256 * 8));
      if (blocks > 148 * 16) blocks = 148 * 16;
      msda::msda_cvt_kernel<T><<<(unsigned)blocks, 256, 0, st>>>((const float*)acc, (T*)grad_value, (long long)n_value);
      if (int rc2 = check_launch("msda_fused_backward(convert grad_value)")) return rc2;
    }
  }
  return 0;
}


// Large batches: image chunks whose accumulation image stays in L2 between its zero-fill and its scatter.
// The backward of a call whose grad_value does not fit the 126 MB L2 (C4DEC, N = 32: 728 MB) pays DRAM three times for
// every touched row -- the fill writes it, the first red fetches it again (evicted meanwhile), the eviction writes it back.
// Images are independent, so the call is issued as consecutive (zero-fill, scatter[, convert])
// groups of `nb` images with <= 64 MB of accumulation image each: the reds then land on L2-resident lines and every line
// goes to DRAM once.  (The reference batches its launches the same way for another reason: im2col_step,
// ms_deform_attn_cuda.cu:126-150.)  knob "bwd_chunk_mb": 0 = auto (64), -1 = off, else the budget in MB.
// Measured (profiles/r2_bwd_chunk_and_deterministic.jsonl, C4DEC shapes): N = 8 (182 MB) 77.8 -> 70.7 us, N = 16 (364 MB)
// 145.7 -> 141.0, N = 32 (728 MB) 282.8 -> 282.5: the groups of a very large call are one-wave, latency-bound launches
// whose sum equals the whole call, so auto applies the grouping up to 400 MB only.
std::atomic<int> g_bwd_chunk_mb{0};

template <typename T>
int backward_chunked(const void* go, const void* value, const int32_t* shapes, const int32_t* start, const void* loc,
                     const void* attn, void* grad_value, void* gloc, void* gattn, void* workspace, const msda_dims& d,
                     long long units, cudaStream_t st) {
  using A = typename msda::AccOf<T>::type;
  const int knob = g_bwd_chunk_mb.load(std::memory_order_relaxed);
  const size_t per_image = sizeof(A) * (size_t)d.spatial_size * d.num_heads * d.channels;
  const size_t budget = (size_t)(knob > 0 ? knob : 64) << 20;
  const bool whole = knob < 0 || t_prezeroed || d.batch <= 1 || per_image == 0 || per_image * d.batch <= (96u << 20) || units == 0 ||
                     (knob == 0 && per_image * d.batch > ((size_t)400 << 20));
  if (whole) return backward_typed<T>(go, value, shapes, start, loc, attn, grad_value, gloc, gattn, workspace, d, units, st);
  int nb = (int)(budget / per_image);
  if (nb < 1) nb = 1;
  const size_t e = sizeof(T);
  const size_t s_value = (size_t)d.spatial_size * d.num_heads * d.channels, s_out = (size_t)d.num_query * d.num_heads * d.channels,
               s_attn = (size_t)d.num_query * d.num_heads * d.num_levels * d.num_point;
  for (int b0 = 0; b0 < d.batch; b0 += nb) {
    msda_dims c = d;
    c.batch = d.batch - b0 < nb ? d.batch - b0 : nb;
    auto at = [&](const void* p, size_t stride, size_t elt) { return p ? (const char*)p + (size_t)b0 * stride * elt : nullptr; };
    const int rc = backward_typed<T>(at(go, s_out, e), at(value, s_value, e), shapes, start, at(loc, 2 * s_attn, loc_elt(e)),
                                     at(attn, s_attn, attn_elt(e)), (void*)at(grad_value, s_value, e),
                                     (void*)at(gloc, 2 * s_attn, loc_elt(e)), (void*)at(gattn, s_attn, attn_elt(e)),
                                     (void*)at(workspace, s_value, sizeof(A)), c, (long long)c.batch * d.num_query * d.num_heads, st);
    if (rc) return rc;
  }
  return 0;
}

}  // namespace

extern "C" {


int msda_version(void) { return MSDA_ABI_VERSION; }

const char* msda_last_error_string(void) { return g_err; }

uint64_t msda_kernel_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

static std::atomic<int>* knob(const char* name) {
  if (!name) return nullptr;
  if (!strcmp(name, "force_generic")) return &g_force_generic;
  if (!strcmp(name, "fwd_unroll")) return &g_fwd_unroll;
  if (!strcmp(name, "bwd_unroll")) return &g_bwd_unroll;
  if (!strcmp(name, "warps_per_block")) return &g_warps_per_block;
  if (!strcmp(name, "no_pdl")) return &g_no_pdl;
  if (!strcmp(name, "head_major")) return &g_head_major;
  if (!strcmp(name, "smem_records")) return &g_smem_records;
  if (!strcmp(name, "patch_mode")) return &g_patch_mode;
  if (!strcmp(name, "patch_px")) return &g_patch_px;
  if (!strcmp(name, "patch_py")) return &g_patch_py;
  if (!strcmp(name, "patch_ctas")) return &g_patch_ctas;
  if (!strcmp(name, "zero_ctas")) return &g_zero_ctas;
  if (!strcmp(name, "spec_mode")) return &g_spec_mode;
  if (!strcmp(name, "zero_mode")) return &g_zero_mode;
  if (!strcmp(name, "zero_chunk_kb")) return &g_zero_chunk_kb;
  if (!strcmp(name, "zero_threads")) return &g_zero_threads;
  if (!strcmp(name, "staged_mode")) return &g_staged_mode;
  if (!strcmp(name, "staged_kb")) return &g_staged_kb;
  if (!strcmp(name, "staged_variant")) return &g_staged_variant;
  if (!strcmp(name, "staged_warps")) return &g_staged_warps;
  if (!strcmp(name, "bwd_tile_mode")) return &g_bwd_tile_mode;
  if (!strcmp(name, "bwd_tile_ctas")) return &g_bwd_tile_ctas;
  if (!strcmp(name, "bwd_two_pass")) return &g_bwd_two_pass;
  if (!strcmp(name, "fwd_win_mode")) return &g_fwd_win_mode;
  if (!strcmp(name, "fwd_win_ctas")) return &g_fwd_win_ctas;
  if (!strcmp(name, "fwd_pair_mode")) return &g_fwd_pair_mode;
  if (!strcmp(name, "fwd_pair_ctas")) return &g_fwd_pair_ctas;
  if (!strcmp(name, "fwd_pair_px")) return &g_fwd_pair_px;
  if (!strcmp(name, "fwd_pair_py")) return &g_fwd_pair_py;
  if (!strcmp(name, "bwd_chunk_mb")) return &g_bwd_chunk_mb;
  return nullptr;
}

int msda_set_tuning(const char* name, int value) {
  std::atomic<int>* k = knob(name);
  if (!k) return fail("unknown tuning knob '%s'", name ? name : "(null)");
  k->store(value, std::memory_order_relaxed);
  return 0;
}

int msda_get_tuning(const char* name, int* value) {
  std::atomic<int>* k = knob(name);
  if (!k || !value) return fail("unknown tuning knob '%s'", name ? name : "(null)");
  *value = k->load(std::memory_order_relaxed);
  return 0;
}

size_t msda_forward_workspace_bytes(const msda_dims* dims, int dtype) {
  if (!dims) return 0;
  dtype &= ~(MSDA_LOC_F32 | MSDA_ATTN_F32);
  if (dtype != MSDA_F32 && dtype != MSDA_BF16 && dtype != MSDA_F16) return 0;
  return fwd_affine_wanted(*dims) ? sizeof(unsigned long long) * MSDA_SCHED_WORDS : 0;
}

int msda_forward(const void* value, const int32_t* spatial_shapes, const int32_t* level_start_index,
                 const void* sampling_loc, const void* attn_weight, void* output, const msda_dims* dims, int dtype,
                 void* stream) {
  return msda_forward_ws(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, output, nullptr, 0, dims, dtype, stream);
}

int msda_forward_ws(const void* value, const int32_t* spatial_shapes, const int32_t* level_start_index,
                    const void* sampling_loc, const void* attn_weight, void* output, void* workspace, size_t workspace_bytes,
                    const msda_dims* dims, int dtype, void* stream) {
  g_err[0] = 0;
  struct SchedScope {  // scoped: every return path clears it
    explicit SchedScope(unsigned long long* p) { t_fwd_sched = p; }
    ~SchedScope() { t_fwd_sched = nullptr; }
  } sched_scope(workspace != nullptr && workspace_bytes >= sizeof(unsigned long long) * MSDA_SCHED_WORDS && aligned(workspace, 8)
                    ? (unsigned long long*)workspace : nullptr);
  const int io = dtype & (MSDA_LOC_F32 | MSDA_ATTN_F32);
  dtype &= ~(MSDA_LOC_F32 | MSDA_ATTN_F32);
  if (int rc = validate_dims(dims, dtype)) return rc;
  if (io && dtype == MSDA_F64) return fail("MSDA_LOC_F32 / MSDA_ATTN_F32 make no sense for MSDA_F64");
  Io32Scope io_scope(dtype == MSDA_F32 ? 0 : io);
  const msda_dims& d = *dims;
  const long long units = (long long)d.batch * d.num_query * d.num_heads;
  if (units == 0 || d.channels == 0) return 0;  // empty output
  if (!output) return fail("output is NULL");
  const bool has_samples = d.num_levels > 0 && d.num_point > 0;
  if (has_samples && (!value || !spatial_shapes || !level_start_index || !sampling_loc || !attn_weight))
    return fail("NULL tensor pointer passed to msda_forward");
  if ((units + 7) / 8 > 0x7fffffffLL) return fail("problem too large for one launch");
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case MSDA_F32: return forward_typed<float>(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, output, d, units, st);
#ifndef MSDA_DEV_FAST
    case MSDA_BF16: return forward_typed<__nv_bfloat16>(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, output, d, units, st);
    case MSDA_F16: return forward_typed<__half>(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, output, d, units, st);
#endif
    case MSDA_F64: return launch_fwd_generic<double>(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, output, d, units, st);
  }
  return fail("unreachable");
}

size_t msda_backward_workspace_bytes_ex(const msda_dims* dims, int dtype, int flags) {
  if (!dims) return 0;
  dtype &= ~(MSDA_LOC_F32 | MSDA_ATTN_F32);
  const size_t n_value = (size_t)dims->batch * dims->spatial_size * dims->num_heads * dims->channels;
  if (flags & MSDA_BWD_DETERMINISTIC) return n_value ? kDetHeader + sizeof(long long) * n_value : 0;
  if (dtype == MSDA_BF16 || dtype == MSDA_F16) return sizeof(float) * n_value;
  return 0;
}

size_t msda_backward_workspace_bytes(const msda_dims* dims, int dtype) { return msda_backward_workspace_bytes_ex(dims, dtype, 0); }

int msda_backward(const void* grad_output, const void* value, const int32_t* spatial_shapes,
                  const int32_t* level_start_index, const void* sampling_loc, const void* attn_weight,
                  void* grad_value, void* grad_sampling_loc, void* grad_attn_weight, void* workspace,
                  size_t workspace_bytes, const msda_dims* dims, int dtype, int flags, void* stream) {
  g_err[0] = 0;
  const int io = dtype & (MSDA_LOC_F32 | MSDA_ATTN_F32);
  dtype &= ~(MSDA_LOC_F32 | MSDA_ATTN_F32);
  if (int rc = validate_dims(dims, dtype)) return rc;
  if (io && dtype == MSDA_F64) return fail("MSDA_LOC_F32 / MSDA_ATTN_F32 make no sense for MSDA_F64");
  Io32Scope io_scope(dtype == MSDA_F32 ? 0 : io);
  if (flags & ~(MSDA_BWD_PREZEROED | MSDA_BWD_DETERMINISTIC)) return fail("unknown flags 0x%x", flags);
  const bool det = (flags & MSDA_BWD_DETERMINISTIC) != 0;
  if (det && (flags & MSDA_BWD_PREZEROED)) return fail("MSDA_BWD_DETERMINISTIC cannot be combined with MSDA_BWD_PREZEROED");
  struct Prezeroed {  // scoped: every return path clears it
    explicit Prezeroed(bool v) { t_prezeroed = v; }
    ~Prezeroed() { t_prezeroed = false; }
  } prezeroed_scope((flags & MSDA_BWD_PREZEROED) != 0);
  const msda_dims& d = *dims;
  const long long units = (long long)d.batch * d.num_query * d.num_heads;
  const size_t n_value = (size_t)d.batch * d.spatial_size * d.num_heads * d.channels;
  const size_t need = msda_backward_workspace_bytes_ex(dims, dtype, flags);
  if (need > 0 && n_value > 0 && (!workspace || workspace_bytes < need))
    return fail("workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
  if (n_value > 0 && !grad_value) return fail("grad_value is NULL");
  const bool has_samples = units > 0 && d.num_levels > 0 && d.num_point > 0;
  if (has_samples && (!grad_output || !value || !spatial_shapes || !level_start_index || !sampling_loc ||
                      !attn_weight || !grad_sampling_loc || !grad_attn_weight))
    return fail("NULL tensor pointer passed to msda_backward");
  if ((units + 7) / 8 > 0x7fffffffLL) return fail("problem too large for one launch");
  cudaStream_t st = (cudaStream_t)stream;
  if (det) {
    const bool al = aligned(value, 16) && aligned(grad_output, 16) && aligned(workspace, 16) && aligned(sampling_loc, 8) &&
                    aligned(grad_sampling_loc, 8);
    if (!det_shape_ok(d, dtype) || (has_samples && !al))
      return fail("MSDA_BWD_DETERMINISTIC is served by the vector kernels only (f32 / bf16 / f16, D in {16, 32, 64, 128}, "
                  "16-byte aligned tensors, Lq * L * P <= 2^25)");
    switch (dtype) {
      case MSDA_F32: return backward_deterministic<float>(grad_output, value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_value, grad_sampling_loc, grad_attn_weight, workspace, d, units, st);
#ifndef MSDA_DEV_FAST
      case MSDA_BF16: return backward_deterministic<__nv_bfloat16>(grad_output, value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_value, grad_sampling_loc, grad_attn_weight, workspace, d, units, st);
      case MSDA_F16: return backward_deterministic<__half>(grad_output, value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_value, grad_sampling_loc, grad_attn_weight, workspace, d, units, st);
#endif
      default: return fail("MSDA_BWD_DETERMINISTIC: unsupported dtype %d", dtype);
    }
  }
  switch (dtype) {
    case MSDA_F32: return backward_chunked<float>(grad_output, value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_value, grad_sampling_loc, grad_attn_weight, workspace, d, units, st);
#ifndef MSDA_DEV_FAST
    case MSDA_BF16: return backward_chunked<__nv_bfloat16>(grad_output, value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_value, grad_sampling_loc, grad_attn_weight, workspace, d, units, st);
    case MSDA_F16: return backward_chunked<__half>(grad_output, value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_value, grad_sampling_loc, grad_attn_weight, workspace, d, units, st);
#endif
    case MSDA_F64: return backward_chunked<double>(grad_output, value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_value, grad_sampling_loc, grad_attn_weight, workspace, d, units, st);
  }
  return fail("unreachable");
}

int msda_zero_fill(void* ptr, size_t bytes, void* stream) {
  g_err[0] = 0;
  if (bytes > 0 && !ptr) return fail("NULL pointer passed to msda_zero_fill");
  const bool keep = t_prezeroed;
  t_prezeroed = false;
  const int rc = zero_fill(ptr, bytes, (cudaStream_t)stream);
  t_prezeroed = keep;
  return rc;
}

int msda_forward_host(const void* value, const int32_t* spatial_shapes, const int32_t* level_start_index,
                      const void* sampling_loc, const void* attn_weight, void* output, const msda_dims* dims, int dtype,
                      void* stream) {
  g_err[0] = 0;
  if (int rc = validate_dims(dims, dtype)) return rc;
  const msda_dims& d = *dims;
  const size_t e = elt_size(dtype);
  const size_t b_value = e * (size_t)d.batch * d.spatial_size * d.num_heads * d.channels;
  const size_t b_attn = e * (size_t)d.batch * d.num_query * d.num_heads * d.num_levels * d.num_point;
  const size_t b_loc = 2 * b_attn;
  const size_t b_out = e * (size_t)d.batch * d.num_query * d.num_heads * d.channels;
  const size_t b_shapes = sizeof(int32_t) * 2 * (size_t)d.num_levels, b_start = sizeof(int32_t) * (size_t)d.num_levels;
  if (b_out == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t o_value = 0, o_loc = o_value + up(b_value), o_attn = o_loc + up(b_loc), o_out = o_attn + up(b_attn),
               o_shapes = o_out + up(b_out), o_start = o_shapes + up(b_shapes), total = o_start + up(b_start);
  char* dev = nullptr;
  cudaError_t ce = cudaMallocAsync((void**)&dev, total, st);
  if (ce != cudaSuccess) return fail("cudaMallocAsync(%zu): %s", total, cudaGetErrorString(ce));
  int rc = 0;
  auto h2d = [&](size_t off, const void* src, size_t n) {
    if (rc == 0 && n > 0) {
      if (!src) { rc = fail("NULL host pointer passed to msda_forward_host"); return; }
      const cudaError_t c = cudaMemcpyAsync(dev + off, src, n, cudaMemcpyHostToDevice, st);
      if (c != cudaSuccess) rc = fail("cudaMemcpyAsync H2D: %s", cudaGetErrorString(c));
    }
  };
  h2d(o_value, value, b_value);
  h2d(o_loc, sampling_loc, b_loc);
  h2d(o_attn, attn_weight, b_attn);
  h2d(o_shapes, spatial_shapes, b_shapes);
  h2d(o_start, level_start_index, b_start);
  if (rc == 0)
    rc = msda_forward(dev + o_value, (const int32_t*)(dev + o_shapes), (const int32_t*)(dev + o_start), dev + o_loc,
                      dev + o_attn, dev + o_out, dims, dtype, stream);
  if (rc == 0) {
    if (!output) rc = fail("output is NULL");
    else {
      const cudaError_t c = cudaMemcpyAsync(output, dev + o_out, b_out, cudaMemcpyDeviceToHost, st);
      if (c != cudaSuccess) rc = fail("cudaMemcpyAsync D2H: %s", cudaGetErrorString(c));
    }
  }
  cudaFreeAsync(dev, st);
  const cudaError_t se = cudaStreamSynchronize(st);
  if (rc == 0 && se != cudaSuccess) rc = fail("stream synchronize: %s", cudaGetErrorString(se));
  return rc;
}

int msda_im2col_inference(void* stream, const void* data_value, const void* data_spatial_shapes,
                          const void* data_level_start_index, const void* data_sampling_loc,
                          const void* data_attn_weight, int batch_size, int spatial_size, int num_heads, int channels,
                          int num_levels, int num_query, int num_point, void* data_col, int data_type) {
  int dtype;
  switch (data_type) {  // nvinfer1::DataType
    case 0: dtype = MSDA_F32; break;
    case 1: dtype = MSDA_F16; break;
    default:
      fail("msda_im2col_inference: unsupported nvinfer1::DataType %d (0 = kFLOAT, 1 = kHALF)", data_type);
      return -1;
  }
  const msda_dims d = {batch_size, spatial_size, num_heads, channels, num_levels, num_query, num_point};
  const int rc = msda_forward(data_value, (const int32_t*)data_spatial_shapes, (const int32_t*)data_level_start_index,
                              data_sampling_loc, data_attn_weight, data_col, &d, dtype, stream);
  return rc == 0 ? 0 : -1;
}

int msda_fused_supported(const msda_dims* dims, int dtype, int ref_dim) {
  if (!dims || elt_size(dtype) == 0) return 0;
  return fused_shape_ok(*dims, dtype, ref_dim) ? 1 : 0;
}

int msda_fused_forward(const void* value, const int32_t* spatial_shapes, const int32_t* level_start_index,
                       const void* reference_points, int ref_dim, const void* sampling_offsets, const void* attn_logits,
                       void* output, const msda_dims* dims, int dtype, void* stream) {
  return msda_fused_forward_ex(value, spatial_shapes, level_start_index, reference_points, ref_dim, sampling_offsets, attn_logits,
                               output, dims, dtype, 0, stream);
}

int msda_fused_forward_ex(const void* value, const int32_t* spatial_shapes, const int32_t* level_start_index,
                          const void* reference_points, int ref_dim, const void* sampling_offsets, const void* attn_logits,
                          void* output, const msda_dims* dims, int dtype, int flags, void* stream) {
  g_err[0] = 0;
  if (int rc = validate_dims(dims, dtype)) return rc;
  if (flags & ~MSDA_FUSED_REF_F32) return fail("msda_fused_forward_ex: unknown flags 0x%x", flags);
  const msda_dims& d = *dims;
  if (!fused_shape_ok(d, dtype, ref_dim)) {
    fail("msda_fused_forward: shape/dtype outside the fused kernels (need L*P <= 32, D in {16,32,64,128}, no f64)");
    return kUnsupported;
  }
  const long long units = (long long)d.batch * d.num_query * d.num_heads;
  if (units == 0) return 0;
  if (!value || !spatial_shapes || !level_start_index || !reference_points || !sampling_offsets || !attn_logits || !output)
    return fail("NULL tensor pointer passed to msda_fused_forward");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = kUnsupported;
  if (flags & MSDA_FUSED_REF_F32) ref_dim |= kRef32Bit;
  switch (dtype) {
    case MSDA_F32: rc = fused_forward_typed<float>(value, spatial_shapes, level_start_index, reference_points, ref_dim, sampling_offsets, attn_logits, output, d, st); break;
#ifndef MSDA_DEV_FAST
    case MSDA_BF16: rc = fused_forward_typed<__nv_bfloat16>(value, spatial_shapes, level_start_index, reference_points, ref_dim, sampling_offsets, attn_logits, output, d, st); break;
    case MSDA_F16: rc = fused_forward_typed<__half>(value, spatial_shapes, level_start_index, reference_points, ref_dim, sampling_offsets, attn_logits, output, d, st); break;
#endif
    default: break;
  }
  if (rc == kUnsupported) fail("msda_fused_forward: misaligned tensors");
  return rc;
}

int msda_fused_backward(const void* grad_output, const void* value, const int32_t* spatial_shapes,
                        const int32_t* level_start_index, const void* reference_points, int ref_dim,
                        const void* sampling_offsets, const void* attn_logits, void* grad_value, void* grad_offsets,
                        void* grad_logits, float* grad_reference_points, void* workspace, size_t workspace_bytes,
                        const msda_dims* dims, int dtype, int flags, void* stream) {
  g_err[0] = 0;
  if (int rc = validate_dims(dims, dtype)) return rc;
  if (flags & ~MSDA_FUSED_REF_F32) return fail("msda_fused_backward: unknown flags 0x%x", flags);
  const msda_dims& d = *dims;
  if (!fused_shape_ok(d, dtype, ref_dim)) {
    fail("msda_fused_backward: shape/dtype outside the fused kernels");
    return kUnsupported;
  }
  const long long units = (long long)d.batch * d.num_query * d.num_heads;
  const size_t n_value = (size_t)d.batch * d.spatial_size * d.num_heads * d.channels;
  const size_t need = msda_backward_workspace_bytes(dims, dtype);
  if (need > 0 && n_value > 0 && (!workspace || workspace_bytes < need))
    return fail("workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
  if (n_value > 0 && !grad_value) return fail("grad_value is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  if (units == 0) return zero_fill(need ? workspace : grad_value, n_value * sizeof(float), st) ||
                         (need ? zero_fill(grad_value, n_value * elt_size(dtype), st) : 0);
  if (!grad_output || !value || !spatial_shapes || !level_start_index || !reference_points || !sampling_offsets ||
      !attn_logits || !grad_offsets || !grad_logits)
    return fail("NULL tensor pointer passed to msda_fused_backward");
  int rc = kUnsupported;
  if (flags & MSDA_FUSED_REF_F32) ref_dim |= kRef32Bit;
  switch (dtype) {
    case MSDA_F32: rc = fused_backward_typed<float>(grad_output, value, spatial_shapes, level_start_index, reference_points, ref_dim, sampling_offsets, attn_logits, grad_value, grad_offsets, grad_logits, grad_reference_points, workspace, d, st); break;
#ifndef MSDA_DEV_FAST
    case MSDA_BF16: rc = fused_backward_typed<__nv_bfloat16>(grad_output, value, spatial_shapes, level_start_index, reference_points, ref_dim, sampling_offsets, attn_logits, grad_value, grad_offsets, grad_logits, grad_reference_points, workspace, d, st); break;
    case MSDA_F16: rc = fused_backward_typed<__half>(grad_output, value, spatial_shapes, level_start_index, reference_points, ref_dim, sampling_offsets, attn_logits, grad_value, grad_offsets, grad_logits, grad_reference_points, workspace, d, st); break;
#endif
    default: break;
  }
  if (rc == kUnsupported) fail("msda_fused_backward: misaligned tensors");
  return rc;
}

}  // extern "C"
