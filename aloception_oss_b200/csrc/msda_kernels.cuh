// msda_kernels.cuh -- sm_100a kernels of the multi-scale deformable attention operator.
//
// Maths (reference: alonet/deformable_detr/ops/src/cuda/ms_deform_im2col_cuda.cuh:33-159,237-299):
//   out[b,q,m,:] = sum_{l,p} A[b,q,m,l,p] * bilinear(V_l[b,:,m,:], x = loc_x*W_l - 0.5, y = loc_y*H_l - 0.5)
// with zero padding per tap and the whole sample dropped unless -1 < y < H_l and -1 < x < W_l.
//
// Work decomposition (not the reference's one-thread-per-output-scalar):
//   * one WARP owns one unit (b, q, m);
//   * vector kernels: a row of D channels is covered by LPR = D*sizeof(T)/16 lanes with one 128-bit load
//     each; the 32/LPR lane groups of the warp take different (level, point) samples, so one warp-wide
//     load instruction fetches 32/LPR complete tap rows; partial sums are combined across groups with
//     warp shuffles and written with one 128-bit store per lane of group 0;
//   * generic kernels: lanes stride over channels, any D / dtype (incl. double).
// The kernels are HBM/L2-latency bound gathers (<1 flop per byte): no tensor cores by design.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef MSDA_MAX_THREADS
#define MSDA_MAX_THREADS 256  // warps_per_block <= 8
#endif
#ifndef MSDA_FWD_MIN_CTAS  // min resident 256-thread CTAs per SM of the U=1 forward kernel: 6 -> 40 registers, 5 -> 48
#define MSDA_FWD_MIN_CTAS 6
#endif

// Register cap of the U=1 backward kernel.  56 = nine 128-thread CTAs per SM = 1 332 resident CTAs: the 1 200 CTAs of a
// C2-sized call (N=2, 300 queries, 8 heads; one warp per unit) run as ONE wave.  At 57..64 registers only eight fit (1 184)
// and the call takes a second, nearly empty wave (measured: backward 13.1 -> 14.7 us).
#ifndef MSDA_BWD_MAX_REGS
#define MSDA_BWD_MAX_REGS 56
#endif
#ifndef MSDA_BWD_MAX_REGS_FUSED  // the fused kernel (softmax + location arithmetic inside) spills 16 bytes at 56
#define MSDA_BWD_MAX_REGS_FUSED 56
#endif

namespace msda {

// ------------------------------------------------------------------------------------------------
// scalar conversions
// ------------------------------------------------------------------------------------------------
template <typename T> struct AccOf { using type = float; };
template <> struct AccOf<double> { using type = double; };

__device__ __forceinline__ float to_acc(float v) { return v; }
__device__ __forceinline__ double to_acc(double v) { return v; }
__device__ __forceinline__ float to_acc(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float to_acc(__half v) { return __half2float(v); }

template <typename T> __device__ __forceinline__ T from_acc(typename AccOf<T>::type v);
template <> __device__ __forceinline__ float from_acc<float>(float v) { return v; }
template <> __device__ __forceinline__ double from_acc<double>(double v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_acc<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_acc<__half>(float v) { return __float2half_rn(v); }

// ------------------------------------------------------------------------------------------------
// geometry of one sampling point inside one level
// ------------------------------------------------------------------------------------------------
template <typename A>
struct Geo {
  A hy, hx, ly, lx;  // fractional parts (l*) and complements (h*)
  int row00;         // y0 * W + x0 (may be "negative-ish"; only used for valid taps)
  int W;
  bool ok00, ok01, ok10, ok11;  // tap inside the level AND sample inside the window
  bool inside;                  // sample inside the (-1, size) window; false also for NaN / Inf coordinates
};

// `valid` = this lane group really has a sample (tail predicate).
template <typename A>
__device__ __forceinline__ Geo<A> make_geo(A loc_x, A loc_y, int H, int W, bool valid) {
  Geo<A> g;
  // ONE rounding (fused multiply-add): this is what the reference's `loc * size - 0.5` compiles to with nvcc's
  // default -fmad=true (verified on the B200: tests/debug_gradloc.py), and it is the closest fp32 gets to the
  // exact coordinate -- floor() decides which pixel pair is interpolated, and grad_loc jumps across that decision.
  const A y = fma(loc_y, (A)H, (A)-0.5);
  const A x = fma(loc_x, (A)W, (A)-0.5);
  const bool inside = valid && y > (A)-1 && x > (A)-1 && y < (A)H && x < (A)W;
  const A fy = floor(y), fx = floor(x);
  const int y0 = (int)fy, x0 = (int)fx;
  g.ly = y - fy;
  g.lx = x - fx;
  g.hy = (A)1 - g.ly;
  g.hx = (A)1 - g.lx;
  g.W = W;
  g.row00 = y0 * W + x0;
  const bool y0ok = y0 >= 0, x0ok = x0 >= 0, y1ok = y0 + 1 <= H - 1, x1ok = x0 + 1 <= W - 1;
  g.ok00 = inside && y0ok && x0ok;
  g.ok01 = inside && y0ok && x1ok;
  g.ok10 = inside && y1ok && x0ok;
  g.ok11 = inside && y1ok && x1ok;
  g.inside = inside;
  if (!inside) {  // the reference skips such a sample altogether: no NaN / Inf coordinate may leak into a 0 * x product
    g.ly = g.lx = g.hy = g.hx = (A)0;
    g.row00 = 0;
  }
  return g;
}

// ------------------------------------------------------------------------------------------------
// 128-bit helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg128(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

template <typename T> struct Vec16;  // 16 bytes of T <-> floats
template <> struct Vec16<float> {
  static constexpr int N = 4;
  __device__ static __forceinline__ void unpack(const uint4& r, float (&f)[4]) {
    f[0] = __uint_as_float(r.x); f[1] = __uint_as_float(r.y); f[2] = __uint_as_float(r.z); f[3] = __uint_as_float(r.w);
  }
  __device__ static __forceinline__ uint4 pack(const float (&f)[4]) {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
  }
};
template <> struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static __forceinline__ void unpack(const uint4& r, float (&f)[8]) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // bf16 -> f32 is a 16-bit shift
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ static __forceinline__ uint4 pack(const float (&f)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
};
template <> struct Vec16<__half> {
  static constexpr int N = 8;
  __device__ static __forceinline__ void unpack(const uint4& r, float (&f)[8]) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
  __device__ static __forceinline__ uint4 pack(const float (&f)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
};

// Gather granule of the BACKWARD kernel: VB bytes per lane.  fp32 uses 16 (4 channels / lane, 8 lanes per 128-B row).
// The 16-bit types use 8 bytes (4 channels / lane, 8 lanes per 64-B row) there, so that the 8 lanes of a row write
// 8 x 16 B = 128 CONTIGUOUS bytes of the fp32 grad_value image per `red.v4.f32` instruction.  With 16-byte gathers a
// lane owns 8 channels = two separate 16-B pieces, every red instruction half-fills its 32-B sectors, and the L2 does
// twice the atomic sector operations (measured: bf16 backward 1.8x SLOWER than fp32 -- profiles/).
template <typename T, int VB> struct VecIO;
template <typename T> struct VecIO<T, 16> {
  static constexpr int N = Vec16<T>::N;
  using Raw = uint4;
  __device__ static __forceinline__ Raw load(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ static __forceinline__ void unpack(const Raw& r, float (&f)[N]) { Vec16<T>::unpack(r, f); }
};
template <> struct VecIO<__nv_bfloat16, 8> {
  static constexpr int N = 4;
  using Raw = uint2;
  __device__ static __forceinline__ Raw load(const void* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }
  __device__ static __forceinline__ void unpack(const Raw& r, float (&f)[4]) {
    f[0] = __uint_as_float(r.x << 16); f[1] = __uint_as_float(r.x & 0xffff0000u);
    f[2] = __uint_as_float(r.y << 16); f[3] = __uint_as_float(r.y & 0xffff0000u);
  }
};
template <> struct VecIO<__half, 8> {
  static constexpr int N = 4;
  using Raw = uint2;
  __device__ static __forceinline__ Raw load(const void* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }
  __device__ static __forceinline__ void unpack(const Raw& r, float (&f)[4]) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
  }
};
template <typename T> struct BwdGranule { static constexpr int VB = sizeof(T) == 4 ? 16 : 8; };

// Store N consecutive accumulators as T (N * sizeof(T) in {2, 4, 8, 16} bytes, naturally aligned).
template <typename T, int N>
__device__ __forceinline__ void store_vals(T* p, const float (&r)[N]) {
  if constexpr (N * sizeof(T) == 16) {
    if constexpr (sizeof(T) == 4) {
      *reinterpret_cast<uint4*>(p) = make_uint4(__float_as_uint(r[0]), __float_as_uint(r[1]), __float_as_uint(r[2]), __float_as_uint(r[3]));
    } else {
      float f[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = r[i];
      *reinterpret_cast<uint4*>(p) = Vec16<T>::pack(f);
    }
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) p[i] = from_acc<T>(r[i]);  // ptxas merges these into one 32/64-bit store where it can
  }
}

// Sum r[] over the G = 32/LPR lane groups of a warp (lane = g*LPR + cl) and leave the result SCATTERED: every
// step a lane keeps one half of its values and trades the other half with its partner group (butterfly
// reduce-scatter), so the warp moves VEC-1 values per lane instead of VEC*log2(G) (3 vs 8 shuffles for fp32 D=32,
// 7 vs 24 for bf16 D=32 -- SHFL shares the L1 data pipe with the tap loads, profiles/).  Same pairs are added in the
// same order as a plain xor-butterfly: results are bit-identical.  On return the lane holds `n` sums starting at
// channel `first` of its VEC-channel slice; `owner` is false on lanes holding a duplicate (G > VEC).
template <int VEC, int LPR>
struct GroupReduceScatter {
  static constexpr int G = 32 / LPR;
  static constexpr int STEPS_HALVING = (G >= VEC) ? (VEC == 1 ? 0 : (VEC == 2 ? 1 : (VEC == 4 ? 2 : 3)))
                                                  : (G == 1 ? 0 : (G == 2 ? 1 : (G == 4 ? 2 : (G == 8 ? 3 : 4))));
  static constexpr int N_OUT = VEC >> STEPS_HALVING;
};

template <int N, int O, int LPR>
__device__ __forceinline__ void reduce_scatter_steps(float (&r)[N], int lane, int& first, bool& owner) {
  if constexpr (O < 32) {
    if constexpr (N > 1) {
      constexpr int H = N / 2;
      const bool upper = (lane & O) != 0;
      float k[H];
#pragma unroll
      for (int i = 0; i < H; ++i) {
        const float send = upper ? r[i] : r[H + i];
        const float keep = upper ? r[H + i] : r[i];
        k[i] = keep + __shfl_xor_sync(0xffffffffu, send, O);
      }
      if (upper) first += H;
      reduce_scatter_steps<H, O * 2, LPR>(k, lane, first, owner);
#pragma unroll
      for (int i = 0; i < H; ++i) r[i] = k[i];
    } else {
      r[0] += __shfl_xor_sync(0xffffffffu, r[0], O);
      if (lane & O) owner = false;
      reduce_scatter_steps<1, O * 2, LPR>(r, lane, first, owner);
    }
  }
}

// (x, y) pair of one sampling location
__device__ __forceinline__ void load_xy(const float* p, float& x, float& y) {
  const float2 t = __ldg(reinterpret_cast<const float2*>(p));
  x = t.x; y = t.y;
}
__device__ __forceinline__ void load_xy(const __nv_bfloat16* p, float& x, float& y) {
  const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(p));
  x = __uint_as_float(w << 16); y = __uint_as_float(w & 0xffff0000u);
}
__device__ __forceinline__ void load_xy(const __half* p, float& x, float& y) {
  const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(p));
  const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w));
  x = t.x; y = t.y;
}
__device__ __forceinline__ float load_s(const float* p) { return __ldg(p); }
__device__ __forceinline__ float load_s(const __nv_bfloat16* p) {
  return __uint_as_float(((uint32_t)__ldg(reinterpret_cast<const unsigned short*>(p))) << 16);
}
__device__ __forceinline__ float load_s(const __half* p) {
  const unsigned short h = __ldg(reinterpret_cast<const unsigned short*>(p));
  return __half2float(*reinterpret_cast<const __half*>(&h));
}

// Reference points of the fused operator: stored as T, or as fp32 next to 16-bit value / offsets / logits (`r32`; C-ABI flag
// MSDA_FUSED_REF_F32) -- a bf16 reference point is quantised to 1/256 of the image, i.e. 0.4-0.8 px on a 100-200 px level.
template <typename T>
__device__ __forceinline__ float load_ref(const T* ref, long long i, bool r32) {
  return r32 ? __ldg(reinterpret_cast<const float*>(ref) + i) : load_s(ref + i);
}

__device__ __forceinline__ void store_xy(float* p, float x, float y) { *reinterpret_cast<float2*>(p) = make_float2(x, y); }
__device__ __forceinline__ void store_xy(__nv_bfloat16* p, float x, float y) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(x, y);
}
__device__ __forceinline__ void store_xy(__half* p, float x, float y) { *reinterpret_cast<__half2*>(p) = __floats2half2_rn(x, y); }

// Programmatic dependent launch (sm_90+).  A kernel launched with the "programmatic stream serialization" attribute may
// become resident while its predecessor in the stream is still running; `pdl_wait` blocks until that predecessor has
// completed and its writes are visible (no-op for a normal launch), `pdl_trigger` lets the NEXT kernel in the stream
// start its launch early (only effective once every CTA of this grid has executed it or exited).  Used for ONE pair:
// zero-fill (normal launch, triggers at once) -> backward kernel (gathers during the fill, waits before its first
// scatter).  Chaining every kernel of a step this way (wait at the top, then trigger) was measured and rejected: the
// fill can then only trigger after ITS wait, which serialises the backward launch behind it (C2 backward 13.0 -> 17.0 us),
// and the forward gains nothing (5.35 -> 5.8 us): profiles/r1_sweep_rejected_pdl_chain.jsonl.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }

// vectorised fp32 reduction into global memory (sm_90+): one 16-byte L2 atomic instead of four
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ------------------------------------------------------------------------------------------------
// VECTOR kernels
// ------------------------------------------------------------------------------------------------
// T in {float, bf16, half}; a row of D channels is covered by LPR = D*sizeof(T)/16 lanes (one 128-bit load
// each), so the warp holds G = 32/LPR lane groups that gather G different samples per load instruction.
//
// "Sample geometry": lane i of the warp turns sampling location i of the unit into {tap offset, row stride |
// validity bits, weights} ONCE (one coalesced read of the unit's location / weight block); the lane groups
// then fetch the record of the sample they gather with warp shuffles.  (Letting every lane of a group redo
// that arithmetic cost ~600 warp instructions per unit and made the kernels issue-bound: profiles/.)
//
// Invalid taps (outside the level, or the sample outside the window) read g_zero_line instead of being
// predicated off: unconditional loads keep the 4*U gathers of a lane back to back in the SASS
// (tools/sass_summary.py).  When every tap of the warp's current samples is valid -- the common case -- a
// warp-uniform fast path skips the pointer selects.
//
// MC > 0 fixes the head count at compile time (M*D*sizeof(T) becomes an immediate load offset); MC == 0 reads
// it from the argument.  Offsets are 32-bit: the host checks S*M*D <= 2^27.  Grid: x = units of one image,
// y = image.
__device__ __align__(16) const unsigned int g_zero_line[4] = {0u, 0u, 0u, 0u};

// level of sample s (= s / P) without an integer division: exact for s, P < 2^20
__device__ __forceinline__ int level_of(int s, float inv_p) { return __float2int_rz(((float)s + 0.5f) * inv_p); }

struct SampleGeo {
  int off00;  // element offset of tap (y0, x0) from the image base (head / channel offset NOT included)
  int rsf;    // (W * M * D) << 4 | ok11 << 3 | ok10 << 2 | ok01 << 1 | ok00
};

// What one sample needs before its geometry: location, attention weight, level shape.
struct SampleParams {
  float lx, ly, a;  // normalised location (x, y), attention weight
  int H, W, st;     // level shape and first row
};

// plain operator: locations and (already normalised) weights are inputs.  l32 / a32: the tensor is fp32 although T is a 16-bit
// type (mixed precision, MSDA_LOC_F32 / MSDA_ATTN_F32: torch.autocast leaves the location arithmetic in fp32 -- a bf16 location
// is quantised to 1/256 of the image); the caller then passes unit pointers whose BYTE offset is that of the fp32 tensor
// (unit_ptr below), and they are re-read as float here.
template <typename T>
__device__ __forceinline__ SampleParams load_params(const T* __restrict__ u_loc, const T* __restrict__ u_att,
                                                    const int32_t* __restrict__ shapes, const int32_t* __restrict__ start,
                                                    int s, bool have, float inv_p, bool l32 = false, bool a32 = false) {
  SampleParams p;
  const int si = have ? s : 0;
  const int l = have ? level_of(si, inv_p) : 0;
  if (sizeof(T) != 4 && l32) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(u_loc) + si);
    p.lx = t.x; p.ly = t.y;
  } else {
    load_xy(u_loc + 2 * si, p.lx, p.ly);
  }
  p.a = (sizeof(T) != 4 && a32) ? __ldg(reinterpret_cast<const float*>(u_att) + si) : load_s(u_att + si);
  p.H = __ldg(shapes + 2 * l);
  p.W = __ldg(shapes + 2 * l + 1);
  p.st = __ldg(start + l);
  return p;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Fused operator (the elementwise part of MSDeformAttn.forward, reference ops/modules/ms_deform_attn.py:121-137, done in
// the kernel): inputs are the RAW sampling offsets and attention logits of the two Linear layers plus the reference
// points; lane i produces  A_i = softmax_i(logits over the unit's L*P samples)  and
//   2-d refs:  loc = ref_xy + off / (W_l, H_l)            4-d refs:  loc = ref_xy + off / P * ref_wh * 0.5
// with the same operation order / roundings as the eager PyTorch expression.  Needs L*P <= 32 (one lane per sample).
// `od` returns off / P (4-d) for the reference-point gradient.
template <typename T>
__device__ __forceinline__ SampleParams fused_params(const T* __restrict__ u_off, const T* __restrict__ u_logit,
                                                     const T* __restrict__ ref, long long rq, int RD, bool r32,
                                                     const int32_t* __restrict__ shapes, const int32_t* __restrict__ start,
                                                     int s, bool have, float inv_p, int P, int& l, float& odx, float& ody) {
  SampleParams p;
  const int si = have ? s : 0;
  l = have ? level_of(si, inv_p) : 0;
  float ox, oy;
  load_xy(u_off + 2 * si, ox, oy);
  const float logit = have ? load_s(u_logit + si) : -INFINITY;
  p.H = __ldg(shapes + 2 * l);
  p.W = __ldg(shapes + 2 * l + 1);
  p.st = __ldg(start + l);
  const float rx = load_ref(ref, rq + l * RD, r32), ry = load_ref(ref, rq + l * RD + 1, r32);
  if (RD == 2) {
    odx = 0.f; ody = 0.f;
    p.lx = __fadd_rn(rx, __fdiv_rn(ox, (float)p.W));
    p.ly = __fadd_rn(ry, __fdiv_rn(oy, (float)p.H));
  } else {
    const float rw = load_ref(ref, rq + l * RD + 2, r32), rh = load_ref(ref, rq + l * RD + 3, r32);
    odx = __fdiv_rn(ox, (float)P);
    ody = __fdiv_rn(oy, (float)P);
    p.lx = __fadd_rn(rx, __fmul_rn(__fmul_rn(odx, rw), 0.5f));
    p.ly = __fadd_rn(ry, __fmul_rn(__fmul_rn(ody, rh), 0.5f));
  }
  const float mx = warp_max(logit);
  const float e = have ? expf(logit - mx) : 0.f;
  const float sum = warp_sum(e);
  p.a = __fdiv_rn(e, sum);
  return p;
}

// ---- the same two loaders split into "memory" and "arithmetic" halves, so that a warp that walks several units (the
// patch-ordered forward) can issue the loads of its NEXT unit before the gathers of the current one ----
struct LevelMeta { int H, W, st, l; };
struct RawSample { float x, y, a, rx, ry, rw, rh; };  // plain: (x, y) = location, a = weight; FUSED: offsets, logit, reference point

__device__ __forceinline__ LevelMeta load_level_meta(const int32_t* __restrict__ shapes, const int32_t* __restrict__ start,
                                                     int s, bool have, float inv_p) {
  LevelMeta lm;
  lm.l = have ? level_of(s, inv_p) : 0;
  lm.H = __ldg(shapes + 2 * lm.l);
  lm.W = __ldg(shapes + 2 * lm.l + 1);
  lm.st = __ldg(start + lm.l);
  return lm;
}

template <typename T, bool FUSED>
__device__ __forceinline__ RawSample load_raw(const T* __restrict__ u_loc, const T* __restrict__ u_att,
                                              const T* __restrict__ ref, long long rq, int RD, bool r32, int s, bool have, int l) {
  RawSample r;
  const int si = have ? s : 0;
  load_xy(u_loc + 2 * si, r.x, r.y);
  r.rx = r.ry = r.rw = r.rh = 0.f;
  if constexpr (FUSED) {
    r.a = have ? load_s(u_att + si) : -INFINITY;
    r.rx = load_ref(ref, rq + l * RD, r32);
    r.ry = load_ref(ref, rq + l * RD + 1, r32);
    if (RD != 2) {
      r.rw = load_ref(ref, rq + l * RD + 2, r32);
      r.rh = load_ref(ref, rq + l * RD + 3, r32);
    }
  } else {
    r.a = load_s(u_att + si);
  }
  return r;
}

// same operation order / roundings as load_params / fused_params
template <bool FUSED>
__device__ __forceinline__ SampleParams params_from_raw(const RawSample& r, const LevelMeta& lm, int RD, int P, bool have) {
  SampleParams p;
  p.H = lm.H; p.W = lm.W; p.st = lm.st;
  if constexpr (!FUSED) {
    p.lx = r.x; p.ly = r.y; p.a = r.a;
  } else {
    if (RD == 2) {
      p.lx = __fadd_rn(r.rx, __fdiv_rn(r.x, (float)p.W));
      p.ly = __fadd_rn(r.ry, __fdiv_rn(r.y, (float)p.H));
    } else {
      p.lx = __fadd_rn(r.rx, __fmul_rn(__fmul_rn(__fdiv_rn(r.x, (float)P), r.rw), 0.5f));
      p.ly = __fadd_rn(r.ry, __fmul_rn(__fmul_rn(__fdiv_rn(r.y, (float)P), r.rh), 0.5f));
    }
    const float mx = warp_max(r.a);
    const float e = have ? expf(r.a - mx) : 0.f;
    const float sum = warp_sum(e);
    p.a = __fdiv_rn(e, sum);
  }
  return p;
}

__device__ __forceinline__ void finish_geometry(const SampleParams& p, bool have, int MD, SampleGeo& sg, Geo<float>& ge) {
  ge = make_geo<float>(p.lx, p.ly, p.H, p.W, have);
  sg.off00 = (p.st + ge.row00) * MD;
  sg.rsf = ((p.W * MD) << 4) | (ge.ok11 ? 8 : 0) | (ge.ok10 ? 4 : 0) | (ge.ok01 ? 2 : 0) | (ge.ok00 ? 1 : 0);
}

// the four tap pointers of one sample; `all_ok` is warp-uniform
template <typename T>
__device__ __forceinline__ void tap_pointers(const T* vb, int off, int rsf, int MD, bool all_ok, const T* (&tp)[4]) {
  const T* t0 = vb + off;
  const T* t1 = t0 + (rsf >> 4);
  if (all_ok) {
    tp[0] = t0; tp[1] = t0 + MD; tp[2] = t1; tp[3] = t1 + MD;
  } else {
    const T* zp = reinterpret_cast<const T*>(g_zero_line);
    tp[0] = (rsf & 1) ? t0 : zp;
    tp[1] = (rsf & 2) ? t0 + MD : zp;
    tp[2] = (rsf & 4) ? t1 : zp;
    tp[3] = (rsf & 8) ? t1 + MD : zp;
  }
}

// Which (query, head) unit a warp owns.
//   unit-major (default): consecutive warps = consecutive heads of one query (units are stored in that order);
//   head-major: a CTA holds ONE head of `warps` consecutive queries -- for raster-ordered queries (encoder
//   self-attention) neighbouring queries of the same head sample overlapping pixel neighbourhoods, so their taps
//   hit in the SM's L1 instead of each going to L2.  Pure scheduling: results are identical.
__device__ __forceinline__ bool unit_of_warp(int M, int QM, int head_major, int& uq, int& m) {
  const int wpb = blockDim.x >> 5, w = threadIdx.x >> 5;
  if (head_major) {
    m = blockIdx.x % M;
    const int q = (blockIdx.x / M) * wpb + w;
    uq = q * M + m;
    return uq < QM;
  }
  uq = blockIdx.x * wpb + w;
  m = uq % M;
  return uq < QM;
}

// packed fp32 pairs (sm_100 FFMA2 / FMUL2: two lanes of fp32 math per issue slot)
__device__ __forceinline__ float2 fma2(float w, float2 v, float2 acc) { return __ffma2_rn(make_float2(w, w), v, acc); }

// ------------------------------------------------------------------------------------------------
// VECTOR FORWARD
// ------------------------------------------------------------------------------------------------
// ---- the speculative regular window in two halves: the lane's record, and the gather rounds over a range of records ----
struct SpecRec { unsigned off00, stride; float w00, w01, w10, w11; };

// Geometry of MY sample (see SPEC below) -> record in registers and, SR, in the warp's shared-memory slice.  `extra` is added to
// the byte offset (the head's column offset when one warp serves two heads from a head-0 base pointer: msda_fwd_pair).
template <typename T, bool SR>
__device__ __forceinline__ void spec_record(const SampleParams& sp, bool have, int MD, unsigned extra, SpecRec& r) {
  const int lane = threadIdx.x & 31;
  const float fH = (float)sp.H, fW = (float)sp.W;
  const float y = fma(sp.ly, fH, -0.5f), x = fma(sp.lx, fW, -0.5f);  // same roundings as make_geo
  const bool inside = have && y > -1.f && x > -1.f && y < fH && x < fW;
  const float fy = floorf(y), fx = floorf(x);
  int y0 = (int)fy, x0 = (int)fx;
  const float ly = y - fy, lx = x - fx, hy = 1.f - ly, hx = 1.f - lx;
  float wy0 = hy, wy1 = ly, wx0 = hx, wx1 = lx;
  if (y0 < 0) { y0 = 0; wy0 = ly; wy1 = 0.f; } else if (y0 > sp.H - 2) { y0 = sp.H - 2; wy1 = hy; wy0 = 0.f; }
  if (x0 < 0) { x0 = 0; wx0 = lx; wx1 = 0.f; } else if (x0 > sp.W - 2) { x0 = sp.W - 2; wx1 = hx; wx0 = 0.f; }
  if (!inside) { y0 = 0; x0 = 0; wy0 = 0.f; wy1 = 0.f; wx0 = 0.f; wx1 = 0.f; }  // (NaN / Inf coordinates land here too)
  const float a = inside ? sp.a : 0.f;
  r.w00 = wy0 * wx0 * a; r.w01 = wy0 * wx1 * a; r.w10 = wy1 * wx0 * a; r.w11 = wy1 * wx1 * a;
  // BYTE offsets from the unit's base pointer (32-bit: the host checks S*M*D*sizeof(T) <= 2^29): a tap pointer is then
  // one 64-bit add instead of an index add + scale + carry chain
  r.off00 = (unsigned)((sp.st + y0 * sp.W + x0) * MD) * (unsigned)sizeof(T) + extra;
  r.stride = (unsigned)(sp.W * MD) * (unsigned)sizeof(T);
  if constexpr (SR) {
    extern __shared__ __align__(16) unsigned char msda_dyn_smem[];
    uint4* rec_a = reinterpret_cast<uint4*>(msda_dyn_smem) + (threadIdx.x & ~31);  // this warp's 32 records
    float2* rec_b = reinterpret_cast<float2*>(reinterpret_cast<uint4*>(msda_dyn_smem) + blockDim.x) + (threadIdx.x & ~31);
    __syncwarp();
    rec_a[lane] = make_uint4(r.off00, r.stride, __float_as_uint(r.w00), __float_as_uint(r.w01));
    rec_b[lane] = make_float2(r.w10, r.w11);
    __syncwarp();
  }
}

// Gather rounds over the records of lanes [k_begin, k_end) (warp-uniform; G * U records per round), accumulating into `acc`.
template <typename T, int D, int U, bool SR>
__device__ __forceinline__ void spec_rounds(const T* __restrict__ vb, const SpecRec& r, int MD, int k_begin, int k_end,
                                            float2 (&acc)[Vec16<T>::N / 2]) {
  constexpr int VEC = Vec16<T>::N;
  constexpr int LPR = D / VEC;
  constexpr int G = 32 / LPR;
  const int g = (threadIdx.x & 31) / LPR;
  extern __shared__ __align__(16) unsigned char msda_dyn_smem[];
  const uint4* rec_a = reinterpret_cast<const uint4*>(msda_dyn_smem) + (threadIdx.x & ~31);
  const float2* rec_b = reinterpret_cast<const float2*>(reinterpret_cast<const uint4*>(msda_dyn_smem) + blockDim.x) + (threadIdx.x & ~31);
  const char* vbc = reinterpret_cast<const char*>(vb);
  const size_t mdb = (size_t)MD * sizeof(T);
  auto round = [&](int k0) {
    unsigned off[U], rs[U];
    float w[U][4];
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const int src = k0 + j * G + g;  // < 32
      if constexpr (SR) {
        const uint4 ra = rec_a[src];
        const float2 rb = rec_b[src];
        off[j] = ra.x; rs[j] = ra.y;
        w[j][0] = __uint_as_float(ra.z); w[j][1] = __uint_as_float(ra.w); w[j][2] = rb.x; w[j][3] = rb.y;
      } else {
        off[j] = __shfl_sync(0xffffffffu, r.off00, src);
        rs[j] = __shfl_sync(0xffffffffu, r.stride, src);
        w[j][0] = __shfl_sync(0xffffffffu, r.w00, src);
        w[j][1] = __shfl_sync(0xffffffffu, r.w01, src);
        w[j][2] = __shfl_sync(0xffffffffu, r.w10, src);
        w[j][3] = __shfl_sync(0xffffffffu, r.w11, src);
      }
    }
    uint4 v[U][4];
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const char* t0 = vbc + off[j];
      const char* t1 = t0 + rs[j];
      v[j][0] = ldg128(t0); v[j][1] = ldg128(t0 + mdb); v[j][2] = ldg128(t1); v[j][3] = ldg128(t1 + mdb);
    }
#pragma unroll
    for (int j = 0; j < U; ++j)
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float f[VEC];
        Vec16<T>::unpack(v[j][t], f);
#pragma unroll
        for (int i = 0; i < VEC / 2; ++i) acc[i] = fma2(w[j][t], make_float2(f[2 * i], f[2 * i + 1]), acc[i]);
      }
  };
  // (a fully unrolled 4-round variant for L*P = 16 was measured 5 % SLOWER -- 100.1 vs 94.8 us on the encoder shape: the
  // kernel grows by 70 instructions and ptxas needs 2 spill slots at 40 registers)
  for (int k0 = k_begin; k0 < k_end; k0 += G * U) round(k0);  // warp-uniform trip count
}

// Dependent-latency chain per warp: {loc, attn, level shapes} -> 4*U tap rows per group -> shuffles -> store.
// Register budgets via min-CTAs/SM at 256 threads: U=1 -> 40 regs (48 warps/SM), U=2 -> 64 regs, U=4 -> 80 regs.
// FUSED: `loc` holds the raw sampling offsets, `attn` the raw attention logits, `ref` the reference points (last dim RD).
// SR: the per-sample records travel from the geometry lanes to the lane groups through a warp-private slice of
// shared memory (2 broadcast LDS per round, one wavefront each) instead of 6 SHFL per round (one wavefront each on
// the same L1 data pipe the tap rows return through).  Dynamic shared memory: 24 bytes per thread.
// Gather + reduce + store of one pass (<= 32 samples, one per lane) of one unit, given the lane's sample parameters.
// `acc` carries the partial sums across passes.
// SPEC ("speculative regular window"): the sample's lane moves the 2x2 tap window fully inside the level (needs H, W >= 2) and
// gives the taps that fell outside a ZERO WEIGHT instead of a zero address: x0 = -1 -> window columns (0, 1) with weights
// (lx, 0); x0 = W-1 -> columns (W-2, W-1) with weights (0, hx); same for rows; a sample outside the (-1, size) window gets
// four zero weights on the level's first pixels.  Every tap address is then valid and regular (base, +MD, +stride,
// +stride+MD), so the rounds need no validity flags, no vote and no pointer selects (the flagged path costs ~30 extra
// instructions on every round that holds ONE invalid tap: 48 % of the rounds of an encoder call with +-4 pixel offsets,
// 83 % of its 13x13-level rounds).  The effective (non-zero-weight) FMAs are the same, in the same order, with the same
// weights: results are bit-identical as long as every loaded value is finite; 0 * Inf / 0 * NaN from a clamped tap would
// differ, so the caller checks the unit's sums and redoes a non-finite unit on the flagged path (msda_fwd_unit).
template <typename T, int D, int U, bool SR, bool SPEC = false>
__device__ __forceinline__ void
msda_fwd_gather_pass(const T* __restrict__ vb, const SampleParams& sp, bool have, int MD, int cnt, float2 (&acc)[Vec16<T>::N / 2]) {
  constexpr int VEC = Vec16<T>::N;
  constexpr int LPR = D / VEC;
  constexpr int G = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int g = lane / LPR;
  extern __shared__ __align__(16) unsigned char msda_dyn_smem[];
  uint4* rec_a = reinterpret_cast<uint4*>(msda_dyn_smem) + (threadIdx.x & ~31);  // this warp's 32 records
  float2* rec_b = reinterpret_cast<float2*>(reinterpret_cast<uint4*>(msda_dyn_smem) + blockDim.x) + (threadIdx.x & ~31);

  if constexpr (SPEC) {
    SpecRec r;
    spec_record<T, SR>(sp, have, MD, 0u, r);
    spec_rounds<T, D, U, SR>(vb, r, MD, 0, cnt, acc);  // lanes >= cnt hold zero-weight records
    return;
  }

  SampleGeo sg;
  Geo<float> ge;
  finish_geometry(sp, have, MD, sg, ge);
  const float a = ge.inside ? sp.a : 0.f;  // an outside sample contributes exactly nothing, whatever its weight is
  const float w00 = ge.hy * ge.hx * a, w01 = ge.hy * ge.lx * a, w10 = ge.ly * ge.hx * a, w11 = ge.ly * ge.lx * a;
  if constexpr (SR) {
    __syncwarp();  // the previous pass / the previous unit of this warp has finished reading the records
    rec_a[lane] = make_uint4((unsigned)sg.off00, (unsigned)sg.rsf, __float_as_uint(w00), __float_as_uint(w01));
    rec_b[lane] = make_float2(w10, w11);
    __syncwarp();
  }
  for (int k0 = 0; k0 < cnt; k0 += G * U) {  // warp-uniform trip count
    int off[U], rsf[U];
    float w[U][4];
    bool full = true;
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const int src = k0 + j * G + g;  // < 32: G * U divides 32
      if constexpr (SR) {
        const uint4 ra = rec_a[src];
        const float2 rb = rec_b[src];
        off[j] = (int)ra.x;
        rsf[j] = (int)ra.y;
        w[j][0] = __uint_as_float(ra.z);
        w[j][1] = __uint_as_float(ra.w);
        w[j][2] = rb.x;
        w[j][3] = rb.y;
      } else {
        off[j] = __shfl_sync(0xffffffffu, sg.off00, src);
        rsf[j] = __shfl_sync(0xffffffffu, sg.rsf, src);
        w[j][0] = __shfl_sync(0xffffffffu, w00, src);
        w[j][1] = __shfl_sync(0xffffffffu, w01, src);
        w[j][2] = __shfl_sync(0xffffffffu, w10, src);
        w[j][3] = __shfl_sync(0xffffffffu, w11, src);
      }
      if (src >= cnt) rsf[j] = 0;  // a lane group past the end must not gather
      full = full && ((rsf[j] & 15) == 15);
    }
    const bool all_ok = __all_sync(0xffffffffu, full);
    uint4 v[U][4];
    if (all_ok) {  // loads duplicated per branch: here the x+1 taps are immediate offsets of the x taps
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const T* t0 = vb + off[j];
        const T* t1 = vb + (off[j] + (rsf[j] >> 4));  // one IMAD.WIDE per pointer
        v[j][0] = ldg128(t0); v[j][1] = ldg128(t0 + MD); v[j][2] = ldg128(t1); v[j][3] = ldg128(t1 + MD);
      }
    } else {
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const T* tp[4];
        tap_pointers<T>(vb, off[j], rsf[j], MD, false, tp);
#pragma unroll
        for (int t = 0; t < 4; ++t) v[j][t] = ldg128(tp[t]);
      }
    }
#pragma unroll
    for (int j = 0; j < U; ++j)
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float f[VEC];
        Vec16<T>::unpack(v[j][t], f);
#pragma unroll
        for (int i = 0; i < VEC / 2; ++i) acc[i] = fma2(w[j][t], make_float2(f[2 * i], f[2 * i + 1]), acc[i]);
      }
  }
}

// sum the lane groups' partial sums (scattered over the warp: GroupReduceScatter) ...
template <typename T, int D>
__device__ __forceinline__ void msda_fwd_reduce(const float2 (&acc)[Vec16<T>::N / 2], float (&o)[GroupReduceScatter<Vec16<T>::N, D / Vec16<T>::N>::N_OUT],
                                                int& first, bool& owner) {
  constexpr int VEC = Vec16<T>::N;
  constexpr int LPR = D / VEC;
  const int lane = threadIdx.x & 31;
  float r[VEC];
#pragma unroll
  for (int i = 0; i < VEC / 2; ++i) { r[2 * i] = acc[i].x; r[2 * i + 1] = acc[i].y; }
  first = 0;
  owner = true;
  reduce_scatter_steps<VEC, LPR, LPR>(r, lane, first, owner);
  constexpr int N_OUT = GroupReduceScatter<VEC, LPR>::N_OUT;
#pragma unroll
  for (int i = 0; i < N_OUT; ++i) o[i] = r[i];
}

// ... and write the unit's D outputs
template <typename T, int D>
__device__ __forceinline__ void msda_fwd_store(T* __restrict__ out_u, const float2 (&acc)[Vec16<T>::N / 2]) {
  constexpr int VEC = Vec16<T>::N;
  constexpr int LPR = D / VEC;
  constexpr int N_OUT = GroupReduceScatter<VEC, LPR>::N_OUT;
  const int cl = (threadIdx.x & 31) % LPR;
  float o[N_OUT];
  int first;
  bool owner;
  msda_fwd_reduce<T, D>(acc, o, first, owner);
  if (owner) store_vals<T, N_OUT>(out_u + cl * VEC + first, o);
}

// One unit (image b, in-image unit index uq = q * M + m) by the calling warp.
template <typename T, int D, int MC, int U, bool FUSED, bool SR>
__device__ __forceinline__ void
msda_fwd_unit(const T* __restrict__ value, const int32_t* __restrict__ shapes,
              const int32_t* __restrict__ start, const T* __restrict__ loc,
              const T* __restrict__ attn, T* __restrict__ out,
              int S, int M, int L, int P, float inv_p, int QM, const T* __restrict__ ref, int RDf, int b, int uq, int m,
              int spec_on = 0) {
  const int RD = RDf & 0xff;            // last dim of the reference points (2 or 4)
  const bool r32 = (RDf & 0x100) != 0;  // ... which are fp32 although T is a 16-bit type
  constexpr int VEC = Vec16<T>::N;
  constexpr int LPR = D / VEC;
  static_assert(D % VEC == 0 && LPR >= 1 && LPR <= 32 && (LPR & (LPR - 1)) == 0, "unsupported D for the vector path");

  const int MD = M * D;
  const int lane = threadIdx.x & 31;
  const int cl = lane % LPR;
  const int LP = L * P;
  const long long u = (long long)b * QM + uq;
  const bool l32 = sizeof(T) != 4 && !FUSED && (spec_on & 2) != 0;  // bits 1 / 2 of the word: fp32 locations / weights next to
  const bool a32 = sizeof(T) != 4 && !FUSED && (spec_on & 4) != 0;  // 16-bit value (load_params); element offsets double
  const T* __restrict__ u_loc = loc + u * (LP * 2) * (l32 ? 2 : 1);
  const T* __restrict__ u_att = attn + u * LP * (a32 ? 2 : 1);
  const T* __restrict__ vb = value + (long long)b * S * MD + (m * D + cl * VEC);

  constexpr int N_OUT = GroupReduceScatter<VEC, LPR>::N_OUT;
  float2 acc[VEC / 2];
  float o[N_OUT];
  int first;
  bool owner;
  bool spec = (spec_on & 1) != 0;  // try the speculative regular-window gather first (see msda_fwd_gather_pass)
  for (;;) {
#pragma unroll
    for (int i = 0; i < VEC / 2; ++i) acc[i] = make_float2(0.f, 0.f);
    bool restart = false;
    for (int base = 0; base < LP; base += 32) {  // 32 samples per pass, one per lane (FUSED: L*P <= 32, one pass)
      SampleParams sp;
      const bool have = base + lane < LP;
      if constexpr (FUSED) {
        int l;
        float odx, ody;
        sp = fused_params<T>(u_loc, u_att, ref, ((long long)b * (QM / M) + uq / M) * (L * RD), RD, r32, shapes, start, lane, have, inv_p, P, l, odx, ody);
      } else {
        sp = load_params<T>(u_loc, u_att, shapes, start, base + lane, have, inv_p, l32, a32);
      }
      // the regular window needs levels of at least 2 x 2 pixels (warp-uniform; the decision sticks for the unit)
      if (spec && !__all_sync(0xffffffffu, !have || (sp.H >= 2 && sp.W >= 2))) {
        spec = false;
        if (base > 0) { restart = true; break; }  // earlier passes were speculative: redo the unit on the flagged path
      }
      if (spec) msda_fwd_gather_pass<T, D, U, SR, true>(vb, sp, have, MD, min(32, LP - base), acc);
      else msda_fwd_gather_pass<T, D, U, SR, false>(vb, sp, have, MD, min(32, LP - base), acc);
    }
    if (restart) continue;
    msda_fwd_reduce<T, D>(acc, o, first, owner);
    if (spec) {
      bool bad = false;
#pragma unroll
      for (int i = 0; i < N_OUT; ++i) bad = bad || !(fabsf(o[i]) <= 3.402823466e38f);
      if (__any_sync(0xffffffffu, bad)) {  // a non-finite value met a zero weight (or is simply there): the flagged path decides
        spec = false;
        continue;
      }
    }
    break;
  }
  if (owner) store_vals<T, N_OUT>(out + u * D + cl * VEC + first, o);
}

// The flagged (exact zero-line) path of one unit as an out-of-line call: the persistent kernels take it only for units whose
// speculative sums came out non-finite or when a level is narrower than 2 pixels, and must not pay its registers.
template <typename T, int D, int MC, bool FUSED, bool SR>
__device__ __noinline__ void
msda_fwd_unit_flagged(const T* __restrict__ value, const int32_t* __restrict__ shapes, const int32_t* __restrict__ start,
                      const T* __restrict__ loc, const T* __restrict__ attn, T* __restrict__ out,
                      int S, int M, int L, int P, float inv_p, int QM, const T* __restrict__ ref, int RDf, int b, int uq, int m,
                      int io = 0) {  // io: bits 1 / 2 of msda_fwd_unit's spec_on word (fp32 locations / weights next to 16-bit value)
  msda_fwd_unit<T, D, MC, 1, FUSED, SR>(value, shapes, start, loc, attn, out, S, M, L, P, inv_p, QM, ref, RDf, b, uq, m, io & 6);
}

// grid: x = units of one image (one warp each), y = image
template <typename T, int D, int MC, int U, bool FUSED, bool SR>
__global__ void __launch_bounds__(MSDA_MAX_THREADS, U == 1 ? MSDA_FWD_MIN_CTAS : (U == 2 ? 4 : 3))
msda_fwd_sg_kernel(const T* __restrict__ value, const int32_t* __restrict__ shapes,
                   const int32_t* __restrict__ start, const T* __restrict__ loc,
                   const T* __restrict__ attn, T* __restrict__ out,
                   int S, int Mrt, int L, int P, float inv_p, int QM, const T* __restrict__ ref, int RDf, int head_major, int spec_on) {
  const int M = MC > 0 ? MC : Mrt;
  int uq, m;  // unit inside image blockIdx.y, its head
  if (!unit_of_warp(M, QM, head_major, uq, m)) return;  // warp-uniform
  msda_fwd_unit<T, D, MC, U, FUSED, SR>(value, shapes, start, loc, attn, out, S, M, L, P, inv_p, QM, ref, RDf, (int)blockIdx.y, uq, m, spec_on);
}

// ------------------------------------------------------------------------------------------------
// PAIRED forward: one warp serves TWO units -- heads m0 and m0 + 1 of one query -- when L * P <= 16.
// With 16 samples per unit half of the warp idles through the per-sample prologue (parameter loads, window geometry, record
// store: ~190 of the ~315 warp instructions a unit costs on the encoder shape, the gather rounds are only ~92).  Here lanes
// 0..15 hold the samples of head m0 and lanes 16..31 those of head m0 + 1 (their locations / weights are one contiguous
// 32-sample run in memory), the prologue runs once for both, and the rounds walk records [0, LP) and then [16, 16 + LP).
// The two heads read different columns of `value`: the records carry the head's column offset, the base pointer is head 0's.
// Arithmetic per unit is the speculative path of msda_fwd_unit, instruction for instruction: results are bit-identical; a unit
// whose sums come out non-finite, or a level narrower than 2 pixels, goes to the flagged path (msda_fwd_unit_flagged).
// ------------------------------------------------------------------------------------------------
template <typename T, int D, int MC, bool SR>
__device__ __forceinline__ void
msda_fwd_pair(const T* __restrict__ value, const int32_t* __restrict__ shapes, const int32_t* __restrict__ start,
              const T* __restrict__ loc, const T* __restrict__ attn, T* __restrict__ out,
              int S, int M, int L, int P, float inv_p, int QM, int b, int q, int m0, int io) {
  constexpr int VEC = Vec16<T>::N;
  constexpr int LPR = D / VEC;
  constexpr int N_OUT = GroupReduceScatter<VEC, LPR>::N_OUT;
  const int MD = M * D;
  const int lane = threadIdx.x & 31;
  const int cl = lane % LPR;
  const int LP = L * P;
  const int h = lane >> 4, s = lane & 15;       // which of the two heads, which of its samples
  const int nh = min(2, M - m0);                // (an odd head count leaves the last pair with one head)
  const bool have = s < LP && h < nh;
  const bool l32 = sizeof(T) != 4 && (io & 2) != 0;
  const bool a32 = sizeof(T) != 4 && (io & 4) != 0;
  const int uq = q * M + m0;
  const long long u0 = (long long)b * QM + uq;  // unit of head m0; head m0 + 1 is unit u0 + 1
  const long long su = (u0 + (have ? h : 0)) * LP;  // first sample of MY unit
  const T* __restrict__ u_loc = loc + su * 2 * (l32 ? 2 : 1);
  const T* __restrict__ u_att = attn + su * (a32 ? 2 : 1);
  const SampleParams sp = load_params<T>(u_loc, u_att, shapes, start, s, have, inv_p, l32, a32);
  if (!__all_sync(0xffffffffu, !have || (sp.H >= 2 && sp.W >= 2))) {  // no regular 2 x 2 window on some level (warp-uniform)
    for (int k = 0; k < nh; ++k)
      msda_fwd_unit_flagged<T, D, MC, false, SR>(value, shapes, start, loc, attn, out, S, M, L, P, inv_p, QM, nullptr, 2, b, uq + k, m0 + k, io);
    return;
  }
  SpecRec r;
  spec_record<T, SR>(sp, have, MD, (unsigned)((m0 + h) * D) * (unsigned)sizeof(T), r);
  const T* __restrict__ vb = value + (long long)b * S * MD + cl * VEC;  // head 0: the records hold the head's column
#pragma unroll 1
  for (int k = 0; k < nh; ++k) {
    float2 acc[VEC / 2];
#pragma unroll
    for (int i = 0; i < VEC / 2; ++i) acc[i] = make_float2(0.f, 0.f);
    spec_rounds<T, D, 1, SR>(vb, r, MD, 16 * k, 16 * k + LP, acc);
    float o[N_OUT];
    int first;
    bool owner;
    msda_fwd_reduce<T, D>(acc, o, first, owner);
    bool bad = false;
#pragma unroll
    for (int i = 0; i < N_OUT; ++i) bad = bad || !(fabsf(o[i]) <= 3.402823466e38f);
    if (__any_sync(0xffffffffu, bad)) {  // a non-finite value met a zero weight (or is simply there): the flagged path decides
      msda_fwd_unit_flagged<T, D, MC, false, SR>(value, shapes, start, loc, attn, out, S, M, L, P, inv_p, QM, nullptr, 2, b, uq + k, m0 + k, io);
      if constexpr (SR) __syncwarp();
    } else if (owner) {
      store_vals<T, N_OUT>(out + (u0 + k) * D + cl * VEC + first, o);
    }
  }
}

// Scheduling words of the SM-affine variant (device memory, zero before the launch): [0] = next chunk to hand out,
// [1 + smid] = the chunk SM `smid` is working on ((chunk + 1) << 32 | next position in it).
#define MSDA_SCHED_WORDS 1025
#define MSDA_SCHED_DONE 0xffffffffu

// grid (static):  x = query-pairs of one image (one warp each), y = image.
// grid (AFFINE):  persistent, x = SMs * CTAs per SM.  For pixel-aligned queries (encoder self-attention: Lq == S, query i is pixel
//   i of the pyramid and samples around its own position) the warps of ONE SM work through PY x PX patches of queries, all head
//   pairs of a patch in turn, so the taps of neighbouring queries are served by that SM's L1 instead of each going to L2 (the
//   unit-ordered kernel: 21 % L1 hits, every miss costs a fill cycle on the L1 data stage it shares with the reads).  A chunk = one
//   patch of one image = PY*PX*ceil(M/2) warp items; the SM's warps take items of its current chunk with one atomicAdd each
//   (issued before the previous item's gathers, so its latency is hidden) and the warp that draws the last one fetches the next
//   chunk from the global counter: SMs stay on their patch whatever the warps' individual speeds are, the chunks balance the SMs
//   dynamically.  Pure scheduling: every (image, query, head) is processed exactly once whatever the level shapes are (queries
//   beyond the level grids come as plain-order chunks); if the queries are not pixel-aligned only locality is lost.
template <typename T, int D, int MC, bool SR, bool AFFINE>
__global__ void __launch_bounds__(256, AFFINE ? 5 : 6)
msda_fwd_pair_kernel(const T* __restrict__ value, const int32_t* __restrict__ shapes, const int32_t* __restrict__ start,
                     const T* __restrict__ loc, const T* __restrict__ attn, T* __restrict__ out,
                     int N, int S, int Mrt, int L, int P, float inv_p, int QM, int io, unsigned long long* __restrict__ sched,
                     int pxs, int pys) {
  const int M = MC > 0 ? MC : Mrt;
  const int HP = (M + 1) >> 1;  // head pairs per query
  const int Lq = QM / M;
  if constexpr (!AFFINE) {
    const long long item = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (item >= (long long)Lq * HP) return;  // warp-uniform
    const int q = (int)(item / HP), hp = (int)(item - (long long)q * HP);
    msda_fwd_pair<T, D, MC, SR>(value, shapes, start, loc, attn, out, S, M, L, P, inv_p, QM, (int)blockIdx.y, q, 2 * hp, io);
  } else {
    const int lane = threadIdx.x & 31;
    const int PX = 1 << pxs, PY = 1 << pys;
    const unsigned C = (unsigned)(HP << (pxs + pys));  // items per chunk
    // patches per image over the level grids, queries they cover (warp-uniform, L <= a few levels)
    int patches = 0, pixels = 0;
    for (int l = 0; l < L; ++l) {
      const int H = __ldg(shapes + 2 * l), W = __ldg(shapes + 2 * l + 1);
      patches += ((H + PY - 1) >> pys) * ((W + PX - 1) >> pxs);
      pixels += H * W;
    }
    pixels = min(pixels, Lq);
    const long long grid_chunks = (long long)patches * N;
    const int tail_q = Lq - pixels;  // queries beyond the level grids: chunks of PY*PX queries in plain order
    const int tail_per_image = (tail_q + (1 << (pxs + pys)) - 1) >> (pxs + pys);
    const long long n_chunks = grid_chunks + (long long)tail_per_image * N;
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    unsigned long long* cur = sched + 1 + (smid & 1023u);
    // one ticket of this SM's current chunk: lane 0 draws, the value is broadcast only where it is needed (`bcast`), so the
    // draw for the NEXT item is in flight while this one is processed
    auto draw = [&]() -> unsigned long long { return lane == 0 ? atomicAdd(cur, 1ull) : 0ull; };
    auto bcast = [&](unsigned long long v) -> unsigned long long { return __shfl_sync(0xffffffffu, v, 0); };
    unsigned long long t_raw = draw();
    for (;;) {
      unsigned long long t = bcast(t_raw);
      unsigned hi = (unsigned)(t >> 32), pos = (unsigned)t;
      if (hi == MSDA_SCHED_DONE) break;
      if (hi == 0 || pos >= C) {  // no chunk installed yet / the chunk is used up
        if (hi == 0 ? pos != 0 : pos != C) {  // ... and another warp (the one that drew ticket 0 / C) is fetching the next one
          __nanosleep(200);
          t_raw = draw();
          continue;
        }
        unsigned long long nc = 0;  // I drew the first ticket past the end (or the very first one): fetch the next chunk
        if (lane == 0) {
          nc = atomicAdd(sched, 1ull);
          atomicExch(cur, nc >= (unsigned long long)n_chunks ? ((unsigned long long)MSDA_SCHED_DONE << 32) : (((nc + 1) << 32) | 1ull));
        }
        nc = __shfl_sync(0xffffffffu, nc, 0);
        if (nc >= (unsigned long long)n_chunks) break;
        hi = (unsigned)nc + 1;
        pos = 0;
      }
      const unsigned long long t_next = draw();  // in flight while this item is processed
      const long long c = (long long)hi - 1;
      const int hp = (int)(pos >> (pxs + pys)), rr = (int)(pos & ((1u << (pxs + pys)) - 1));
      int b, q = -1;
      if (c < grid_chunks) {
        b = (int)(c / patches);
        int patch = (int)(c - (long long)b * patches);
        int l = 0, H = 0, W = 0, npx = 1, first = 0;  // first = index of the level's first query
        for (; l < L; ++l) {
          H = __ldg(shapes + 2 * l);
          W = __ldg(shapes + 2 * l + 1);
          npx = (W + PX - 1) >> pxs;
          const int np = ((H + PY - 1) >> pys) * npx;
          if (patch < np) break;
          patch -= np;
          first += H * W;
        }
        const int prow = patch / npx;
        const int y = (prow << pys) + (rr >> pxs), x = ((patch - prow * npx) << pxs) + (rr & (PX - 1));
        if (y < H && x < W) q = first + y * W + x;
        if (q >= Lq) q = -1;
      } else {
        const long long ct = c - grid_chunks;
        b = (int)(ct / tail_per_image);
        q = pixels + (((int)(ct - (long long)b * tail_per_image)) << (pxs + pys)) + rr;
        if (q >= Lq) q = -1;
      }
      if (q >= 0) msda_fwd_pair<T, D, MC, SR>(value, shapes, start, loc, attn, out, S, M, L, P, inv_p, QM, b, q, 2 * hp, io);
      t_raw = t_next;
    }
  }
}

// PATCH-ORDERED forward for pixel-aligned queries (encoder self-attention: Lq == S, query i is pixel i of the pyramid
// and samples around its own position).  A CTA owns one head of a PY x PX patch of queries of one level -- warp w walks
// the PX queries of patch row w -- so the taps of neighbouring queries are served by this SM's L1 instead of each
// going to L2 (unit-ordered kernel: 22 % L1 hit rate on the encoder shape, long-scoreboard stalls 63 %: profiles/).
// Persistent CTAs pick work items (level patch, image, head) round-robin; levels are enumerated finest first, so
// the big items come first and the static schedule ends balanced.  Pure scheduling: the arithmetic of a unit is
// msda_fwd_unit, identical to the unit-ordered kernel, and every query in [0, Lq) is processed exactly once whatever
// the level shapes are; if the queries are NOT pixel-aligned only locality is lost.  blockDim.x = 32 * PY.
#ifndef MSDA_PATCH_MAX_THREADS
#define MSDA_PATCH_MAX_THREADS 512
#endif
template <typename T, int D, int MC, bool FUSED, bool SR, int MINB = 2>
__global__ void __launch_bounds__(MSDA_PATCH_MAX_THREADS, MINB)
msda_fwd_patch_kernel(const T* __restrict__ value, const int32_t* __restrict__ shapes,
                      const int32_t* __restrict__ start, const T* __restrict__ loc,
                      const T* __restrict__ attn, T* __restrict__ out,
                      int N, int S, int Mrt, int L, int P, float inv_p, int QM, const T* __restrict__ ref, int RDf, int PX, int spec_on) {
  constexpr int VEC = Vec16<T>::N;
  constexpr int LPR = D / VEC;
  const int RD = RDf & 0xff;
  const bool r32 = (RDf & 0x100) != 0;
  const int M = MC > 0 ? MC : Mrt;
  const int MD = M * D;
  const int PY = blockDim.x >> 5;
  const int w = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cl = lane % LPR;
  const int NM = N * M;
  const int Lq = QM / M;
  const int LP = L * P;
  const bool have = lane < LP;
  const LevelMeta lm = load_level_meta(shapes, start, lane, have, inv_p);  // of MY sample: the same for every unit
  const bool spec_ok = spec_on != 0 && __all_sync(0xffffffffu, !have || (lm.H >= 2 && lm.W >= 2));
  int total = 0, pixels = 0;  // patches / queries covered by the level grids
  for (int l = 0; l < L; ++l) {
    const int H = __ldg(shapes + 2 * l), W = __ldg(shapes + 2 * l + 1);
    total += ((H + PY - 1) / PY) * ((W + PX - 1) / PX);
    pixels += H * W;
  }
  const long long items = (long long)total * NM;
  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
    const int bm = (int)(item % NM);
    int patch = (int)(item / NM);
    int l = 0, H = 0, W = 0, npx = 1, first = 0;  // first = index of the level's first query (queries are NOT tied to
    for (; l < L; ++l) {                          // level_start_index, which describes `value`)
      H = __ldg(shapes + 2 * l);
      W = __ldg(shapes + 2 * l + 1);
      npx = (W + PX - 1) / PX;
      const int np = ((H + PY - 1) / PY) * npx;
      if (patch < np) break;
      patch -= np;
      first += H * W;
    }
    const int y = (patch / npx) * PY + w;
    const int x0 = (patch % npx) * PX;
    if (y >= H) continue;  // warp-uniform: this patch row lies below the level
    const int b = bm / M, m = bm % M;
    const int q0 = first + y * W + x0;
    const int x1 = min(PX, W - x0);
    const int n_q = min(x1, Lq - q0);  // <= 0 when the level grids overshoot Lq
    if (n_q <= 0) continue;
    // walk the row: the loads of unit i + 1 are issued before the gathers of unit i (needs L * P <= 32: host-checked)
    const T* __restrict__ vb = value + (long long)b * S * MD + (m * D + cl * VEC);
    long long u = (long long)b * QM + (long long)q0 * M + m;
    RawSample raw = load_raw<T, FUSED>(loc + u * (LP * 2), attn + u * LP, ref, ((long long)b * Lq + q0) * (L * RD), RD, r32, lane, have, lm.l);
    for (int i = 0; i < n_q; ++i) {
      const RawSample cur = raw;
      if (i + 1 < n_q) {
        const long long un = u + M;
        raw = load_raw<T, FUSED>(loc + un * (LP * 2), attn + un * LP, ref, ((long long)b * Lq + q0 + i + 1) * (L * RD), RD, r32, lane, have, lm.l);
      }
      const SampleParams sp = params_from_raw<FUSED>(cur, lm, RD, P, have);
      constexpr int N_OUT = GroupReduceScatter<VEC, LPR>::N_OUT;
      float2 acc[VEC / 2];
      float o[N_OUT];
      int first;
      bool owner;
      bool done = false;
      if (spec_ok) {  // speculative regular-window gather (msda_fwd_gather_pass); a non-finite sum sends the unit to the flagged path
#pragma unroll
        for (int k = 0; k < VEC / 2; ++k) acc[k] = make_float2(0.f, 0.f);
        msda_fwd_gather_pass<T, D, 1, SR, true>(vb, sp, have, MD, LP, acc);
        msda_fwd_reduce<T, D>(acc, o, first, owner);
        bool bad = false;
#pragma unroll
        for (int k = 0; k < N_OUT; ++k) bad = bad || !(fabsf(o[k]) <= 3.402823466e38f);
        done = !__any_sync(0xffffffffu, bad);
        if (done && owner) store_vals<T, N_OUT>(out + u * D + cl * VEC + first, o);
      }
      if (!done)
        msda_fwd_unit_flagged<T, D, MC, FUSED, SR>(value, shapes, start, loc, attn, out, S, M, L, P, inv_p, QM, ref, RDf, b,
                                                   (int)(u - (long long)b * QM), m);
      u += M;
    }
  }
  // queries beyond the level grids (Lq > sum H*W: not pixel-aligned after all) in plain unit order
  const long long tail = (long long)max(0, Lq - pixels) * NM;
  for (long long t = (long long)blockIdx.x * PY + w; t < tail; t += (long long)gridDim.x * PY) {
    const int bm = (int)(t % NM);
    const int q = pixels + (int)(t / NM);
    msda_fwd_unit<T, D, MC, 1, FUSED, SR>(value, shapes, start, loc, attn, out, S, M, L, P, inv_p, QM, ref, RDf, bm / M, q * M + bm % M, bm % M, spec_on);
  }
}

// ------------------------------------------------------------------------------------------------
// STAGED FORWARD: the coarsest pyramid levels of one (image, head) live in shared memory, brought there by TMA
// ------------------------------------------------------------------------------------------------
// For calls with many queries per image (encoder self-attention: Lq = S) every (image, head) slab of the coarse levels
// is gathered from thousands of times: with 4 levels and 4 points, HALF of all taps fall into the two coarsest
// levels, which hold 6 % of the rows (794 rows x 128 B = 101 KB per head for the 100^2..13^2 pyramid, 169 KB for
// 800x1333).  A persistent CTA (one per SM) owns a contiguous range of work items (image b, head m, 32 consecutive
// queries); whenever (b, m) changes it stages the trailing levels that fit its tile -- rows [start[L0], S) of head m,
// D*sizeof(T) bytes each, pitch M*D*sizeof(T) -- with one `cp.async.bulk` (TMA, 1-D bulk copy) per row, all
// completing on one mbarrier.  Taps of staged levels are then LDS.128 from the tile (29-cycle latency, no L2 round
// trip), the fine levels keep the LDG path.  Which levels are staged is decided IN the kernel (the level tensors are
// device memory): the largest suffix of levels that is contiguous in `value` and fits `tile_rows`; nothing fits or
// the levels are not laid out back to back -> L0 = L and the kernel degenerates to the plain gather.  Pure
// scheduling/data placement: the arithmetic of a unit is the same as in msda_fwd_sg_kernel (bit-identical results).
// Dynamic shared memory: [records 24 B/thread][tile_rows rows][one zero row][mbarrier].
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t a = smem_u32(bar);
  while (!done)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
}
// TMA 1-D bulk copy global -> shared (bytes: multiple of 16; both addresses 16-byte aligned), completes on `bar`
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint4 lds128(const void* p) { return *reinterpret_cast<const uint4*>(p); }

// TPB / MINB: launch bounds (1024 x 1: 64 registers, 32 warps per SM).  LPC: compile-time L*P (0 = run time) -- with
// LPC = 16 the sample rounds are fully unrolled and the tail checks disappear.  Needs L*P <= 32 (one sample per lane:
// the lane's level is fixed, so its metadata, tile / image offsets and row stride are loop invariants, the pointers
// of a warp's consecutive units advance by constants, and the raw location / weight of the NEXT unit is fetched
// before the gathers of the current one).
__device__ __forceinline__ uint4 lds128_u32(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}

template <typename T, int D, int MC, bool FUSED, int TPB, int MINB, int LPC>
__global__ void __launch_bounds__(TPB, MINB)
msda_fwd_staged_kernel(const T* __restrict__ value, const int32_t* __restrict__ shapes,
                       const int32_t* __restrict__ start, const T* __restrict__ loc,
                       const T* __restrict__ attn, T* __restrict__ out,
                       int N, int S, int Mrt, int L, int P, float inv_p, int QM, const T* __restrict__ ref, int RDf,
                       int tile_rows, int spec_on) {
  const int RD = RDf & 0xff;
  const bool r32 = (RDf & 0x100) != 0;
  constexpr int VEC = Vec16<T>::N;
  constexpr int LPR = D / VEC;
  constexpr int G = 32 / LPR;
  constexpr int ES = (int)sizeof(T);
  constexpr int ROWB = D * ES;
  const int M = MC > 0 ? MC : Mrt;
  const int MD = M * D;
  const int nw = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane / LPR, cl = lane % LPR;
  const int Lq = QM / M;
  const int LP = LPC > 0 ? LPC : L * P;

  extern __shared__ __align__(16) unsigned char msda_dyn_smem[];
  uint4* rec_a = reinterpret_cast<uint4*>(msda_dyn_smem) + (threadIdx.x & ~31);
  float2* rec_b = reinterpret_cast<float2*>(reinterpret_cast<uint4*>(msda_dyn_smem) + blockDim.x) + (threadIdx.x & ~31);
  T* tile = reinterpret_cast<T*>(msda_dyn_smem + (((size_t)blockDim.x * 24 + 127) & ~(size_t)127));
  T* zero_row = tile + (size_t)tile_rows * D;
  uint64_t* bar = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(zero_row) + ROWB);

  // the trailing levels [L0, L) that are stored back to back at the end of the image and fit the tile
  int L0 = L, row0 = S;
  for (int l = L - 1; l >= 0; --l) {
    const int st = __ldg(start + l), hw = __ldg(shapes + 2 * l) * __ldg(shapes + 2 * l + 1);
    if (st + hw != row0 || S - st > tile_rows) break;
    L0 = l;
    row0 = st;
  }
  const int staged_rows = S - row0;
  const int s0 = L0 * P;  // samples >= s0 of a unit read the tile

  if (threadIdx.x == 0) mbar_init(bar, 1);
  for (int i = threadIdx.x; i < ROWB / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(zero_row)[i] = 0u;
  __syncthreads();

  // ---- per-lane loop invariants: my sample's level ----
  const bool have = lane < LP;
  const LevelMeta lm = load_level_meta(shapes, start, lane, have, inv_p);
  const bool mine_staged = have && lane >= s0;  // lanes past L*P publish a zero-weight record on level 0 in GLOBAL memory
  // speculative regular-window gather (msda_fwd_gather_pass): needs every level at least 2 x 2 pixels
  const bool spec_ok = spec_on != 0 && __all_sync(0xffffffffu, !have || (lm.H >= 2 && lm.W >= 2));
  const unsigned xstep = mine_staged ? (unsigned)ROWB : (unsigned)(MD * ES);   // bytes per pixel step in x, of MY sample
  const unsigned ystep = (unsigned)lm.W * xstep;                               // bytes per row of my level
  const unsigned lvl_off = (mine_staged ? (unsigned)(lm.st - row0) : (unsigned)lm.st) * xstep;  // my level's first pixel
  const float fH = (float)lm.H, fW = (float)lm.W;
  const uint32_t tile_s = smem_u32(tile) + (uint32_t)(cl * VEC * ES);
  const int slane = have ? lane : 0;

  const int nchunk = (Lq + nw - 1) / nw;  // work item = (image, head, nw consecutive queries: one per warp)
  const int items = N * M * nchunk;       // < 2^31: host-checked
  const int i0 = (int)((long long)items * blockIdx.x / gridDim.x), i1 = (int)((long long)items * (blockIdx.x + 1) / gridDim.x);
  if (i0 >= i1) return;
  int bm = i0 / nchunk, c = i0 % nchunk;  // the only divisions: items are walked incrementally
  int cur_bm = -1;
  uint32_t parity = 0;

  // pointers of my warp's unit of item (bm, c); they advance by constants while the item stays in the same (image, head)
  const T* p_loc; const T* p_att; long long p_ref = 0; T* p_out;  // p_ref: element index into the reference points
  auto seek = [&](int bm_, int c_) {
    const int q = c_ * nw + w, b = bm_ / M, m = bm_ % M;
    const long long u = (long long)b * QM + (long long)q * M + m;
    p_loc = loc + (u * LP + slane) * 2;
    p_att = attn + u * LP + slane;
    p_out = out + u * D;
    if constexpr (FUSED) p_ref = (((long long)b * Lq + q) * L + lm.l) * RD;
  };
  const int d_att = nw * M * LP, d_out = nw * M * D, d_ref = nw * L * RD;
  auto fetch = [&]() {
    RawSample r;
    load_xy(p_loc, r.x, r.y);
    r.rx = r.ry = r.rw = r.rh = 0.f;
    if constexpr (FUSED) {
      r.a = have ? load_s(p_att) : -INFINITY;
      r.rx = load_ref(ref, p_ref, r32);
      r.ry = load_ref(ref, p_ref + 1, r32);
      if (RD != 2) { r.rw = load_ref(ref, p_ref + 2, r32); r.rh = load_ref(ref, p_ref + 3, r32); }
    } else {
      r.a = load_s(p_att);
    }
    return r;
  };
  seek(bm, c);
  RawSample raw;
  raw.x = raw.y = raw.a = raw.rx = raw.ry = raw.rw = raw.rh = 0.f;
  if (c * nw + w < Lq) raw = fetch();

  for (int item = i0; item < i1; ++item) {
    const int q = c * nw + w;
    const int this_bm = bm;
    const RawSample cur = raw;
    T* const o_cur = p_out;
    // advance to the next item and put its parameter loads in flight
    if (++c == nchunk) { c = 0; ++bm; seek(bm, c); }
    else { p_loc += 2 * d_att; p_att += d_att; p_out += d_out; if constexpr (FUSED) p_ref += d_ref; }
    if (item + 1 < i1 && c * nw + w < Lq) raw = fetch();

    if (this_bm != cur_bm) {  // CTA-uniform
      cur_bm = this_bm;
      if (staged_rows > 0) {
        __syncthreads();  // every warp has finished reading the previous slab
        if (threadIdx.x == 0) mbar_arrive_expect_tx(bar, (uint32_t)staged_rows * ROWB);
        const T* src = value + ((long long)(this_bm / M) * S + row0) * MD + (this_bm % M) * D;
        for (int r = threadIdx.x; r < staged_rows; r += blockDim.x)
          tma_bulk_g2s(tile + (size_t)r * D, src + (long long)r * MD, ROWB, bar);
        mbar_wait(bar, parity);
        parity ^= 1u;
      }
    }
    if (q >= Lq) continue;  // warp-uniform
    const int b = this_bm / M, m = this_bm % M;
    const char* vbc = reinterpret_cast<const char*>(value + (long long)b * S * MD + (m * D + cl * VEC));

    bool done = false;
    if (spec_ok) {
      const SampleParams sp = params_from_raw<FUSED>(cur, lm, RD, P, have);
      const float y = fma(sp.ly, fH, -0.5f), x = fma(sp.lx, fW, -0.5f);  // same roundings as make_geo
      const bool inside = have && y > -1.f && x > -1.f && y < fH && x < fW;
      const float fy = floorf(y), fx = floorf(x);
      int y0 = (int)fy, x0 = (int)fx;
      const float ly = y - fy, lx = x - fx, hy = 1.f - ly, hx = 1.f - lx;
      float wy0 = hy, wy1 = ly, wx0 = hx, wx1 = lx;
      if (y0 < 0) { y0 = 0; wy0 = ly; wy1 = 0.f; } else if (y0 > lm.H - 2) { y0 = lm.H - 2; wy1 = hy; wy0 = 0.f; }
      if (x0 < 0) { x0 = 0; wx0 = lx; wx1 = 0.f; } else if (x0 > lm.W - 2) { x0 = lm.W - 2; wx1 = hx; wx0 = 0.f; }
      if (!inside) { y0 = 0; x0 = 0; wy0 = 0.f; wy1 = 0.f; wx0 = 0.f; wx1 = 0.f; }
      const float a = inside ? sp.a : 0.f;
      const unsigned off00 = lvl_off + (unsigned)(y0 * lm.W + x0) * xstep;
      __syncwarp();
      rec_a[lane] = make_uint4(off00, ystep, __float_as_uint(wy0 * wx0 * a), __float_as_uint(wy0 * wx1 * a));
      rec_b[lane] = make_float2(wy1 * wx0 * a, wy1 * wx1 * a);
      __syncwarp();

      float2 acc[VEC / 2];
#pragma unroll
      for (int i = 0; i < VEC / 2; ++i) acc[i] = make_float2(0.f, 0.f);
      for (int k0 = 0; k0 < LP; k0 += G) {  // lanes >= LP hold zero-weight records on valid addresses
        const int src = k0 + g;
        const uint4 ra = rec_a[src];
        const float2 rb = rec_b[src];
        const float wt[4] = {__uint_as_float(ra.z), __uint_as_float(ra.w), rb.x, rb.y};
        uint4 v[4];
        if (k0 >= s0 && k0 + G <= LP) {  // whole round in the tile
          const uint32_t a0 = tile_s + ra.x, a1 = a0 + ra.y;
          v[0] = lds128_u32(a0); v[1] = lds128_u32(a0 + ROWB); v[2] = lds128_u32(a1); v[3] = lds128_u32(a1 + ROWB);
        } else if (k0 + G <= s0) {  // whole round in global memory
          const char* t0 = vbc + ra.x;
          const char* t1 = t0 + ra.y;
          const size_t mdb = (size_t)MD * ES;
          v[0] = ldg128(t0); v[1] = ldg128(t0 + mdb); v[2] = ldg128(t1); v[3] = ldg128(t1 + mdb);
        } else {  // a round that mixes staged and unstaged levels: generic loads (shared or global window)
          const bool stg = src >= s0 && src < LP;
          const char* t0 = (stg ? reinterpret_cast<const char*>(tile + cl * VEC) : vbc) + ra.x;
          const char* t1 = t0 + ra.y;
          const size_t xs = stg ? (size_t)ROWB : (size_t)MD * ES;
          v[0] = *reinterpret_cast<const uint4*>(t0); v[1] = *reinterpret_cast<const uint4*>(t0 + xs);
          v[2] = *reinterpret_cast<const uint4*>(t1); v[3] = *reinterpret_cast<const uint4*>(t1 + xs);
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          float f[VEC];
          Vec16<T>::unpack(v[t], f);
#pragma unroll
          for (int i = 0; i < VEC / 2; ++i) acc[i] = fma2(wt[t], make_float2(f[2 * i], f[2 * i + 1]), acc[i]);
        }
      }
      constexpr int N_OUT = GroupReduceScatter<VEC, LPR>::N_OUT;
      float o[N_OUT];
      int first;
      bool owner;
      msda_fwd_reduce<T, D>(acc, o, first, owner);
      bool bad = false;
#pragma unroll
      for (int k = 0; k < N_OUT; ++k) bad = bad || !(fabsf(o[k]) <= 3.402823466e38f);
      done = !__any_sync(0xffffffffu, bad);
      if (done && owner) store_vals<T, N_OUT>(o_cur + cl * VEC + first, o);
    }
    if (!done)  // a level narrower than 2 pixels, or non-finite sums: the exact zero-line path, out of line
      msda_fwd_unit_flagged<T, D, MC, FUSED, true>(value, shapes, start, loc, attn, out, S, M, L, P, inv_p, QM, ref, RDf, b, q * M + m, m);
  }
}

// ------------------------------------------------------------------------------------------------
// DETERMINISTIC accumulation of grad_value (C-ABI flag MSDA_BWD_DETERMINISTIC)
// ------------------------------------------------------------------------------------------------
// fp32 atomics make grad_value depend on the order in which the reds reach L2 (the reference's atomicAdd scatter has the same
// property: ms_deform_im2col_cuda.cuh:125-152).  In deterministic mode every contribution  c = weight * grad_out  (one fp32
// product, as in the default path) is converted to FIXED POINT,  q = rint(c * 2^k),  and accumulated with 64-bit INTEGER reds:
// integer addition is associative, so the sum is the same bit pattern whatever the order.  2^k is derived from max|grad_out|
// and max|attn| of the call (header of the workspace, written by msda_absmax_kernel) so that |q| <= 2^37 and 2^25 contributions
// per element cannot overflow; the absolute error per contribution is 2^-37 of the largest possible contribution -- finer than
// the fp32 accumulation it replaces.  msda_det_cvt_kernel turns the sums into grad_value (one rounding per element).
__device__ __forceinline__ int det_shift(const unsigned* __restrict__ hdr) {
  const int eg = (int)((__ldg(hdr) >> 23) & 255u) - 127;      // floor(log2(max |grad_out|)); -127 for zero / denormals
  const int ea = (int)((__ldg(hdr + 1) >> 23) & 255u) - 127;  // floor(log2(max |attn|))
  const int k = 36 - (eg + 1) - max(ea + 1, 0);
  return max(-250, min(250, k));  // applied as two exact power-of-two factors 2^(k/2) * 2^(k - k/2)
}
__device__ __forceinline__ float pow2_int(int k) { return __int_as_float((k + 127) << 23); }  // exact 2^k, -126 <= k <= 127

template <typename T>
__global__ void __launch_bounds__(256) msda_absmax_kernel(const T* __restrict__ a, long long na, const T* __restrict__ b, long long nb,
                                                          unsigned* __restrict__ hdr) {
  float ma = 0.f, mb = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x, i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (long long i = i0; i < na; i += stride) ma = fmaxf(ma, fabsf((float)to_acc(a[i])));
  for (long long i = i0; i < nb; i += stride) mb = fmaxf(mb, fabsf((float)to_acc(b[i])));
  ma = warp_max(ma);
  mb = warp_max(mb);
  if ((threadIdx.x & 31) == 0) {  // non-negative floats order like their bit patterns
    atomicMax(hdr, __float_as_uint(ma));
    atomicMax(hdr + 1, __float_as_uint(mb));
  }
}

template <typename T>
__global__ void __launch_bounds__(256) msda_det_cvt_kernel(const long long* __restrict__ src, T* __restrict__ dst, long long n,
                                                           const unsigned* __restrict__ hdr) {
  const int k = det_shift(hdr);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = from_acc<T>((float)scalbn((double)src[i], -k));
}

// ------------------------------------------------------------------------------------------------
// VECTOR BACKWARD
// ------------------------------------------------------------------------------------------------
// grad_value accumulates in fp32 (`gv`): the caller's tensor for T=float, a workspace for 16-bit T.
//
// Per round a lane group gathers the 4 taps of its sample, forms the four dot products <grad_out, tap> over its
// channels (FFMA2), reduces them across the LPR lanes and hands them back to the lane that owns the sample (which
// combines them into grad_attn / grad_loc and writes those coalesced at the end); then it scatters
// weight * attention * grad_out into grad_value with 16-byte `red.global.add.v4.f32` (fire-and-forget, resolved in
// L2, so the scatter of one round overlaps the gathers of the next).
//
// Programmatic dependent launch: the kernel is released while the zero-fill of grad_value (previous kernel in the
// stream, which triggers `griddepcontrol.launch_dependents` at its start) is still running; everything up to the first
// scatter -- parameter loads, geometry, the first gather round -- overlaps the fill, and `griddepcontrol.wait`
// orders the first `red` after its completion.
// FUSED: `loc` / `attn` are the raw offsets / logits, `gloc` / `gattn` receive the gradients w.r.t. THOSE (softmax and
// location arithmetic differentiated in the kernel), `gref` (fp32, pre-zeroed, may be null) accumulates the gradient
// w.r.t. the reference points with scalar reds (M*P contributions per element).
// DET: `gv` is the int64 fixed-point image (see above), `det_hdr` its header; launched WITHOUT programmatic dependency.
template <typename T, int D, int MC, int U, bool FUSED, bool DET = false>
__global__ void __maxnreg__(U == 1 ? (FUSED ? MSDA_BWD_MAX_REGS_FUSED : (DET ? 64 : MSDA_BWD_MAX_REGS)) : (U == 2 ? 80 : 128))
msda_bwd_sg_kernel(const T* __restrict__ grad_out, const T* __restrict__ value,
                   const int32_t* __restrict__ shapes, const int32_t* __restrict__ start,
                   const T* __restrict__ loc, const T* __restrict__ attn, float* __restrict__ gv,
                   T* __restrict__ gloc, T* __restrict__ gattn,
                   int S, int Mrt, int L, int P, float inv_p, int QM, const T* __restrict__ ref, int RDf,
                   float* __restrict__ gref, int head_major, const unsigned* __restrict__ det_hdr = nullptr) {
  const int RD = RDf & 0xff;
  const bool r32 = (RDf & 0x100) != 0;
  using IO = VecIO<T, BwdGranule<T>::VB>;
  constexpr int VEC = IO::N;
  constexpr int LPR = D / VEC;
  constexpr int G = 32 / LPR;
  static_assert(D % VEC == 0 && LPR >= 1 && LPR <= 32 && (LPR & (LPR - 1)) == 0, "unsupported D for the vector path");

  const int M = MC > 0 ? MC : Mrt;
  const int MD = M * D;
  const int lane = threadIdx.x & 31;
  const bool two_pass = (head_major & 2) != 0;  // bit 1 of the scheduling word: gather pass, fence, scatter pass
  const bool l32 = sizeof(T) != 4 && !FUSED && (head_major & 4) != 0;  // bits 2 / 3: fp32 locations / weights (and their gradients)
  const bool a32 = sizeof(T) != 4 && !FUSED && (head_major & 8) != 0;  // next to 16-bit value (mixed precision, see load_params)
  int uq, m;
  if (!unit_of_warp(M, QM, head_major & 1, uq, m)) return;  // warp-uniform (an exited warp needs no fence)
  const int g = lane / LPR, cl = lane % LPR;
  const int LP = L * P;
  const long long u = (long long)blockIdx.y * QM + uq;
  const T* __restrict__ u_loc = loc + u * (LP * 2) * (l32 ? 2 : 1);
  const T* __restrict__ u_att = attn + u * LP * (a32 ? 2 : 1);
  const long long voff = (long long)blockIdx.y * S * MD + (m * D + cl * VEC);
  const T* __restrict__ vb = value + voff;
  float* __restrict__ gb = gv + voff;
  float det_s1 = 0.f, det_s2 = 0.f;
  if constexpr (DET) {
    const int k = det_shift(det_hdr);
    det_s1 = pow2_int(k / 2);
    det_s2 = pow2_int(k - k / 2);
  }

  float go[VEC];
  IO::unpack(IO::load(grad_out + u * D + cl * VEC), go);
  bool fenced = false;

  for (int base = 0; base < LP; base += 32) {
    SampleGeo sg;
    Geo<float> ge;
    SampleParams sp;
    int lvl = 0;
    float odx = 0.f, ody = 0.f;
    const bool have = base + lane < LP;
    if constexpr (FUSED) {
      sp = fused_params<T>(u_loc, u_att, ref, ((long long)blockIdx.y * (QM / M) + uq / M) * (L * RD), RD, r32, shapes, start, lane, have, inv_p, P, lvl, odx, ody);
    } else {
      sp = load_params<T>(u_loc, u_att, shapes, start, base + lane, have, inv_p, l32, a32);
    }
    finish_geometry(sp, have, MD, sg, ge);
    const float a = (FUSED || ge.inside) ? sp.a : 0.f;  // outside sample: every gradient is exactly zero (FUSED keeps the softmax weight)
    const int H = sp.H, W = sp.W;
    const float w00 = ge.hy * ge.hx * a, w01 = ge.hy * ge.lx * a, w10 = ge.ly * ge.hx * a, w11 = ge.ly * ge.lx * a;
    const int cnt = min(32, LP - base);
    float r00 = 0.f, r01 = 0.f, r10 = 0.f, r11 = 0.f;  // <grad_out, tap> of MY sample, from the group that gathered it
    // which samples the lane groups handle in round k0; true if every tap of the round is valid (warp-uniform)
    auto round_index = [&](int k0, int (&off)[U], int (&rsf)[U]) -> bool {
      bool full = true;
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const int src = k0 + j * G + g;
        off[j] = __shfl_sync(0xffffffffu, sg.off00, src);
        rsf[j] = __shfl_sync(0xffffffffu, sg.rsf, src);
        if (src >= cnt) rsf[j] = 0;  // shfl wraps modulo 32: a lane group past the end must not gather / scatter
        full = full && ((rsf[j] & 15) == 15);
      }
      return __all_sync(0xffffffffu, full);
    };
    // gather half of a round: <grad_out, tap> of the round's samples, returned to the lanes that own them
    auto gather_round = [&](int k0, const int (&off)[U], const int (&rsf)[U], bool all_ok) {
      typename IO::Raw v[U][4];
      if (all_ok) {  // loads duplicated per branch: here the x+1 taps are immediate offsets of the x taps
#pragma unroll
        for (int j = 0; j < U; ++j) {
          const T* t0 = vb + off[j];
          const T* t1 = t0 + (rsf[j] >> 4);
          v[j][0] = IO::load(t0); v[j][1] = IO::load(t0 + MD); v[j][2] = IO::load(t1); v[j][3] = IO::load(t1 + MD);
        }
      } else {
#pragma unroll
        for (int j = 0; j < U; ++j) {
          const T* tp[4];
          tap_pointers<T>(vb, off[j], rsf[j], MD, false, tp);
#pragma unroll
          for (int t = 0; t < 4; ++t) v[j][t] = IO::load(tp[t]);
        }
      }
#pragma unroll
      for (int j = 0; j < U; ++j) {
        float d[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          float f[VEC];
          IO::unpack(v[j][t], f);
          float2 p = make_float2(0.f, 0.f);
#pragma unroll
          for (int i = 0; i < VEC / 2; ++i) p = __ffma2_rn(make_float2(go[2 * i], go[2 * i + 1]), make_float2(f[2 * i], f[2 * i + 1]), p);
          d[t] = p.x + p.y;
        }
#pragma unroll
        for (int o = 1; o < LPR; o <<= 1)  // channel sums over the LPR lanes of this group
#pragma unroll
          for (int t = 0; t < 4; ++t) d[t] += __shfl_xor_sync(0xffffffffu, d[t], o);
        // hand the sums back to the lane that owns the sample: lane (k0 + j*G + g') reads from group g'
        const int rel = lane - k0 - j * G;
        const int from = (rel & (G - 1)) * LPR;
        const float t00 = __shfl_sync(0xffffffffu, d[0], from);
        const float t01 = __shfl_sync(0xffffffffu, d[1], from);
        const float t10 = __shfl_sync(0xffffffffu, d[2], from);
        const float t11 = __shfl_sync(0xffffffffu, d[3], from);
        if (rel >= 0 && rel < G) { r00 = t00; r01 = t01; r10 = t10; r11 = t11; }
      }
    };
    // scatter half of a round: needs the zero-fill of grad_value (previous kernel) complete and visible; it does not
    // depend on the gathered taps at all (contribution = weight * attention * grad_out)
    auto scatter_round = [&](int k0, const int (&off)[U], const int (&rsf)[U]) {
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const int src = k0 + j * G + g;
        float w[4];
        w[0] = __shfl_sync(0xffffffffu, w00, src);
        w[1] = __shfl_sync(0xffffffffu, w01, src);
        w[2] = __shfl_sync(0xffffffffu, w10, src);
        w[3] = __shfl_sync(0xffffffffu, w11, src);
        if constexpr (DET) {
          unsigned long long* g0 = reinterpret_cast<unsigned long long*>(gv) + (voff + off[j]);
          unsigned long long* g1 = g0 + (rsf[j] >> 4);
          unsigned long long* gp[4] = {g0, g0 + MD, g1, g1 + MD};
#pragma unroll
          for (int t = 0; t < 4; ++t)
            if (rsf[j] & (1 << t)) {
#pragma unroll
              for (int i = 0; i < VEC; ++i)  // same fp32 product as the default path, then exact scaling by 2^k and one rint
                atomicAdd(gp[t] + i, (unsigned long long)__float2ll_rn(__fmul_rn(__fmul_rn(__fmul_rn(w[t], go[i]), det_s1), det_s2)));
            }
        } else {
          float* g0 = gb + off[j];
          float* g1 = g0 + (rsf[j] >> 4);
          float* gp[4] = {g0, g0 + MD, g1, g1 + MD};
#pragma unroll
          for (int t = 0; t < 4; ++t)
            if (rsf[j] & (1 << t)) {
#pragma unroll
              for (int i = 0; i < VEC; i += 4) {
                const float2 lo = __fmul2_rn(make_float2(w[t], w[t]), make_float2(go[i], go[i + 1]));
                const float2 hi = __fmul2_rn(make_float2(w[t], w[t]), make_float2(go[i + 2], go[i + 3]));
                red_add_v4(gp[t] + i, lo.x, lo.y, hi.x, hi.y);
              }
            }
        }
      }
    };
    // grad_attn / grad_loc of MY sample from the four returned dot products
    auto epilogue = [&]() {
      const float top = ge.hx * r00 + ge.lx * r01;  // interpolated along x on row y0
      const float bot = ge.hx * r10 + ge.lx * r11;  // ... on row y0 + 1
      const float g_a = ge.hy * top + ge.ly * bot;                                  // d out / d A_i
      const float g_lx = (float)W * a * (ge.hy * (r01 - r00) + ge.ly * (r11 - r10));  // d out / d loc_x
      const float g_ly = (float)H * a * (bot - top);                                  // d out / d loc_y
      const long long sidx = u * LP + base + lane;
      if constexpr (!FUSED) {
        if (have) {  // coalesced: 32 consecutive samples of the unit
          if (a32) reinterpret_cast<float*>(gattn)[sidx] = g_a;
          else gattn[sidx] = from_acc<T>(g_a);
          if (l32) store_xy(reinterpret_cast<float*>(gloc) + 2 * sidx, g_lx, g_ly);
          else store_xy(gloc + 2 * sidx, g_lx, g_ly);
        }
      } else {
        // softmax backward over the unit's samples:  g_logit_i = A_i * (g_A_i - sum_j A_j g_A_j)
        const float dot = warp_sum(have ? a * g_a : 0.f);
        if (have) {
          gattn[sidx] = from_acc<T>(a * (g_a - dot));
          const long long rbase = (((long long)blockIdx.y * (QM / M) + uq / M) * L + lvl) * RD;
          if (RD == 2) {
            store_xy(gloc + 2 * sidx, __fdiv_rn(g_lx, (float)W), __fdiv_rn(g_ly, (float)H));
            if (gref != nullptr) {
              atomicAdd(gref + rbase, g_lx);
              atomicAdd(gref + rbase + 1, g_ly);
            }
          } else {
            const float rw = load_ref(ref, rbase + 2, r32), rh = load_ref(ref, rbase + 3, r32);
            store_xy(gloc + 2 * sidx, __fdiv_rn(g_lx * 0.5f * rw, (float)P), __fdiv_rn(g_ly * 0.5f * rh, (float)P));
            if (gref != nullptr) {
              atomicAdd(gref + rbase, g_lx);
              atomicAdd(gref + rbase + 1, g_ly);
              atomicAdd(gref + rbase + 2, g_lx * 0.5f * odx);
              atomicAdd(gref + rbase + 3, g_ly * 0.5f * ody);
            }
          }
        }
      }
    };
    if (two_pass) {
      // every gather round (a dependent-latency chain per round) runs while the zero-fill is still in flight; behind the
      // fence only the reds remain
      for (int k0 = 0; k0 < cnt; k0 += G * U) {
        int off[U], rsf[U];
        const bool all_ok = round_index(k0, off, rsf);
        gather_round(k0, off, rsf, all_ok);
      }
      epilogue();
      if (!fenced) {
        pdl_wait();
        fenced = true;
      }
      for (int k0 = 0; k0 < cnt; k0 += G * U) {
        int off[U], rsf[U];
        round_index(k0, off, rsf);
        scatter_round(k0, off, rsf);
      }
    } else {
      for (int k0 = 0; k0 < cnt; k0 += G * U) {
        int off[U], rsf[U];
        const bool all_ok = round_index(k0, off, rsf);
        gather_round(k0, off, rsf, all_ok);
        if (!fenced) {
          pdl_wait();
          fenced = true;
        }
        scatter_round(k0, off, rsf);
      }
      epilogue();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// GENERIC kernels: any D, L, P; T in {float, double, bf16, half}.  Warp per unit, lanes over channels.
// TL / TA: storage type of sampling_loc / attn_weight and of their gradients (float next to a 16-bit T: mixed precision).
// ------------------------------------------------------------------------------------------------
template <typename T, typename TL = T, typename TA = T>
__global__ void __launch_bounds__(256)
msda_fwd_generic_kernel(const T* __restrict__ value, const int32_t* __restrict__ shapes,
                        const int32_t* __restrict__ start, const TL* __restrict__ loc,
                        const TA* __restrict__ attn, T* __restrict__ out,
                        int S, int M, int D, int L, int Lq, int P, long long units) {
  using A = typename AccOf<T>::type;
  const int lane = threadIdx.x & 31;
  const long long u = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (u >= units) return;
  const int m = (int)(u % M);
  const long long b = (u / M) / Lq;
  const int LP = L * P;
  const long long MD = (long long)M * D;
  const TL* __restrict__ u_loc = loc + u * LP * 2;
  const TA* __restrict__ u_att = attn + u * LP;
  const T* __restrict__ vb = value + b * (long long)S * MD + (long long)m * D;

  for (int c0 = 0; c0 < D; c0 += 32) {
    const int c = c0 + lane;
    const bool cok = c < D;
    A acc = (A)0;
    for (int l = 0; l < L; ++l) {
      const int H = __ldg(shapes + 2 * l), W = __ldg(shapes + 2 * l + 1);
      const T* __restrict__ lv = vb + (long long)__ldg(start + l) * MD + c;
      for (int p = 0; p < P; ++p) {
        const int s = l * P + p;
        const Geo<A> ge = make_geo<A>(to_acc(u_loc[2 * s]), to_acc(u_loc[2 * s + 1]), H, W, true);
        const A a = ge.inside ? to_acc(u_att[s]) : (A)0;
        const T* t0 = lv + (long long)ge.row00 * MD;
        const long long rs = (long long)W * MD;
        const A v00 = (cok && ge.ok00) ? to_acc(t0[0]) : (A)0;
        const A v01 = (cok && ge.ok01) ? to_acc(t0[MD]) : (A)0;
        const A v10 = (cok && ge.ok10) ? to_acc(t0[rs]) : (A)0;
        const A v11 = (cok && ge.ok11) ? to_acc(t0[rs + MD]) : (A)0;
        acc += (ge.hy * ge.hx * v00 + ge.hy * ge.lx * v01 + ge.ly * ge.hx * v10 + ge.ly * ge.lx * v11) * a;
      }
    }
    if (cok) out[u * D + c] = from_acc<T>(acc);
  }
}

template <typename T, typename TL = T, typename TA = T>
__global__ void __launch_bounds__(256)
msda_bwd_generic_kernel(const T* __restrict__ grad_out, const T* __restrict__ value,
                        const int32_t* __restrict__ shapes, const int32_t* __restrict__ start,
                        const TL* __restrict__ loc, const TA* __restrict__ attn,
                        typename AccOf<T>::type* __restrict__ gv, TL* __restrict__ gloc, TA* __restrict__ gattn,
                        int S, int M, int D, int L, int Lq, int P, long long units) {
  using A = typename AccOf<T>::type;
  const int lane = threadIdx.x & 31;
  const long long u = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (u >= units) return;
  const int m = (int)(u % M);
  const long long b = (u / M) / Lq;
  const int LP = L * P;
  const long long MD = (long long)M * D;
  const TL* __restrict__ u_loc = loc + u * LP * 2;
  const TA* __restrict__ u_att = attn + u * LP;
  const T* __restrict__ u_go = grad_out + u * D;
  const long long voff = b * (long long)S * MD + (long long)m * D;

  for (int l = 0; l < L; ++l) {
    const int H = __ldg(shapes + 2 * l), W = __ldg(shapes + 2 * l + 1);
    const long long loff = voff + (long long)__ldg(start + l) * MD;
    for (int p = 0; p < P; ++p) {
      const int s = l * P + p;
      const Geo<A> ge = make_geo<A>(to_acc(u_loc[2 * s]), to_acc(u_loc[2 * s + 1]), H, W, true);
      const A a = ge.inside ? to_acc(u_att[s]) : (A)0;
      const long long o00 = loff + (long long)ge.row00 * MD;
      const long long rs = (long long)W * MD;
      const A w00 = ge.hy * ge.hx * a, w01 = ge.hy * ge.lx * a, w10 = ge.ly * ge.hx * a, w11 = ge.ly * ge.lx * a;
      A s_a = (A)0, s_x = (A)0, s_y = (A)0;
      for (int c = lane; c < D; c += 32) {
        const A go = to_acc(u_go[c]);
        const A v00 = ge.ok00 ? to_acc(value[o00 + c]) : (A)0;
        const A v01 = ge.ok01 ? to_acc(value[o00 + MD + c]) : (A)0;
        const A v10 = ge.ok10 ? to_acc(value[o00 + rs + c]) : (A)0;
        const A v11 = ge.ok11 ? to_acc(value[o00 + rs + MD + c]) : (A)0;
        const A top = ge.hx * v00 + ge.lx * v01, bot = ge.hx * v10 + ge.lx * v11;
        s_a += go * (ge.hy * top + ge.ly * bot);
        s_y += go * (bot - top);
        s_x += go * (ge.hy * (v01 - v00) + ge.ly * (v11 - v10));
        if (ge.ok00) atomicAdd(gv + o00 + c, w00 * go);
        if (ge.ok01) atomicAdd(gv + o00 + MD + c, w01 * go);
        if (ge.ok10) atomicAdd(gv + o00 + rs + c, w10 * go);
        if (ge.ok11) atomicAdd(gv + o00 + rs + MD + c, w11 * go);
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        s_a += __shfl_xor_sync(0xffffffffu, s_a, off);
        s_x += __shfl_xor_sync(0xffffffffu, s_x, off);
        s_y += __shfl_xor_sync(0xffffffffu, s_y, off);
      }
      if (lane == 0) {
        const long long sidx = u * LP + s;
        gattn[sidx] = from_acc<TA>(s_a);
        gloc[2 * sidx] = from_acc<TL>((A)W * a * s_x);
        gloc[2 * sidx + 1] = from_acc<TL>((A)H * a * s_y);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// helpers: zero fill (128-bit stores, grid-stride) and fp32 -> 16-bit conversion of grad_value
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) msda_zero_kernel(uint4* __restrict__ p, long long n16, unsigned char* __restrict__ tail, int ntail) {
  // Launched NORMALLY (everything earlier in the stream is complete when it starts); releases the programmatically-
  // dependent backward kernel at once, whose gather pass reads the op's inputs while this fill is running.
  pdl_trigger();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) p[i] = make_uint4(0, 0, 0, 0);
  if (blockIdx.x == 0 && (int)threadIdx.x < ntail) tail[threadIdx.x] = 0;
}

// Zero fill through the TMA store path: each CTA clears one shared-memory buffer once and then ONE thread streams it
// out with `cp.async.bulk.global.shared::cta` (bulk stores bypass the LSU store path and occupy one warp and no
// registers worth mentioning, so the programmatically-dependent backward kernel gets the SMs to itself).
// `chunk` bytes per bulk store (multiple of 16, <= the dynamic shared memory size); n16 = number of 16-byte pieces.
__global__ void __launch_bounds__(128) msda_zero_tma_kernel(unsigned char* __restrict__ p, long long n16, int ntail, int chunk) {
  extern __shared__ __align__(128) unsigned char zbuf[];
  pdl_trigger();
  for (int i = threadIdx.x; i < chunk / 16; i += blockDim.x) reinterpret_cast<uint4*>(zbuf)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const long long bytes = n16 * 16;
  if (threadIdx.x == 0) {
    const uint32_t src = (uint32_t)__cvta_generic_to_shared(zbuf);
    for (long long off = (long long)blockIdx.x * chunk; off < bytes; off += (long long)gridDim.x * chunk) {
      const long long left = bytes - off;
      const uint32_t n = (uint32_t)(left < chunk ? left : chunk);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p + off), "r"(src), "r"(n) : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  if (blockIdx.x == 0 && (int)threadIdx.x < ntail) p[bytes + threadIdx.x] = 0;
}

template <typename T>
__global__ void __launch_bounds__(256) msda_cvt_kernel(const float* __restrict__ src, T* __restrict__ dst, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = from_acc<T>(src[i]);
}

}  // namespace msda
