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
#define MSDA_MAX_THREADS 256  // warps_per_block <= 8: leaves ptxas the register room for 16 x 128-bit loads in flight
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
};

__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }

// `valid` = this lane group really has a sample (tail predicate).
template <typename A>
__device__ __forceinline__ Geo<A> make_geo(A loc_x, A loc_y, int H, int W, bool valid) {
  Geo<A> g;
  // separate multiply and subtract (no FMA contraction) = the reference's rounding of the coordinate
  const A y = sub_rn(mul_rn(loc_y, (A)H), (A)0.5);
  const A x = sub_rn(mul_rn(loc_x, (A)W), (A)0.5);
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

__device__ __forceinline__ void store_xy(float* p, float x, float y) { *reinterpret_cast<float2*>(p) = make_float2(x, y); }
__device__ __forceinline__ void store_xy(__nv_bfloat16* p, float x, float y) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(x, y);
}
__device__ __forceinline__ void store_xy(__half* p, float x, float y) { *reinterpret_cast<__half2*>(p) = __floats2half2_rn(x, y); }

// vectorised fp32 reduction into global memory (sm_90+): one 16-byte L2 atomic instead of four
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ------------------------------------------------------------------------------------------------
// VECTOR kernels -- shared pieces
// ------------------------------------------------------------------------------------------------
// Invalid taps (outside the level, or sample outside the window) read this line instead of being
// predicated off: unconditional loads let ptxas issue all 4*U gathers of a lane back to back
// (checked with tools/sass_summary.py); a predicated load sequence was serialised into load->FMA pairs.
__device__ __align__(16) const unsigned int g_zero_line[4] = {0u, 0u, 0u, 0u};

// Per-warp level table: lane l holds (H_l, W_l, start_l); read back with __shfl_sync.  Needs L <= 32.
struct LevelTable {
  int H, W, st;
  __device__ __forceinline__ void load(const int32_t* __restrict__ shapes, const int32_t* __restrict__ start, int L, int lane) {
    H = 1; W = 1; st = 0;
    if (lane < L) {
      H = __ldg(shapes + 2 * lane);
      W = __ldg(shapes + 2 * lane + 1);
      st = __ldg(start + lane);
    }
  }
};

// level of sample s (= s / P) without an integer division: exact for s, P < 2^20
__device__ __forceinline__ int level_of(int s, float inv_p) { return __float2int_rz(((float)s + 0.5f) * inv_p); }

// ------------------------------------------------------------------------------------------------
// VECTOR FORWARD
// ------------------------------------------------------------------------------------------------
// T in {float, bf16, half}; D*sizeof(T)/16 = LPR in {1,2,4,8,16,32}; U samples in flight per lane group.
// Dependent-latency chain per warp: {loc, attn, level table} -> 4*U tap rows -> shuffles -> store.
template <typename T, int D, int U>
__global__ void __launch_bounds__(MSDA_MAX_THREADS)
msda_fwd_vec_kernel(const T* __restrict__ value, const int32_t* __restrict__ shapes,
                    const int32_t* __restrict__ start, const T* __restrict__ loc,
                    const T* __restrict__ attn, T* __restrict__ out,
                    int S, int M, int L, int Lq, int P, float inv_p, long long units) {
  constexpr int VEC = Vec16<T>::N;
  constexpr int LPR = D / VEC;
  constexpr int G = 32 / LPR;
  static_assert(D % VEC == 0 && LPR >= 1 && LPR <= 32 && (LPR & (LPR - 1)) == 0, "unsupported D for the vector path");

  const int lane = threadIdx.x & 31;
  const long long u = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (u >= units) return;  // warp-uniform
  const int g = lane / LPR, cl = lane % LPR;
  const int m = (int)(u % M);
  const long long b = (u / M) / Lq;
  const int LP = L * P;
  const int MD = M * D;
  const T* __restrict__ u_loc = loc + u * LP * 2;
  const T* __restrict__ u_att = attn + u * LP;
  const T* __restrict__ vb = value + b * (long long)S * MD + m * D + cl * VEC;
  const T* zp = reinterpret_cast<const T*>(g_zero_line);

  LevelTable lt;
  lt.load(shapes, start, L, lane);

  float acc[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] = 0.f;

  for (int s0 = g; s0 < LP + g; s0 += G * U) {  // same trip count for every group (warp-uniform loop)
    float lx[U], ly[U], a[U];
    int sc[U];
#pragma unroll
    for (int j = 0; j < U; ++j) {  // phase 1: the small loads, all independent
      const int s = s0 + j * G;
      sc[j] = s < LP ? s : -1;
      const int si = s < LP ? s : 0;
      load_xy(u_loc + 2 * si, lx[j], ly[j]);
      a[j] = load_s(u_att + si);
    }
    const T* tp[U][4];
    float w[U][4];
#pragma unroll
    for (int j = 0; j < U; ++j) {  // phase 2: geometry -> 4 tap addresses + weights
      const bool valid = sc[j] >= 0;
      const int l = valid ? level_of(sc[j], inv_p) : 0;
      const int H = __shfl_sync(0xffffffffu, lt.H, l), W = __shfl_sync(0xffffffffu, lt.W, l);
      const int st = __shfl_sync(0xffffffffu, lt.st, l);
      const Geo<float> ge = make_geo<float>(lx[j], ly[j], H, W, valid);
      w[j][0] = ge.hy * ge.hx * a[j]; w[j][1] = ge.hy * ge.lx * a[j];
      w[j][2] = ge.ly * ge.hx * a[j]; w[j][3] = ge.ly * ge.lx * a[j];
      const T* t0 = vb + ((long long)st + ge.row00) * MD;
      const int rs = W * MD;
      tp[j][0] = ge.ok00 ? t0 : zp;
      tp[j][1] = ge.ok01 ? t0 + MD : zp;
      tp[j][2] = ge.ok10 ? t0 + rs : zp;
      tp[j][3] = ge.ok11 ? t0 + rs + MD : zp;
    }
    uint4 v[U][4];
#pragma unroll
    for (int j = 0; j < U; ++j)  // phase 3: 4*U independent 128-bit gathers
#pragma unroll
      for (int t = 0; t < 4; ++t) v[j][t] = ldg128(tp[j][t]);
#pragma unroll
    for (int j = 0; j < U; ++j)  // phase 4: weighted accumulation
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float f[VEC];
        Vec16<T>::unpack(v[j][t], f);
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = fmaf(w[j][t], f[i], acc[i]);
      }
  }
#pragma unroll
  for (int off = LPR; off < 32; off <<= 1)
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], off);
  if (g == 0) *reinterpret_cast<uint4*>(out + u * D + cl * VEC) = Vec16<T>::pack(acc);
}

// ------------------------------------------------------------------------------------------------
// VECTOR BACKWARD
// ------------------------------------------------------------------------------------------------
// grad_value accumulates in fp32 (`gv`): the caller's tensor for T=float, a workspace for 16-bit T.
// Scatter = 16-byte `red.global.add.v4.f32` per tap per lane (no return value, resolved in L2).
template <typename T, int D, int U>
__global__ void __launch_bounds__(MSDA_MAX_THREADS)
msda_bwd_vec_kernel(const T* __restrict__ grad_out, const T* __restrict__ value,
                    const int32_t* __restrict__ shapes, const int32_t* __restrict__ start,
                    const T* __restrict__ loc, const T* __restrict__ attn, float* __restrict__ gv,
                    T* __restrict__ gloc, T* __restrict__ gattn,
                    int S, int M, int L, int Lq, int P, float inv_p, long long units) {
  constexpr int VEC = Vec16<T>::N;
  constexpr int LPR = D / VEC;
  constexpr int G = 32 / LPR;
  static_assert(D % VEC == 0 && LPR >= 1 && LPR <= 32 && (LPR & (LPR - 1)) == 0, "unsupported D for the vector path");

  const int lane = threadIdx.x & 31;
  const long long u = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (u >= units) return;
  const int g = lane / LPR, cl = lane % LPR;
  const int m = (int)(u % M);
  const long long b = (u / M) / Lq;
  const int LP = L * P;
  const int MD = M * D;
  const T* __restrict__ u_loc = loc + u * LP * 2;
  const T* __restrict__ u_att = attn + u * LP;
  const long long voff = b * (long long)S * MD + m * D + cl * VEC;
  const T* __restrict__ vb = value + voff;
  float* __restrict__ gb = gv + voff;
  const T* zp = reinterpret_cast<const T*>(g_zero_line);

  LevelTable lt;
  lt.load(shapes, start, L, lane);

  float go[VEC];
  Vec16<T>::unpack(ldg128(grad_out + u * D + cl * VEC), go);

  for (int s0 = g; s0 < LP + g; s0 += G * U) {
    float lx[U], ly[U], a[U];
    int sc[U];
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const int s = s0 + j * G;
      sc[j] = s < LP ? s : -1;
      const int si = s < LP ? s : 0;
      load_xy(u_loc + 2 * si, lx[j], ly[j]);
      a[j] = load_s(u_att + si);
    }
    Geo<float> ge[U];
    long long row[U];
    int Hs[U];
    const T* tp[U][4];
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const bool valid = sc[j] >= 0;
      const int l = valid ? level_of(sc[j], inv_p) : 0;
      Hs[j] = __shfl_sync(0xffffffffu, lt.H, l);
      const int W = __shfl_sync(0xffffffffu, lt.W, l);
      const int st = __shfl_sync(0xffffffffu, lt.st, l);
      ge[j] = make_geo<float>(lx[j], ly[j], Hs[j], W, valid);
      row[j] = ((long long)st + ge[j].row00) * MD;
      const T* t0 = vb + row[j];
      const int rs = W * MD;
      tp[j][0] = ge[j].ok00 ? t0 : zp;
      tp[j][1] = ge[j].ok01 ? t0 + MD : zp;
      tp[j][2] = ge[j].ok10 ? t0 + rs : zp;
      tp[j][3] = ge[j].ok11 ? t0 + rs + MD : zp;
    }
    uint4 v[U][4];
#pragma unroll
    for (int j = 0; j < U; ++j)
#pragma unroll
      for (int t = 0; t < 4; ++t) v[j][t] = ldg128(tp[j][t]);
#pragma unroll
    for (int j = 0; j < U; ++j) {
      float f00[VEC], f01[VEC], f10[VEC], f11[VEC];
      Vec16<T>::unpack(v[j][0], f00);
      Vec16<T>::unpack(v[j][1], f01);
      Vec16<T>::unpack(v[j][2], f10);
      Vec16<T>::unpack(v[j][3], f11);
      const float hy = ge[j].hy, hx = ge[j].hx, ly_ = ge[j].ly, lx_ = ge[j].lx;
      float s_a = 0.f, s_x = 0.f, s_y = 0.f;
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const float top = hx * f00[i] + lx_ * f01[i];  // interpolated along x on row y0
        const float bot = hx * f10[i] + lx_ * f11[i];  // ... on row y0+1
        s_a = fmaf(go[i], hy * top + ly_ * bot, s_a);
        s_y = fmaf(go[i], bot - top, s_y);
        s_x = fmaf(go[i], hy * (f01[i] - f00[i]) + ly_ * (f11[i] - f10[i]), s_x);
      }
      {  // scatter into grad_value: bilinear weight * attention * grad_out
        float* g0 = gb + row[j];
        const int rs = ge[j].W * MD;
        const float w00 = hy * hx * a[j], w01 = hy * lx_ * a[j], w10 = ly_ * hx * a[j], w11 = ly_ * lx_ * a[j];
#pragma unroll
        for (int i = 0; i < VEC; i += 4) {
          if (ge[j].ok00) red_add_v4(g0 + i, w00 * go[i], w00 * go[i + 1], w00 * go[i + 2], w00 * go[i + 3]);
          if (ge[j].ok01) red_add_v4(g0 + MD + i, w01 * go[i], w01 * go[i + 1], w01 * go[i + 2], w01 * go[i + 3]);
          if (ge[j].ok10) red_add_v4(g0 + rs + i, w10 * go[i], w10 * go[i + 1], w10 * go[i + 2], w10 * go[i + 3]);
          if (ge[j].ok11) red_add_v4(g0 + rs + MD + i, w11 * go[i], w11 * go[i + 1], w11 * go[i + 2], w11 * go[i + 3]);
        }
      }
#pragma unroll
      for (int off = 1; off < LPR; off <<= 1) {  // channel sums over the LPR lanes of this group
        s_a += __shfl_xor_sync(0xffffffffu, s_a, off);
        s_x += __shfl_xor_sync(0xffffffffu, s_x, off);
        s_y += __shfl_xor_sync(0xffffffffu, s_y, off);
      }
      if (cl == 0 && sc[j] >= 0) {
        const long long sidx = u * LP + sc[j];
        gattn[sidx] = from_acc<T>(s_a);
        T* gl = gloc + 2 * sidx;
        gl[0] = from_acc<T>((float)ge[j].W * a[j] * s_x);
        gl[1] = from_acc<T>((float)Hs[j] * a[j] * s_y);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// VECTOR kernels, "sample-geometry" variant (default)
// ------------------------------------------------------------------------------------------------
// In the kernels above all LPR lanes of a group redo the coordinate -> tap-address arithmetic of their
// sample (~55 instructions), so a (b,q,m) unit with 16 samples costs ~600 warp instructions and the SMs are
// issue-bound for a good part of the run.  Here lane i of the warp does that arithmetic ONCE for sample i
// (one coalesced read of the unit's 128-byte location block, 64-byte weight block), and the lane groups
// fetch {tap offset, row stride | validity bits, 4 weights} of the sample they gather with warp shuffles.
// Offsets are 32-bit (host checks S*M*D <= 2^27).  Grid: x = units of one image, y = image.
struct SampleGeo {
  int off00;   // element offset of tap (y0, x0) from the image base (head/channel offset NOT included)
  int rsf;     // (W * M * D) << 4 | ok11 << 3 | ok10 << 2 | ok01 << 1 | ok00
};

template <typename T>
__device__ __forceinline__ void sample_geometry(const T* __restrict__ u_loc, const T* __restrict__ u_att,
                                                const int32_t* __restrict__ shapes, const int32_t* __restrict__ start,
                                                int s, bool have, float inv_p, int MD, SampleGeo& sg, Geo<float>& ge,
                                                float& a, int& H, int& W) {
  const int si = have ? s : 0;
  const int l = have ? level_of(si, inv_p) : 0;
  float lx, ly;
  load_xy(u_loc + 2 * si, lx, ly);
  a = load_s(u_att + si);
  H = __ldg(shapes + 2 * l);
  W = __ldg(shapes + 2 * l + 1);
  const int st = __ldg(start + l);
  ge = make_geo<float>(lx, ly, H, W, have);
  sg.off00 = (st + ge.row00) * MD;
  sg.rsf = ((W * MD) << 4) | (ge.ok11 ? 8 : 0) | (ge.ok10 ? 4 : 0) | (ge.ok01 ? 2 : 0) | (ge.ok00 ? 1 : 0);
}

template <typename T, int D, int U>
__global__ void __launch_bounds__(MSDA_MAX_THREADS)
msda_fwd_sg_kernel(const T* __restrict__ value, const int32_t* __restrict__ shapes,
                   const int32_t* __restrict__ start, const T* __restrict__ loc,
                   const T* __restrict__ attn, T* __restrict__ out,
                   int S, int M, int L, int P, float inv_p, int QM) {
  constexpr int VEC = Vec16<T>::N;
  constexpr int LPR = D / VEC;
  constexpr int G = 32 / LPR;
  static_assert(D % VEC == 0 && LPR >= 1 && LPR <= 32 && (LPR & (LPR - 1)) == 0, "unsupported D for the vector path");

  const int lane = threadIdx.x & 31;
  const int uq = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // unit inside image blockIdx.y
  if (uq >= QM) return;                                                // warp-uniform
  const int g = lane / LPR, cl = lane % LPR;
  const int m = uq % M;
  const int LP = L * P;
  const int MD = M * D;
  const long long u = (long long)blockIdx.y * QM + uq;
  const T* __restrict__ u_loc = loc + u * LP * 2;
  const T* __restrict__ u_att = attn + u * LP;
  const T* __restrict__ vb = value + (long long)blockIdx.y * S * MD + m * D + cl * VEC;
  const T* zp = reinterpret_cast<const T*>(g_zero_line);

  float acc[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] = 0.f;

  for (int base = 0; base < LP; base += 32) {  // 32 samples per pass, one per lane
    SampleGeo sg;
    Geo<float> ge;
    float a;
    int H, W;
    sample_geometry<T>(u_loc, u_att, shapes, start, base + lane, base + lane < LP, inv_p, MD, sg, ge, a, H, W);
    const float w00 = ge.hy * ge.hx * a, w01 = ge.hy * ge.lx * a, w10 = ge.ly * ge.hx * a, w11 = ge.ly * ge.lx * a;
    const int cnt = min(32, LP - base);
    for (int k0 = 0; k0 < cnt; k0 += G * U) {  // warp-uniform trip count
      const T* tp[U][4];
      float w[U][4];
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const int src = k0 + j * G + g;
        const int off = __shfl_sync(0xffffffffu, sg.off00, src);
        int rsf = __shfl_sync(0xffffffffu, sg.rsf, src);
        w[j][0] = __shfl_sync(0xffffffffu, w00, src);
        w[j][1] = __shfl_sync(0xffffffffu, w01, src);
        w[j][2] = __shfl_sync(0xffffffffu, w10, src);
        w[j][3] = __shfl_sync(0xffffffffu, w11, src);
        if (src >= cnt) rsf = 0;  // shfl wraps modulo 32: a lane group past the end must not gather
        const T* t0 = vb + off;
        const int rs = rsf >> 4;
        tp[j][0] = (rsf & 1) ? t0 : zp;
        tp[j][1] = (rsf & 2) ? t0 + MD : zp;
        tp[j][2] = (rsf & 4) ? t0 + rs : zp;
        tp[j][3] = (rsf & 8) ? t0 + rs + MD : zp;
      }
      uint4 v[U][4];
#pragma unroll
      for (int j = 0; j < U; ++j)
#pragma unroll
        for (int t = 0; t < 4; ++t) v[j][t] = ldg128(tp[j][t]);
#pragma unroll
      for (int j = 0; j < U; ++j)
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          float f[VEC];
          Vec16<T>::unpack(v[j][t], f);
#pragma unroll
          for (int i = 0; i < VEC; ++i) acc[i] = fmaf(w[j][t], f[i], acc[i]);
        }
    }
  }
#pragma unroll
  for (int off = LPR; off < 32; off <<= 1)
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], off);
  if (g == 0) *reinterpret_cast<uint4*>(out + u * D + cl * VEC) = Vec16<T>::pack(acc);
}

template <typename T, int D, int U>
__global__ void __launch_bounds__(MSDA_MAX_THREADS)
msda_bwd_sg_kernel(const T* __restrict__ grad_out, const T* __restrict__ value,
                   const int32_t* __restrict__ shapes, const int32_t* __restrict__ start,
                   const T* __restrict__ loc, const T* __restrict__ attn, float* __restrict__ gv,
                   T* __restrict__ gloc, T* __restrict__ gattn,
                   int S, int M, int L, int P, float inv_p, int QM) {
  constexpr int VEC = Vec16<T>::N;
  constexpr int LPR = D / VEC;
  constexpr int G = 32 / LPR;
  static_assert(D % VEC == 0 && LPR >= 1 && LPR <= 32 && (LPR & (LPR - 1)) == 0, "unsupported D for the vector path");

  const int lane = threadIdx.x & 31;
  const int uq = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (uq >= QM) return;
  const int g = lane / LPR, cl = lane % LPR;
  const int m = uq % M;
  const int LP = L * P;
  const int MD = M * D;
  const long long u = (long long)blockIdx.y * QM + uq;
  const T* __restrict__ u_loc = loc + u * LP * 2;
  const T* __restrict__ u_att = attn + u * LP;
  const long long voff = (long long)blockIdx.y * S * MD + m * D + cl * VEC;
  const T* __restrict__ vb = value + voff;
  float* __restrict__ gb = gv + voff;
  const T* zp = reinterpret_cast<const T*>(g_zero_line);

  float go[VEC];
  Vec16<T>::unpack(ldg128(grad_out + u * D + cl * VEC), go);

  for (int base = 0; base < LP; base += 32) {
    SampleGeo sg;
    Geo<float> ge;
    float a;
    int H, W;
    const bool have = base + lane < LP;
    sample_geometry<T>(u_loc, u_att, shapes, start, base + lane, have, inv_p, MD, sg, ge, a, H, W);
    const int cnt = min(32, LP - base);
    float r_a = 0.f, r_x = 0.f, r_y = 0.f;  // channel sums of MY sample, collected from the group that gathered it
    for (int k0 = 0; k0 < cnt; k0 += G * U) {
      const T* tp[U][4];
      int off[U], rsf[U];
      float fly[U], flx[U], fa[U];
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const int src = k0 + j * G + g;
        off[j] = __shfl_sync(0xffffffffu, sg.off00, src);
        rsf[j] = __shfl_sync(0xffffffffu, sg.rsf, src);
        fly[j] = __shfl_sync(0xffffffffu, ge.ly, src);
        flx[j] = __shfl_sync(0xffffffffu, ge.lx, src);
        fa[j] = __shfl_sync(0xffffffffu, a, src);
        if (src >= cnt) rsf[j] = 0;
        const T* t0 = vb + off[j];
        const int rs = rsf[j] >> 4;
        tp[j][0] = (rsf[j] & 1) ? t0 : zp;
        tp[j][1] = (rsf[j] & 2) ? t0 + MD : zp;
        tp[j][2] = (rsf[j] & 4) ? t0 + rs : zp;
        tp[j][3] = (rsf[j] & 8) ? t0 + rs + MD : zp;
      }
      uint4 v[U][4];
#pragma unroll
      for (int j = 0; j < U; ++j)
#pragma unroll
        for (int t = 0; t < 4; ++t) v[j][t] = ldg128(tp[j][t]);
#pragma unroll
      for (int j = 0; j < U; ++j) {
        float f00[VEC], f01[VEC], f10[VEC], f11[VEC];
        Vec16<T>::unpack(v[j][0], f00);
        Vec16<T>::unpack(v[j][1], f01);
        Vec16<T>::unpack(v[j][2], f10);
        Vec16<T>::unpack(v[j][3], f11);
        const float ly_ = fly[j], lx_ = flx[j], hy = 1.f - ly_, hx = 1.f - lx_;
        float s_a = 0.f, s_x = 0.f, s_y = 0.f;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          const float top = hx * f00[i] + lx_ * f01[i];  // interpolated along x on row y0
          const float bot = hx * f10[i] + lx_ * f11[i];  // ... on row y0+1
          s_a = fmaf(go[i], hy * top + ly_ * bot, s_a);
          s_y = fmaf(go[i], bot - top, s_y);
          s_x = fmaf(go[i], hy * (f01[i] - f00[i]) + ly_ * (f11[i] - f10[i]), s_x);
        }
        {  // scatter into grad_value: bilinear weight * attention * grad_out
          float* g0 = gb + off[j];
          const int rs = rsf[j] >> 4;
          const float w00 = hy * hx * fa[j], w01 = hy * lx_ * fa[j], w10 = ly_ * hx * fa[j], w11 = ly_ * lx_ * fa[j];
#pragma unroll
          for (int i = 0; i < VEC; i += 4) {
            if (rsf[j] & 1) red_add_v4(g0 + i, w00 * go[i], w00 * go[i + 1], w00 * go[i + 2], w00 * go[i + 3]);
            if (rsf[j] & 2) red_add_v4(g0 + MD + i, w01 * go[i], w01 * go[i + 1], w01 * go[i + 2], w01 * go[i + 3]);
            if (rsf[j] & 4) red_add_v4(g0 + rs + i, w10 * go[i], w10 * go[i + 1], w10 * go[i + 2], w10 * go[i + 3]);
            if (rsf[j] & 8) red_add_v4(g0 + rs + MD + i, w11 * go[i], w11 * go[i + 1], w11 * go[i + 2], w11 * go[i + 3]);
          }
        }
#pragma unroll
        for (int o = 1; o < LPR; o <<= 1) {  // channel sums over the LPR lanes of this group
          s_a += __shfl_xor_sync(0xffffffffu, s_a, o);
          s_x += __shfl_xor_sync(0xffffffffu, s_x, o);
          s_y += __shfl_xor_sync(0xffffffffu, s_y, o);
        }
        // hand the sums back to the lane that owns the sample: lane (k0 + j*G + g') reads from group g'
        const int rel = lane - k0 - j * G;
        const int from = (rel & (G - 1)) * LPR;
        const float t_a = __shfl_sync(0xffffffffu, s_a, from);
        const float t_x = __shfl_sync(0xffffffffu, s_x, from);
        const float t_y = __shfl_sync(0xffffffffu, s_y, from);
        if (rel >= 0 && rel < G) { r_a = t_a; r_x = t_x; r_y = t_y; }
      }
    }
    if (have) {  // coalesced: 32 consecutive samples of the unit
      const long long sidx = u * LP + base + lane;
      gattn[sidx] = from_acc<T>(r_a);
      store_xy(gloc + 2 * sidx, (float)W * a * r_x, (float)H * a * r_y);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// GENERIC kernels: any D, L, P; T in {float, double, bf16, half}.  Warp per unit, lanes over channels.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(MSDA_MAX_THREADS)
msda_fwd_generic_kernel(const T* __restrict__ value, const int32_t* __restrict__ shapes,
                        const int32_t* __restrict__ start, const T* __restrict__ loc,
                        const T* __restrict__ attn, T* __restrict__ out,
                        int S, int M, int D, int L, int Lq, int P, long long units) {
  using A = typename AccOf<T>::type;
  const int lane = threadIdx.x & 31;
  const long long u = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (u >= units) return;
  const int m = (int)(u % M);
  const long long b = (u / M) / Lq;
  const int LP = L * P;
  const long long MD = (long long)M * D;
  const T* __restrict__ u_loc = loc + u * LP * 2;
  const T* __restrict__ u_att = attn + u * LP;
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
        const A a = to_acc(u_att[s]);
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

template <typename T>
__global__ void __launch_bounds__(MSDA_MAX_THREADS)
msda_bwd_generic_kernel(const T* __restrict__ grad_out, const T* __restrict__ value,
                        const int32_t* __restrict__ shapes, const int32_t* __restrict__ start,
                        const T* __restrict__ loc, const T* __restrict__ attn,
                        typename AccOf<T>::type* __restrict__ gv, T* __restrict__ gloc, T* __restrict__ gattn,
                        int S, int M, int D, int L, int Lq, int P, long long units) {
  using A = typename AccOf<T>::type;
  const int lane = threadIdx.x & 31;
  const long long u = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (u >= units) return;
  const int m = (int)(u % M);
  const long long b = (u / M) / Lq;
  const int LP = L * P;
  const long long MD = (long long)M * D;
  const T* __restrict__ u_loc = loc + u * LP * 2;
  const T* __restrict__ u_att = attn + u * LP;
  const T* __restrict__ u_go = grad_out + u * D;
  const long long voff = b * (long long)S * MD + (long long)m * D;

  for (int l = 0; l < L; ++l) {
    const int H = __ldg(shapes + 2 * l), W = __ldg(shapes + 2 * l + 1);
    const long long loff = voff + (long long)__ldg(start + l) * MD;
    for (int p = 0; p < P; ++p) {
      const int s = l * P + p;
      const Geo<A> ge = make_geo<A>(to_acc(u_loc[2 * s]), to_acc(u_loc[2 * s + 1]), H, W, true);
      const A a = to_acc(u_att[s]);
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
        gattn[sidx] = from_acc<T>(s_a);
        gloc[2 * sidx] = from_acc<T>((A)W * a * s_x);
        gloc[2 * sidx + 1] = from_acc<T>((A)H * a * s_y);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// helpers: zero fill (128-bit stores, grid-stride) and fp32 -> 16-bit conversion of grad_value
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) msda_zero_kernel(uint4* __restrict__ p, long long n16, unsigned char* __restrict__ tail, int ntail) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) p[i] = make_uint4(0, 0, 0, 0);
  if (blockIdx.x == 0 && (int)threadIdx.x < ntail) tail[threadIdx.x] = 0;
}

template <typename T>
__global__ void __launch_bounds__(256) msda_cvt_kernel(const float* __restrict__ src, T* __restrict__ dst, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = from_acc<T>(src[i]);
}

}  // namespace msda
