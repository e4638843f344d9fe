// sortv_capi.cu -- sm_100a `sort_vertices` (include/sortv_b200.h): order the vertices of a convex intersection polygon.
//
// Reference: aloscene/utils/rotated_iou/cuda_op/sort_vert_kernel.cu:42-134 -- one CTA per BATCH ELEMENT, threads stride over
// its polygons, every comparison re-reads vertices / mask / idx from global memory (m * num_valid * 3 loads per polygon).
// Here: one thread per polygon over the flattened (b * n) range (a b = 1 call still fills the GPU), the polygon's m vertices
// and its mask are read ONCE with 128-bit / 64-bit loads into registers, the selection runs entirely in registers (the order
// is kept as 5-bit fields of one 64-bit word, the previously selected vertex is carried along instead of being re-read through
// its index), and the 9 indices leave through shared memory as coalesced stores.  Integer result: bit-exact by construction --
// the comparator below performs the reference's operations in the reference's types.
#include "../../include/sortv_b200.h"

#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

namespace {

thread_local char g_err[256] = "";
std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_variant{0};  // 0 = default (balanced tile kernel for m = 24), 1 = register kernels, 2 = generic kernel, 3 = unbalanced tile kernel, 4 = balanced tile kernel without the sorted-order fast path

int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

constexpr int kIdx = SORTV_MAX_NUM_VERT_IDX;  // 9
constexpr int kOff = 8;                       // first intersection candidate (sort_vert_kernel.cu:7)
#define SORTV_EPS 1e-8                        // a DOUBLE constant, as in the reference (sort_vert_kernel.cu:8)

// "vertex 1 comes before vertex 2": smallest angle first, counter-clockwise (sort_vert_kernel.cu:16-40).  Same operations
// in the same types: float differences compared against the double epsilon; squared norms as float fma + double add rounded
// back to float; IEEE float division.  The reference has no return statement for the remaining case (a y that is exactly 0
// without opposite signs); its sm_100a build returns false there (checked in its PTX and against the compiled reference).
__device__ __forceinline__ bool before(float x1, float y1, float x2, float y2) {
  if ((double)fabsf(x1 - x2) < SORTV_EPS && (double)fabsf(y2 - y1) < SORTV_EPS) return false;
  if (y1 > 0.f && y2 < 0.f) return true;
  if (y1 < 0.f && y2 > 0.f) return false;
  const float n1 = (float)((double)fmaf(x1, x1, y1 * y1) + SORTV_EPS);
  const float n2 = (float)((double)fmaf(x2, x2, y2 * y2) + SORTV_EPS);
  const float d = __fsub_rn(__fdiv_rn(fabsf(x1) * x1, n1), __fdiv_rn(fabsf(x2) * x2, n2));
  if (y1 > 0.f && y2 > 0.f) return (double)d > SORTV_EPS;
  if (y1 < 0.f && y2 < 0.f) return (double)d < SORTV_EPS;
  return false;
}

// M = candidates per polygon (compile time, <= 32): everything in registers.
template <int M>
__global__ void __launch_bounds__(128) sortv_kernel(const float* __restrict__ vertices, const uint8_t* __restrict__ mask,
                                                    const int32_t* __restrict__ num_valid, int32_t* __restrict__ idx, long long total) {
  __shared__ int32_t s_idx[128 * kIdx];
  const long long p0 = (long long)blockIdx.x * blockDim.x;
  const long long p = p0 + threadIdx.x;
  if (p < total) {
    float vx[M], vy[M];
    unsigned valid = 0;
    {
      const float* v = vertices + p * (2 * M);
      if constexpr ((2 * M) % 4 == 0) {  // polygon = 8 * M bytes: 16-byte aligned when the tensor is
#pragma unroll
        for (int k = 0; k < M / 2; ++k) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(v) + k);
          vx[2 * k] = t.x; vy[2 * k] = t.y; vx[2 * k + 1] = t.z; vy[2 * k + 1] = t.w;
        }
      } else {
#pragma unroll
        for (int k = 0; k < M; ++k) {
          const float2 t = __ldg(reinterpret_cast<const float2*>(v) + k);
          vx[k] = t.x; vy[k] = t.y;
        }
      }
      const uint8_t* mk = mask + p * M;
      if constexpr (M % 8 == 0) {
#pragma unroll
        for (int k = 0; k < M / 8; ++k) {
          const unsigned long long w = __ldg(reinterpret_cast<const unsigned long long*>(mk) + k);
#pragma unroll
          for (int t = 0; t < 8; ++t) valid |= (((w >> (8 * t)) & 0xffull) != 0 ? 1u : 0u) << (8 * k + t);
        }
      } else {
#pragma unroll
        for (int k = 0; k < M; ++k) valid |= (__ldg(mk + k) != 0 ? 1u : 0u) << k;
      }
    }
    // index of an invalid intersection candidate (sort_vert_kernel.cu:56-62)
    const unsigned inv = ~valid & (M < 32 ? ((1u << M) - 1u) : 0xffffffffu) & ~((1u << kOff) - 1u);
    const int pad = inv ? (__ffs(inv) - 1) : M - 1;
    const int nv_in = __ldg(num_valid + p);
    unsigned long long order = 0;  // idx[j] in bits [5j, 5j+5)
    auto put = [&](int j, int v) { order = (order & ~(31ull << (5 * j))) | ((unsigned long long)v << (5 * j)); };
    auto get = [&](int j) { return (int)((order >> (5 * j)) & 31ull); };
    if (nv_in < 3) {
#pragma unroll
      for (int j = 0; j < kIdx; ++j) put(j, pad);
    } else {
      const int nv = nv_in > kIdx - 1 ? kIdx - 1 : nv_in;
      float px = 0.f, py = 0.f;  // vertex selected in the previous round (the reference re-reads it through idx[j-1])
      for (int j = 0; j < nv; ++j) {
        float x_min = 1.f, y_min = (float)(-SORTV_EPS);  // "big" start value (sort_vert_kernel.cu:77-79)
        int take = 0;
#pragma unroll
        for (int k = 0; k < M; ++k) {
          if (!((valid >> k) & 1u)) continue;
          const float x = vx[k], y = vy[k];
          if (before(x, y, x_min, y_min) && (j == 0 || before(px, py, x, y))) {
            x_min = x; y_min = y; take = k;
          }
        }
        put(j, take);
        // idx[j] = take; the next round compares against vertices[idx[j]] -- vertex 0 when nothing was selected
        float sx = vx[0], sy = vy[0];
#pragma unroll
        for (int k = 1; k < M; ++k)
          if (take == k) { sx = vx[k]; sy = vy[k]; }
        px = sx; py = sy;
      }
      put(nv, get(0));  // close the polygon (sort_vert_kernel.cu:104)
      for (int j = nv + 1; j < kIdx; ++j) put(j, pad);
      // the two boxes are identical: corners appear twice among the first 8 (sort_vert_kernel.cu:111-131)
      if (nv_in == 8) {
        int counter = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int k = 4; k < kOff; ++k) counter += get(k) == get(j) ? 1 : 0;
        if (counter == 4) {
          put(4, get(0));
#pragma unroll
          for (int j = 5; j < kIdx; ++j) put(j, pad);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kIdx; ++j) s_idx[threadIdx.x * kIdx + j] = get(j);
  }
  __syncthreads();
  // coalesced write of this CTA's (<= 128 * 9) contiguous indices
  const long long rows = total - p0 < (long long)blockDim.x ? total - p0 : (long long)blockDim.x;
  int32_t* dst = idx + p0 * kIdx;
  for (int i = threadIdx.x; i < (int)rows * kIdx; i += blockDim.x) dst[i] = s_idx[i];
}

// ---------------------------------------------------------------------------------------------------------------------
// Tile kernel (m = 24, the reference's only shape; 16-byte aligned tensors): the default path.
//
//  * a CTA owns 128 consecutive polygons = ONE contiguous 24 KB range of `vertices`, 3 KB of `mask`, 512 B of `num_valid`:
//    one elected thread brings the three ranges into shared memory with three 1-D TMA bulk copies (cp.async.bulk) completing
//    on one mbarrier -- every HBM sector is read once, fully used, with no per-thread strided loads;
//  * the comparator's expensive half depends on ONE vertex only: q(v) = |x| * x / (float)(fma(x, x, y * y) + 1e-8) (an IEEE
//    division and a double-precision add).  The reference re-evaluates it for both operands of every comparison
//    (4 * m * num_valid divisions per polygon); here it is evaluated once per VALID vertex while the thread compacts its valid
//    candidates (x, y, q, index) into a private shared-memory column, and `before_q` compares the stored values with the same
//    float subtraction and the same thresholds -- identical results, ~30 x fewer instructions;
//  * the selection rounds walk the c <= 8 compacted candidates instead of all 24;
//  * the double-precision threshold tests on float values are replaced by exactly equivalent float comparisons (below);
//  * the 9 indices of the CTA's polygons leave as one contiguous 4.5 KB range of 128-bit stores.
// A polygon with more than 8 valid candidates (impossible for two rectangles) walks the staged tile directly.
constexpr int kTile = 128;
constexpr int kSlots = 8;

// (double)f < 1e-8 and (double)f > 1e-8 for a float f, without the conversion: 1e-8 is not a float, F is its nearest float.
constexpr float kF = (float)SORTV_EPS;
constexpr bool kFAbove = (double)kF > SORTV_EPS;
__device__ __forceinline__ bool lt_eps(float f) { return kFAbove ? (f < kF) : (f <= kF); }
__device__ __forceinline__ bool gt_eps(float f) { return kFAbove ? (f >= kF) : (f > kF); }

__device__ __forceinline__ float q_of(float x, float y) {
  const float n = (float)((double)fmaf(x, x, y * y) + SORTV_EPS);
  return __fdiv_rn(fabsf(x) * x, n);
}

// before(x1, y1, x2, y2) with q1 = q_of(x1, y1), q2 = q_of(x2, y2) supplied.  Branch-free: the reference's chain of early
// returns picks ONE of four mutually exclusive sign cases (+-: true, -+: false, ++: d > eps, --: d < eps; a zero or NaN y:
// false) unless the two vertices coincide, which is what the predicate expression below evaluates.
__device__ __forceinline__ bool before_q(float x1, float y1, float q1, float x2, float y2, float q2) {
  const bool tie = lt_eps(fabsf(x1 - x2)) & lt_eps(fabsf(y2 - y1));
  const bool p1 = y1 > 0.f, n1 = y1 < 0.f, p2 = y2 > 0.f, n2 = y2 < 0.f;
  const float d = __fsub_rn(q1, q2);
  return !tie & ((p1 & n2) | (p1 & p2 & gt_eps(d)) | (n1 & n2 & lt_eps(d)));
}

__device__ __forceinline__ uint32_t sv_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// BAL: in-CTA load balancing.  The selection costs c * num_valid comparator pairs and c varies from polygon to polygon (0 for
// disjoint boxes ... 8), so a warp whose lanes own consecutive polygons idles most lanes while the largest polygon finishes
// (measured: 9.7 of 32 lanes active).  With BAL the CTA counting-sorts its 128 polygons by c (shared-memory atomics, 9
// buckets; c is known from the mask before anything is compacted), every owner writes its compacted candidates into the
// column of its SORTED position, thread t runs the selection of column t (conflict-free 128-bit reads, near-uniform trip
// counts inside a warp) and hands the packed order word back to the owner through shared memory.
// Shared memory: the compacted columns take the place of the staged vertices (the owner carries its <= 8 candidates in
// registers across the barrier that retires the tile), so a CTA needs 29.7 KB and 7 CTAs = 28 warps fit an SM.
template <bool BAL>
__global__ void __launch_bounds__(kTile) sortv_tile_kernel(const float* __restrict__ vertices, const uint8_t* __restrict__ mask,
                                                           const int32_t* __restrict__ num_valid, int32_t* __restrict__ idx,
                                                           long long total, bool fast_order) {
  constexpr int M = 24;
  // 24 576 B: staged vertices; then the columns s_slot[kSlots + 1][kTile] (18 432 B; column p = candidates (x, y, q, index)
  // of the polygon at sorted position p, row kSlots = (vertex 0, its q, c | rounds << 8)); then the indices (4 608 B)
  __shared__ __align__(128) float2 s_v[kTile * M];
  __shared__ __align__(16) uint8_t s_m[kTile * M];  //  3 072 B
  __shared__ __align__(16) int32_t s_nv[kTile];     //    512 B
  __shared__ unsigned long long s_order[kTile];     //  1 024 B: packed result, indexed by owner
  __shared__ int s_owner[kTile];                    //    512 B: owner thread of the polygon at sorted position p (-1: none)
  __shared__ int s_cnt[kSlots + 1];
  __shared__ __align__(8) unsigned long long s_bar;
  float4(*s_slot)[kTile] = reinterpret_cast<float4(*)[kTile]>(s_v);
  const int t = threadIdx.x;
  const long long p0 = (long long)blockIdx.x * kTile;
  const int rows = total - p0 < (long long)kTile ? (int)(total - p0) : kTile;

  if (BAL && t <= kSlots) s_cnt[t] = 0;
  if (rows == kTile) {
    if (t == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sv_smem_u32(&s_bar)) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      constexpr uint32_t bv = kTile * M * 8, bm = kTile * M, bn = kTile * 4;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sv_smem_u32(&s_bar)), "r"(bv + bm + bn) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(sv_smem_u32(s_v)), "l"(vertices + p0 * (2 * M)), "r"(bv), "r"(sv_smem_u32(&s_bar)) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(sv_smem_u32(s_m)), "l"(mask + p0 * M), "r"(bm), "r"(sv_smem_u32(&s_bar)) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(sv_smem_u32(s_nv)), "l"(num_valid + p0), "r"(bn), "r"(sv_smem_u32(&s_bar)) : "memory");
    }
    __syncthreads();  // the barrier is initialised before anybody polls it
    uint32_t done = 0;
    const uint32_t a = sv_smem_u32(&s_bar);
    while (!done)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(a) : "memory");
  } else {  // tail tile: byte counts are not multiples of 16
    const float2* gv = reinterpret_cast<const float2*>(vertices) + p0 * M;
    for (int i = t; i < rows * M; i += kTile) {
      s_v[i] = __ldg(gv + i);
      s_m[i] = __ldg(mask + p0 * M + i);
    }
    if (t < rows) s_nv[t] = __ldg(num_valid + p0 + t);
    __syncthreads();
  }

  // ---- phase A (owner thread): mask bits, pad index, candidates into registers; polygons the fast path cannot take are finished
  int pad = 0, nv_in = 0, nv = 0, c = 0;
  unsigned long long order = 0;  // idx[j] in bits [5j, 5j + 5)
  bool fast = false;             // selection still to be run over the compacted column
  const float y0 = -kF;          // (float)(-EPSILON), sort_vert_kernel.cu:78-79
  const float q0 = q_of(1.f, y0);
  float4 cand[kSlots];
  float4 head = make_float4(0.f, 0.f, 0.f, 0.f);
  if (t < rows) {
    unsigned valid = 0;
    {
      const uint2* mw = reinterpret_cast<const uint2*>(s_m + t * M);  // 24 bytes, 8-byte aligned
#pragma unroll
      for (int k = 0; k < M / 8; ++k) {
        const uint2 w = mw[k];
        // byte != 0 -> bit: OR the byte's bits down to its bit 0, then gather the four bit-0s with one multiply
        unsigned lo = w.x | (w.x >> 4); lo |= lo >> 2; lo |= lo >> 1; lo &= 0x01010101u;
        unsigned hi = w.y | (w.y >> 4); hi |= hi >> 2; hi |= hi >> 1; hi &= 0x01010101u;
        const unsigned b4lo = (lo * 0x01020408u) >> 24 & 0xfu;  // bytes 0..3 -> bits 0..3
        const unsigned b4hi = (hi * 0x01020408u) >> 24 & 0xfu;
        valid |= (b4lo | (b4hi << 4)) << (8 * k);
      }
    }
    const unsigned inv = ~valid & 0x00ffffffu & ~((1u << kOff) - 1u);
    pad = inv ? (__ffs(inv) - 1) : M - 1;  // sort_vert_kernel.cu:56-62
    nv_in = s_nv[t];
    nv = nv_in > kIdx - 1 ? kIdx - 1 : nv_in;
    c = __popc(valid);
    const float2* v = s_v + t * M;
    if (nv_in >= 3 && c <= kSlots) {
      fast = true;
      unsigned bits = valid;
#pragma unroll
      for (int i = 0; i < kSlots; ++i) {
        if (i < c) {
          const int k = __ffs(bits) - 1;
          bits &= bits - 1;
          const float2 xy = v[k];
          cand[i] = make_float4(xy.x, xy.y, q_of(xy.x, xy.y), __int_as_float(k));
        }
      }
      const float2 v0 = v[0];
      head = make_float4(v0.x, v0.y, q_of(v0.x, v0.y), __int_as_float(c | (nv << 8)));
    } else if (nv_in >= 3) {  // more than 8 valid candidates: same scan over the staged tile
      float px = 0.f, py = 0.f, pq = 0.f;
      for (int j = 0; j < nv; ++j) {
        float x_min = 1.f, y_min = y0, q_min = q0;
        int take = -1;
        unsigned bits = valid;
        while (bits) {
          const int k = __ffs(bits) - 1;
          bits &= bits - 1;
          const float2 xy = v[k];
          const float q = q_of(xy.x, xy.y);
          if (before_q(xy.x, xy.y, q, x_min, y_min, q_min) && (j == 0 || before_q(px, py, pq, xy.x, xy.y, q))) {
            x_min = xy.x; y_min = xy.y; q_min = q; take = k;
          }
        }
        if (take < 0) {  // nothing selected: idx[j] = 0 and the next round compares against vertex 0
          take = 0;
          const float2 xy = v[0];
          px = xy.x; py = xy.y; pq = q_of(px, py);
        } else {
          px = x_min; py = y_min; pq = q_min;
        }
        order |= (unsigned long long)take << (5 * j);
      }
    }
  }

  // ---- sorted position of this thread's polygon; the staged tile retires, the columns take its place
  int position = t;
  if constexpr (BAL) {
    const int key = fast ? c : 0;  // bucket 0 also holds the polygons with nothing left to do
    const int pos = atomicAdd(&s_cnt[key], 1);
    __syncthreads();
    int base = 0;
#pragma unroll
    for (int k = 0; k < kSlots; ++k) base += k < key ? s_cnt[k] : 0;
    position = base + pos;
  } else {
    __syncthreads();
  }
  if (fast) {
#pragma unroll
    for (int i = 0; i < kSlots; ++i)
      if (i < c) s_slot[i][position] = cand[i];
    s_slot[kSlots][position] = head;
  }
  s_owner[position] = fast ? t : -1;
  __syncthreads();

  // ---- phase B: the order of column t.
  // Fast path: when `before` restricted to this polygon's candidates is a strict total order -- irreflexive (it is not for
  // y = -inf: inf - inf defeats the coincidence test), every pair ordered one way exactly, the in-degrees a permutation of
  // 0..c-1 (a tournament with that score sequence is transitive) -- every candidate comes before the start value and there
  // is one round per candidate, the reference's rounds ("smallest
  // candidate after the previous one", each a min-scan under that order) return the candidates in sorted order: idx[r] =
  // the candidate with in-degree r.  That takes c (c - 1) / 2 joint evaluations of before(a, b) / before(b, a) -- they
  // share the coincidence test, the sign classes and d = q_a - q_b (q_b - q_a = -d exactly) -- instead of 2 c^2 single ones.
  // Anything else (coincident or zero-y vertices, near-equal quotients that break antisymmetry, more rounds than
  // candidates) runs the rounds themselves.  Same comparator values either way, so the indices are identical.
  {
    const int owner = s_owner[t];
    if (owner >= 0) {
      const float4 h = s_slot[kSlots][t];
      const int cc = __float_as_int(h.w) & 0xff, rounds = __float_as_int(h.w) >> 8;
      unsigned long long ord = 0;
      bool total = rounds == cc && fast_order;
      {
        unsigned rankword = 0, seen = 0, onehot_a = 1u;  // in-degree of candidate a in bits [4a, 4a + 4)
        for (int a = 0; a < cc; ++a, onehot_a <<= 4) {
          const float4 A = s_slot[a][t];
          total &= before_q(A.x, A.y, A.z, 1.f, y0, q0) & !before_q(A.x, A.y, A.z, A.x, A.y, A.z);
          const bool pa = A.y > 0.f, na = A.y < 0.f;
          unsigned onehot_b = onehot_a << 4;
          for (int b = a + 1; b < cc; ++b, onehot_b <<= 4) {
            const float4 B = s_slot[b][t];
            const bool tie = lt_eps(fabsf(A.x - B.x)) & lt_eps(fabsf(B.y - A.y));
            const bool pb = B.y > 0.f, nb = B.y < 0.f;
            const float d = __fsub_rn(A.z, B.z), nd = -d;
            const bool ab = !tie & ((pa & nb) | (pa & pb & gt_eps(d)) | (na & nb & lt_eps(d)));
            const bool ba = !tie & ((pb & na) | (pb & pa & gt_eps(nd)) | (nb & na & lt_eps(nd)));
            total &= ab != ba;
            rankword += ab ? onehot_b : 0u;
            rankword += ba ? onehot_a : 0u;
          }
          const unsigned r = (rankword >> (4 * a)) & 15u;  // final: every pair with a has been seen
          seen |= 1u << r;
          ord |= (unsigned long long)__float_as_int(A.w) << (5 * r);
        }
        total &= seen == (1u << cc) - 1u;
      }
      if (!total) {
        ord = 0;
        float px = 0.f, py = 0.f, pq = 0.f;
        for (int j = 0; j < rounds; ++j) {
          float x_min = 1.f, y_min = y0, q_min = q0;
          int take = -1;
          for (int i = 0; i < cc; ++i) {
            const float4 s = s_slot[i][t];
            const bool sel = before_q(s.x, s.y, s.z, x_min, y_min, q_min) & ((j == 0) | before_q(px, py, pq, s.x, s.y, s.z));
            x_min = sel ? s.x : x_min; y_min = sel ? s.y : y_min; q_min = sel ? s.z : q_min;
            take = sel ? __float_as_int(s.w) : take;
          }
          // nothing selected: idx[j] = 0 and the next round compares against vertex 0
          const bool none = take < 0;
          px = none ? h.x : x_min; py = none ? h.y : y_min; pq = none ? h.z : q_min;
          ord |= (unsigned long long)(none ? 0 : take) << (5 * j);
        }
      }
      s_order[owner] = ord;
    }
  }
  __syncthreads();  // results published; the columns retire, the indices take their place
  if (fast) order = s_order[t];

  // ---- phase C (owner thread): close the polygon, pad, identical-boxes corner case; coalesced store
  int32_t* s_idx = reinterpret_cast<int32_t*>(s_v);
  if (t < rows) {
    int o[kIdx];
    if (nv_in < 3) {
#pragma unroll
      for (int j = 0; j < kIdx; ++j) o[j] = pad;
    } else {
#pragma unroll
      for (int jj = 0; jj < kIdx - 1; ++jj) o[jj] = (int)((order >> (5 * jj)) & 31ull);
      // close the polygon (sort_vert_kernel.cu:104), pad (:107-109)
#pragma unroll
      for (int jj = 1; jj < kIdx; ++jj) o[jj] = jj == nv ? o[0] : (jj > nv ? pad : o[jj]);
      // the two boxes are identical: corners appear twice among the first 8 (sort_vert_kernel.cu:111-131)
      if (nv_in == 8) {
        int counter = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int k = 4; k < kOff; ++k) counter += o[k] == o[j] ? 1 : 0;
        if (counter == 4) {
          o[4] = o[0];
#pragma unroll
          for (int j = 5; j < kIdx; ++j) o[j] = pad;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kIdx; ++j) s_idx[t * kIdx + j] = o[j];
  }
  __syncthreads();
  int32_t* dst = idx + p0 * kIdx;
  if (rows == kTile) {  // 4 608 B, 16-byte aligned
    const int4* src4 = reinterpret_cast<const int4*>(s_idx);
    int4* dst4 = reinterpret_cast<int4*>(dst);
    for (int i = t; i < kTile * kIdx / 4; i += kTile) dst4[i] = src4[i];
  } else {
    for (int i = t; i < rows * kIdx; i += kTile) dst[i] = s_idx[i];
  }
}

// any m (>= 9): local-memory restatement of the same scan, one thread per polygon
__global__ void __launch_bounds__(128) sortv_kernel_generic(const float* __restrict__ vertices, const uint8_t* __restrict__ mask,
                                                            const int32_t* __restrict__ num_valid, int32_t* __restrict__ idx, long long total, int m) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total) return;
  const float* v = vertices + p * 2 * m;
  const uint8_t* mk = mask + p * m;
  int32_t* out = idx + p * kIdx;
  int pad = m - 1;
  for (int j = kOff; j < m; ++j)
    if (!mk[j]) { pad = j; break; }
  const int nv_in = num_valid[p];
  if (nv_in < 3) {
    for (int j = 0; j < kIdx; ++j) out[j] = pad;
    return;
  }
  const int nv = nv_in > kIdx - 1 ? kIdx - 1 : nv_in;
  int o[kIdx];
  for (int j = 0; j < nv; ++j) {
    float x_min = 1.f, y_min = (float)(-SORTV_EPS);
    int take = 0;
    for (int k = 0; k < m; ++k) {
      if (!mk[k]) continue;
      const float x = v[2 * k], y = v[2 * k + 1];
      if (before(x, y, x_min, y_min) && (j == 0 || before(v[2 * o[j - 1]], v[2 * o[j - 1] + 1], x, y))) {
        x_min = x; y_min = y; take = k;
      }
    }
    o[j] = take;
  }
  o[nv] = o[0];
  for (int j = nv + 1; j < kIdx; ++j) o[j] = pad;
  if (nv_in == 8) {
    int counter = 0;
    for (int j = 0; j < 4; ++j)
      for (int k = 4; k < kOff; ++k) counter += o[k] == o[j] ? 1 : 0;
    if (counter == 4) {
      o[4] = o[0];
      for (int j = 5; j < kIdx; ++j) o[j] = pad;
    }
  }
  for (int j = 0; j < kIdx; ++j) out[j] = o[j];
}

}  // namespace

extern "C" {

int sortv_version(void) { return SORTV_ABI_VERSION; }

const char* sortv_last_error_string(void) { return g_err; }

uint64_t sortv_kernel_launch_count(void) { return g_launches.load(); }

int sortv_set_variant(int variant) {
  if (variant < 0 || variant > 4) return fail("sortv_set_variant: unknown variant %d", variant);
  g_variant.store(variant);
  return 0;
}

int sortv_sort_vertices(const float* vertices, const uint8_t* mask, const int32_t* num_valid, int32_t* idx, int b, int n, int m,
                        void* stream) {
  g_err[0] = 0;
  if (b < 0 || n < 0 || m < 0) return fail("negative dimension");
  const long long total = (long long)b * n;
  if (total == 0) return 0;
  if (m < kIdx) return fail("m = %d: need at least %d candidates per polygon (8 box corners + intersections)", m, kIdx);
  if (!vertices || !mask || !num_valid || !idx) return fail("NULL tensor pointer passed to sortv_sort_vertices");
  const long long blocks = (total + 127) / 128;
  if (blocks > 0x7fffffffLL) return fail("problem too large for one launch");
  cudaStream_t st = (cudaStream_t)stream;
  const bool a16 = (reinterpret_cast<uintptr_t>(vertices) % 16) == 0 && (reinterpret_cast<uintptr_t>(mask) % 8) == 0;
  const bool tma_ok = a16 && (reinterpret_cast<uintptr_t>(mask) % 16) == 0 && (reinterpret_cast<uintptr_t>(num_valid) % 16) == 0 &&
                      (reinterpret_cast<uintptr_t>(idx) % 16) == 0;
  if (m == 24 && tma_ok && g_variant == 0) sortv_tile_kernel<true><<<(unsigned)blocks, kTile, 0, st>>>(vertices, mask, num_valid, idx, total, true);
  else if (m == 24 && tma_ok && g_variant == 3) sortv_tile_kernel<false><<<(unsigned)blocks, kTile, 0, st>>>(vertices, mask, num_valid, idx, total, true);
  else if (m == 24 && tma_ok && g_variant == 4) sortv_tile_kernel<true><<<(unsigned)blocks, kTile, 0, st>>>(vertices, mask, num_valid, idx, total, false);
  else if (m == 24 && a16 && g_variant != 2) sortv_kernel<24><<<(unsigned)blocks, 128, 0, st>>>(vertices, mask, num_valid, idx, total);
  else if (m == 16 && a16 && g_variant != 2) sortv_kernel<16><<<(unsigned)blocks, 128, 0, st>>>(vertices, mask, num_valid, idx, total);
  else if (m == 32 && a16 && g_variant != 2) sortv_kernel<32><<<(unsigned)blocks, 128, 0, st>>>(vertices, mask, num_valid, idx, total);
  else sortv_kernel_generic<<<(unsigned)blocks, 128, 0, st>>>(vertices, mask, num_valid, idx, total, m);
  const cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail("sortv_sort_vertices: CUDA launch failed: %s", cudaGetErrorString(e));
  }
  g_launches.fetch_add(1);
  return 0;
}

}  // extern "C"
