// msda_fwd_win.cuh -- WINDOWED forward for pixel-aligned queries (encoder self-attention: Lq == S, query i is pixel i of the
// pyramid and samples around its own position).  BASELINE.json's north star: "per-level value tiles staged into shared memory".
//
// The unit-ordered forward gathers every tap row (128 B) through the L1 miss path: 64 rows per unit, 2.0 clk each, and each row of
// `value` is requested by ~64 different units of the call.  Here a CTA owns one head of a TY x TX tile of queries and
//   A. computes the window geometry of all its samples (same arithmetic as spec_record) and, per level, the bounding box of
//      the 2x2 tap windows of the tile;
//   B. copies those boxes -- the only rows of `value` the tile can touch -- into shared memory once (cp.async, 16 B per lane,
//      all rows in flight together: ONE global round trip per tile instead of one per gather round);
//   C. runs the gather rounds of its units out of shared memory (LDS.128: 1.05 clk per row, ~30 clk latency).
// A level whose box does not fit the budget (a tile of a coarse level looking at a fine one, or non-local sampling locations)
// is not staged: its samples keep their global tap addresses and are gathered as in the unit-ordered kernel, round by round.
// Pure re-scheduling: per unit the FMAs run in the order of msda_fwd_unit's speculative path with the same weights, so the result
// is bit-identical as long as every loaded value is finite; a unit whose sums are not finite is redone on the flagged path.
// One warp serves UPW = 4 queries of the tile, two at a time (lanes 0..15: the samples of one query, lanes 16..31 those of the
// next; needs L * P <= 16), like msda_fwd_pair.
//
// MEASURED (B200, ENC = N 2, 100^2 pyramid, Lq = S, fp32; profiles/r2_ncu_enc_fwd_paired.md): bit-identical, L2 throughput 53 -> 18 %
// of peak, 165 us against 95 us for the unit-ordered kernel (187 us with non-local locations).  Phase isolation: tile loop +
// barriers 29 us, + geometry / boxes 68 us, + copy 87 us, and the shared-memory gather rounds add 80 us by themselves (72 LDS
// wavefronts per unit = 53 us at the pipe's peak).  Opt-in (knob "fwd_win_mode" = 2), not selected automatically.
#pragma once
#include "msda_kernels.cuh"

namespace msda {

#define MSDA_WIN_LEVELS 8

template <int D_, int TXS_, int TYS_, int CAP_>
struct FwdWinCfg {
  static constexpr int D = D_;
  static constexpr int TXS = TXS_, TYS = TYS_;            // log2 of the tile edge in queries
  static constexpr int TQ = 1 << (TXS_ + TYS_);           // queries per tile
  static constexpr int UPW = 4;                           // queries per warp
  static constexpr int WARPS = TQ / UPW;
  static constexpr int THREADS = 32 * WARPS;
  static constexpr int CAP = CAP_;                        // window rows (of D floats) per CTA, + 1 zero row in front
  static constexpr int ROWB = D_ * 4;                     // bytes per row
  static constexpr size_t WIN_BYTES = (size_t)(CAP_ + 1) * ROWB;
  static constexpr size_t REC_BYTES = (size_t)THREADS * 24;
  static constexpr size_t SMEM = WIN_BYTES + REC_BYTES + 16 * MSDA_WIN_LEVELS * 4;
};

__device__ __forceinline__ void win_cp_async16(unsigned dst_smem, const void* src) {  // L2 -> shared memory, no L1 allocation
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void win_cp_async_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ uint4 lds128(unsigned a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}

// Per-level words in shared memory (int each, [MSDA_WIN_LEVELS]):
//   H, W, st, first (index of the level's first query)          -- constant for the kernel
//   xmin, xmax, ymin, ymax                                      -- bounding box of the tile's window origins (x0, y0), per tile
//   wbase (first window row, 1-based: row 0 is the zero row), ww (window width in pixels), staged (0 / 1)
struct WinLevels {
  int H[MSDA_WIN_LEVELS], W[MSDA_WIN_LEVELS], st[MSDA_WIN_LEVELS], first[MSDA_WIN_LEVELS];
  int xmin[MSDA_WIN_LEVELS], xmax[MSDA_WIN_LEVELS], ymin[MSDA_WIN_LEVELS], ymax[MSDA_WIN_LEVELS];
  int wbase[MSDA_WIN_LEVELS], ww[MSDA_WIN_LEVELS], staged[MSDA_WIN_LEVELS], rows[MSDA_WIN_LEVELS];
  int total_rows, t_m, t_b, t_level, t_y, t_x, pad[2];
};

template <typename Cfg, int MC>
__global__ void __launch_bounds__(Cfg::THREADS, 2)
msda_fwd_win_kernel(const float* __restrict__ value, const int32_t* __restrict__ shapes, const int32_t* __restrict__ start,
                    const float* __restrict__ loc, const float* __restrict__ attn, float* __restrict__ out,
                    int N, int S, int Mrt, int L, int P, float inv_p, int QM) {
  using T = float;
  constexpr int D = Cfg::D;
  constexpr int VEC = Vec16<T>::N;       // 4
  constexpr int LPR = D / VEC;           // 8 lanes per row
  constexpr int G = 32 / LPR;            // 4 rows per warp instruction
  constexpr int N_OUT = GroupReduceScatter<VEC, LPR>::N_OUT;
  constexpr int TX = 1 << Cfg::TXS, TY = 1 << Cfg::TYS;
  const int M = MC > 0 ? MC : Mrt;
  const int MD = M * D;
  const int Lq = QM / M;
  const int LP = L * P;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int g = lane / LPR, cl = lane % LPR;
  const int h = lane >> 4, s = lane & 15;   // which query of the pair, which of its samples
  const bool have_s = s < LP;
  const int lvl = have_s ? level_of(s, inv_p) : 0;

  extern __shared__ __align__(128) unsigned char win_smem[];
  float* win = reinterpret_cast<float*>(win_smem);                                     // (CAP + 1) rows
  uint4* rec_a = reinterpret_cast<uint4*>(win_smem + Cfg::WIN_BYTES) + (threadIdx.x & ~31);   // this warp's 32 records
  float2* rec_b = reinterpret_cast<float2*>(win_smem + Cfg::WIN_BYTES + (size_t)Cfg::THREADS * 16) + (threadIdx.x & ~31);
  WinLevels* lv = reinterpret_cast<WinLevels*>(win_smem + Cfg::WIN_BYTES + Cfg::REC_BYTES);
  const unsigned win_s = (unsigned)__cvta_generic_to_shared(win);

  // ---- once per CTA: level table, zero row ----
  int patches = 0, pixels = 0;
  for (int l = 0; l < L; ++l) {
    const int H = __ldg(shapes + 2 * l), W = __ldg(shapes + 2 * l + 1);
    if (threadIdx.x == 0) { lv->H[l] = H; lv->W[l] = W; lv->st[l] = __ldg(start + l); lv->first[l] = pixels; }
    patches += ((H + TY - 1) >> Cfg::TYS) * ((W + TX - 1) >> Cfg::TXS);
    pixels += H * W;
  }
  if (threadIdx.x < D) win[threadIdx.x] = 0.f;
  __syncthreads();
  const bool narrow = !__all_sync(0xffffffffu, !have_s || (lv->H[lvl] >= 2 && lv->W[lvl] >= 2));  // uniform over the CTA
  const int myH = lv->H[lvl], myW = lv->W[lvl], mySt = lv->st[lvl];
  const int NM = N * M;
  const long long n_tiles = (long long)patches * NM;

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    // ---- tile -> (patch, image, head) -> level, origin: one thread decodes, everybody reads ----
    if (threadIdx.x == 0) {
      const unsigned t32 = (unsigned)tile;  // (the host launches this kernel only for < 2^31 tiles)
      const unsigned rest = t32 / (unsigned)M;
      lv->t_m = (int)(t32 - rest * (unsigned)M);
      int patch = (int)(rest / (unsigned)N);
      lv->t_b = (int)(rest - (unsigned)patch * (unsigned)N);
      int tl = 0, npx = 1;
      for (; tl < L; ++tl) {
        npx = (lv->W[tl] + TX - 1) >> Cfg::TXS;
        const int np = ((lv->H[tl] + TY - 1) >> Cfg::TYS) * npx;
        if (patch < np) break;
        patch -= np;
      }
      const int prow = patch / npx;
      lv->t_level = tl;
      lv->t_y = prow << Cfg::TYS;
      lv->t_x = (patch - prow * npx) << Cfg::TXS;
    }
    if (threadIdx.x < MSDA_WIN_LEVELS) {
      lv->xmin[threadIdx.x] = 0x7fffffff; lv->ymin[threadIdx.x] = 0x7fffffff;
      lv->xmax[threadIdx.x] = -1; lv->ymax[threadIdx.x] = -1;
    }
    __syncthreads();
    const int m = lv->t_m, b = lv->t_b;
    const int tH = lv->H[lv->t_level], tW = lv->W[lv->t_level], tfirst = lv->first[lv->t_level];
    const int y_base = lv->t_y, x_base = lv->t_x;

    // ---- A: geometry of my samples (two pairs of queries per warp) ----
    int xy[2];         // x0 | y0 << 16 of the 2x2 window, or -1: no contribution (outside sample / idle lane / no such query)
    float wt[2][4];
    int q_of[2];       // query of my half of pair i, or -1
    int bx0 = 0x7fffffff, bx1 = -1, by0 = 0x7fffffff, by1 = -1;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int j = w * Cfg::UPW + 2 * i + h;                       // tile-local query
      const int y = y_base + (j >> Cfg::TXS), x = x_base + (j & (TX - 1));
      int q = (y < tH && x < tW) ? tfirst + y * tW + x : -1;
      if (q >= Lq) q = -1;
      q_of[i] = q;
      xy[i] = -1;
      wt[i][0] = wt[i][1] = wt[i][2] = wt[i][3] = 0.f;
      if (q >= 0 && have_s && !narrow) {
        const long long su = (((long long)b * Lq + q) * M + m) * LP + s;
        const float2 lxy = __ldg(reinterpret_cast<const float2*>(loc) + su);
        const float a_in = __ldg(attn + su);
        const float fH = (float)myH, fW = (float)myW;
        const float yy = fma(lxy.y, fH, -0.5f), xx = fma(lxy.x, fW, -0.5f);  // same roundings as make_geo / spec_record
        const bool inside = yy > -1.f && xx > -1.f && yy < fH && xx < fW;
        const float fy = floorf(yy), fx = floorf(xx);
        int y0 = (int)fy, x0 = (int)fx;
        const float ly = yy - fy, lx = xx - fx, hy = 1.f - ly, hx = 1.f - lx;
        float wy0 = hy, wy1 = ly, wx0 = hx, wx1 = lx;
        if (y0 < 0) { y0 = 0; wy0 = ly; wy1 = 0.f; } else if (y0 > myH - 2) { y0 = myH - 2; wy1 = hy; wy0 = 0.f; }
        if (x0 < 0) { x0 = 0; wx0 = lx; wx1 = 0.f; } else if (x0 > myW - 2) { x0 = myW - 2; wx1 = hx; wx0 = 0.f; }
        if (inside) {
          const float a = a_in;
          wt[i][0] = wy0 * wx0 * a; wt[i][1] = wy0 * wx1 * a; wt[i][2] = wy1 * wx0 * a; wt[i][3] = wy1 * wx1 * a;
          xy[i] = x0 | (y0 << 16);
          bx0 = min(bx0, x0); bx1 = max(bx1, x0); by0 = min(by0, y0); by1 = max(by1, y0);
        }
      }
    }
    // bounding boxes: lanes of one level reduce among themselves (the level of a lane is fixed), one lane per level publishes
    for (int l = 0; l < L; ++l) {
      const unsigned mk = __ballot_sync(0xffffffffu, have_s && lvl == l && bx1 >= 0);
      if (mk == 0) continue;
      if (have_s && lvl == l && bx1 >= 0) {
        const int a0 = __reduce_min_sync(mk, bx0), a1 = __reduce_max_sync(mk, bx1);
        const int c0 = __reduce_min_sync(mk, by0), c1 = __reduce_max_sync(mk, by1);
        if (lane == (__ffs(mk) - 1)) {
          atomicMin(&lv->xmin[l], a0); atomicMax(&lv->xmax[l], a1);
          atomicMin(&lv->ymin[l], c0); atomicMax(&lv->ymax[l], c1);
        }
      }
    }
    __syncthreads();

    // ---- B: window allocation (one thread), then the copy (everybody) ----
    if (threadIdx.x == 0) {
      int used = 0;
      for (int l = 0; l < L; ++l) {
        int ww = 0, rows = 0, staged = 1;
        if (lv->xmax[l] >= 0) {
          ww = lv->xmax[l] - lv->xmin[l] + 2;
          rows = ww * (lv->ymax[l] - lv->ymin[l] + 2);
          if (rows > Cfg::CAP - used) { staged = 0; rows = 0; }
        }
        lv->wbase[l] = 1 + used; lv->ww[l] = ww; lv->staged[l] = staged; lv->rows[l] = rows;
        used += rows;
      }
      lv->total_rows = used;
    }
    __syncthreads();
    {
      const char* vimg = reinterpret_cast<const char*>(value + (long long)b * S * MD + m * D) + (threadIdx.x & (LPR - 1)) * 16;
      for (int l = 0; l < L; ++l) {
        const int rows = lv->rows[l];
        if (rows == 0) continue;
        const int ww = lv->ww[l], Wl = lv->W[l];
        const int pix0 = lv->st[l] + lv->ymin[l] * Wl + lv->xmin[l];        // first pixel of the box
        const unsigned magic = (unsigned)(0xffffffffu / (unsigned)ww) + 1u;  // exact r / ww for r * ww < 2^32
        const unsigned dst0 = win_s + (unsigned)lv->wbase[l] * Cfg::ROWB + (unsigned)(threadIdx.x & (LPR - 1)) * 16u;
        const unsigned rowb = (unsigned)MD * 4u;                             // bytes between pixels (32-bit: S*M*D*4 <= 2^29)
        for (int r = threadIdx.x / LPR; r < rows; r += Cfg::THREADS / LPR) {
          const int ry = (int)__umulhi((unsigned)r, magic);
          win_cp_async16(dst0 + (unsigned)r * Cfg::ROWB, vimg + (unsigned)(pix0 + ry * (Wl - ww) + r) * rowb);
        }
      }
      win_cp_async_wait();
    }
    __syncthreads();

    // ---- C: the gather rounds, out of the windows ----
    const char* vbc = reinterpret_cast<const char*>(value + (long long)b * S * MD + (m * D + cl * VEC));
    const unsigned base_s = win_s + (unsigned)cl * 16u;
    const size_t mdb = (size_t)MD * sizeof(T);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (narrow) {  // a level narrower than 2 pixels has no regular 2x2 window: the flagged path does these units
        for (int k = 0; k < 2; ++k) {
          const int q = __shfl_sync(0xffffffffu, q_of[i], 16 * k);
          if (q >= 0)
            msda_fwd_unit_flagged<T, D, MC, false, false>(value, shapes, start, loc, attn, out, S, M, L, P, inv_p, QM, nullptr, 2, b, q * M + m, m);
        }
        continue;
      }
      // final record of my sample: a shared-memory window address, or (bit 31 of the stride) a global one
      unsigned off = 0, stride = 0;
      if (xy[i] >= 0) {
        const int x0 = xy[i] & 0xffff, y0 = xy[i] >> 16;
        if (lv->staged[lvl]) {
          const int ww = lv->ww[lvl];
          off = (unsigned)(lv->wbase[lvl] + (y0 - lv->ymin[lvl]) * ww + (x0 - lv->xmin[lvl])) * (unsigned)Cfg::ROWB;
          stride = (unsigned)ww * (unsigned)Cfg::ROWB;
        } else {
          off = (unsigned)((mySt + y0 * myW + x0) * MD) * (unsigned)sizeof(T);
          stride = ((unsigned)(myW * MD) * (unsigned)sizeof(T)) | 0x80000000u;
        }
      }
      __syncwarp();
      rec_a[lane] = make_uint4(off, stride, __float_as_uint(wt[i][0]), __float_as_uint(wt[i][1]));
      rec_b[lane] = make_float2(wt[i][2], wt[i][3]);
      __syncwarp();
#pragma unroll 1
      for (int k = 0; k < 2; ++k) {  // the two queries of the pair (warp-uniform)
        const int q = __shfl_sync(0xffffffffu, q_of[i], 16 * k);
        if (q < 0) continue;
        float2 acc[VEC / 2];
#pragma unroll
        for (int t = 0; t < VEC / 2; ++t) acc[t] = make_float2(0.f, 0.f);
        for (int k0 = 16 * k; k0 < 16 * k + LP; k0 += G) {
          const int src = k0 + g;
          const uint4 ra = rec_a[src];
          const float2 rb = rec_b[src];
          uint4 v0, v1, v2, v3;
          if (!__any_sync(0xffffffffu, (int)ra.y < 0)) {  // the common round: every row of it is in a window
            const unsigned a0 = base_s + ra.x, a1 = a0 + ra.y;
            v0 = lds128(a0); v1 = lds128(a0 + Cfg::ROWB); v2 = lds128(a1); v3 = lds128(a1 + Cfg::ROWB);
          } else if ((int)ra.y < 0) {                      // my sample's level is not staged: global tap rows
            const char* t0 = vbc + ra.x;
            const char* t1 = t0 + (ra.y & 0x7fffffffu);
            v0 = ldg128(t0); v1 = ldg128(t0 + mdb); v2 = ldg128(t1); v3 = ldg128(t1 + mdb);
          } else {
            const unsigned a0 = base_s + ra.x, a1 = a0 + ra.y;
            v0 = lds128(a0); v1 = lds128(a0 + Cfg::ROWB); v2 = lds128(a1); v3 = lds128(a1 + Cfg::ROWB);
          }
          const float wq[4] = {__uint_as_float(ra.z), __uint_as_float(ra.w), rb.x, rb.y};
          const uint4 vv[4] = {v0, v1, v2, v3};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            float f[VEC];
            Vec16<T>::unpack(vv[t], f);
#pragma unroll
            for (int c = 0; c < VEC / 2; ++c) acc[c] = fma2(wq[t], make_float2(f[2 * c], f[2 * c + 1]), acc[c]);
          }
        }
        float o[N_OUT];
        int first;
        bool owner;
        msda_fwd_reduce<T, D>(acc, o, first, owner);
        bool bad = false;
#pragma unroll
        for (int t = 0; t < N_OUT; ++t) bad = bad || !(fabsf(o[t]) <= 3.402823466e38f);
        const long long u = (long long)b * QM + (long long)q * M + m;
        if (__any_sync(0xffffffffu, bad)) {  // a non-finite value met a zero weight (or is simply there): the flagged path decides
          msda_fwd_unit_flagged<T, D, MC, false, false>(value, shapes, start, loc, attn, out, S, M, L, P, inv_p, QM, nullptr, 2, b, q * M + m, m);
        } else if (owner) {
          store_vals<T, N_OUT>(out + u * D + cl * VEC + first, o);
        }
      }
    }
    __syncthreads();  // the windows and the box words are rewritten by the next tile
  }

  // ---- queries beyond the level grids (Lq > sum H*W: not pixel-aligned after all), in plain unit order ----
  const int pix = min(pixels, Lq);
  const long long tail = (long long)(Lq - pix) * NM;
  for (long long t = (long long)blockIdx.x * Cfg::WARPS + w; t < tail; t += (long long)gridDim.x * Cfg::WARPS) {
    const int bm = (int)(t % NM);
    const int q = pix + (int)(t / NM);
    msda_fwd_unit_flagged<T, D, MC, false, false>(value, shapes, start, loc, attn, out, S, M, L, P, inv_p, QM, nullptr, 2, bm / M, q * M + bm % M, bm % M);
  }
}

}  // namespace msda
