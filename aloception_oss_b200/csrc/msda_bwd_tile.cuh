// msda_bwd_tile.cuh -- TILE-BINNED backward of the multi-scale deformable attention operator (sm_100a, fp32).
//
// BASELINE.json north_star: "the backward scatter into grad_value done by query-tiled privatised accumulation instead of
// global atomics".  The unit-ordered backward (msda_bwd_sg_kernel) sends every tap row (4 per sample) to L2 as a 128-byte
// `red` and is bound by the SM->L2 red rate (5.9 clk per row per SM: profiles/README.md); shared-memory float atomics are
// slower still (sm_100a has no fp32 shared-memory add).  This kernel privatises by OWNERSHIP instead of by atomics:
//
//   * a CTA owns a tile of TQ = 16 x 16 raster queries of one pyramid level (encoder self-attention: query i IS pixel i of
//     the pyramid), one head, one image, and walks the sampled levels one after the other;
//   * per sampled level every (sample, tap row) -- the two x-adjacent taps (y, x0), (y, x0 + 1) of a bilinear sample -- is a
//     16-byte RECORD {weight_left, weight_right, sample id}; the CTA counting-sorts its 2 * TQ * P records by destination
//     (y, x0) inside a window of at most MAXB destinations that it derives from the data (bounding box of its own samples,
//     shrunk around their mean if too large) with INTEGER shared-memory atomics (count, scan, fill);
//   * a group of D/4 lanes then owns one destination at a time: it loads the destination's two `value` rows ONCE, walks the
//     destination's records, reads each record's grad_out row from the tile's shared-memory copy (one LDS.128 per lane, a
//     whole 128-byte row per quarter warp: conflict-free), accumulates  weight * grad_out  for both pixels in REGISTERS and
//     forms <grad_out, value row> for both taps (what grad_attn / grad_loc need) in the same pass; it issues ONE red per
//     (pixel, destination, tile) -- 5 x fewer red rows than one per tap on the encoder shape -- and returns the two dot
//     products to the sample's owner thread through shared memory;
//   * records whose destination falls outside the window (arbitrary sampling locations: the operator cannot assume
//     locality) are "destinations of one record": same code, same results, only the privatisation is lost.
// So BOTH halves of the backward are privatised: the scatter (reds per destination instead of per tap) and the gather
// (each `value` row of the window is fetched once per tile instead of once per tap).
//
// MEASURED (profiles/r2_ncu_bwd_tile.md): 5 x fewer red rows and a quarter of the L2 traffic, but the schedule is bound by
// instruction issue -- 11.4 warp instructions per tap, as many as the unit-ordered kernel spends while IT waits on the red
// rate -- and ends at 309 us against 265 us on the encoder shape (local +-4 px offsets; random locations 518 vs 292 us).
// It is therefore an OPT-IN schedule (knob "bwd_tile_mode" = 2), not the default.
//
// Results: same maths as msda_bwd_sg_kernel / the reference (ms_deform_im2col_cuda.cuh:87-159); only the order of the
// fp32 additions differs (tolerances: tests/).  Queries need not be pixel-aligned for correctness -- Lq != sum H*W is
// handled with linear tiles -- but the speed-up comes from tiles whose samples are local.
//
// Reference semantics kept: sample dropped unless -1 < y < H and -1 < x < W (also NaN / Inf locations); zero padding per
// tap; grad_loc scaled by W_l / H_l; outside samples give exactly zero gradients.
#pragma once

#include <climits>

#include "msda_kernels.cuh"

namespace msda {

template <int D, int P, int TPQ, int MAXB>
struct BwdTileCfg {
  static constexpr int T = 16;             // tile edge (queries)
  static constexpr int TQ = T * T;         // queries per tile
  static constexpr int THREADS = TQ * TPQ;
  static constexpr int SPT = P / TPQ;      // samples per thread and level
  static constexpr int NS = TQ * P;        // samples per level pass
  static constexpr int NREC = 2 * NS;      // (sample, tap row) records per level pass
  static constexpr int LPG = D / 4;        // lanes per destination group: 16 bytes (4 channels) per lane
  static constexpr int GPW = 32 / LPG;     // destination groups per warp
  static constexpr int BPT = MAXB / THREADS;  // scan: bins per thread
  static constexpr int MAXL = 16;          // levels (shared-memory tables)
  static constexpr int SIDE = MAXB >= 2025 ? 45 : (MAXB >= 1024 ? 32 : 22);  // SIDE * SIDE <= MAXB
  static constexpr size_t GO_BYTES = (size_t)(TQ + 1) * D * 4;   // + one zero row (padding records point at it)
  static constexpr size_t REC_BYTES = (size_t)(NREC + 1) * 16;   // + one zero record
  static constexpr size_t DOT_BYTES = (size_t)NS * 4 * 4;
  static constexpr size_t CNT_BYTES = (size_t)(MAXB + 4) * 4;
  static constexpr int NCLS = 17;          // destinations are ordered by ceil(records / 4), capped at NCLS - 1
  static constexpr size_t ORD_BYTES = (size_t)MAXB * 8;  // destination descriptors
  static constexpr size_t MISC_BYTES = (size_t)(16 + THREADS / 32 + 9 * MAXL + 4 + 2 * 32) * 4;
  static constexpr size_t SMEM_BYTES = GO_BYTES + REC_BYTES + DOT_BYTES + CNT_BYTES + MISC_BYTES + ORD_BYTES;
  static_assert(P % TPQ == 0 && (P & (P - 1)) == 0, "P must be a power of two, divisible by the threads per query");
  static_assert(LPG == 8, "the destination pass is written for 8 lanes per row (D = 32)");
  static_assert(THREADS % 32 == 0 && THREADS <= 1024 && MAXB % THREADS == 0, "bad tile configuration");
  static_assert(NS <= 1024, "sample id has 10 bits");
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// shared-memory slots of the per-level reductions / counters
enum { BB_YMIN = 0, BB_YMAX, BB_XMIN, BB_XMAX, BB_SUMY, BB_SUMX, BB_CNT, BB_NDIR, BB_NNE, BB_NEXT, BB_N };
// per-level tables (shared memory, filled once per CTA)
enum { LV_H = 0, LV_W, LV_ST, LV_FIRST, LV_NX, LV_TW, LV_TH, LV_CUM, LV_INVTW, LV_N };

// grid: persistent CTAs (static round-robin over work items); block: THREADS; dynamic shared memory: SMEM_BYTES.
// Work item = (tile, image, head); tiles: linear tail tiles (queries beyond sum H*W) first, then the tiles of the
// coarsest query level ... finest (items with the least locality start first).
template <int D, int P, int TPQ, int MAXB>
__global__ void __launch_bounds__(BwdTileCfg<D, P, TPQ, MAXB>::THREADS, 2)
msda_bwd_tile_kernel(const float* __restrict__ go, const float* __restrict__ value, const int32_t* __restrict__ shapes,
                     const int32_t* __restrict__ start, const float* __restrict__ loc, const float* __restrict__ attn,
                     float* __restrict__ gv, float* __restrict__ gloc, float* __restrict__ gattn,
                     int N, int S, int M, int L, int Lq) {
  using C = BwdTileCfg<D, P, TPQ, MAXB>;
  constexpr int T = C::T, TQ = C::TQ, THREADS = C::THREADS, SPT = C::SPT, NS = C::NS, NREC = C::NREC, LPG = C::LPG,
                GPW = C::GPW, BPT = C::BPT, MAXL = C::MAXL, NCLS = C::NCLS;
  extern __shared__ __align__(16) unsigned char msda_dyn_smem[];
  float4* go_s = reinterpret_cast<float4*>(msda_dyn_smem);                                   // [TQ + 1][D / 4]
  uint4* rec = reinterpret_cast<uint4*>(msda_dyn_smem + C::GO_BYTES);                        // [NREC + 1]
  float* dots = reinterpret_cast<float*>(msda_dyn_smem + C::GO_BYTES + C::REC_BYTES);        // [4 taps][NS]
  int* cnt = reinterpret_cast<int*>(msda_dyn_smem + C::GO_BYTES + C::REC_BYTES + C::DOT_BYTES);  // [MAXB + 1]
  int* bb = cnt + MAXB + 4;                                                                  // [BB_N]
  int* wsum = bb + 16;                                                                       // [THREADS / 32]
  int* lv = wsum + THREADS / 32;                                                             // [LV_N][MAXL]
  int* glob = lv + LV_N * MAXL;                                                              // pixels, tail_tiles, ntiles
  int* cls = glob + 4;                                                                       // [32] destinations per class
  int* cur2 = cls + 32;                                                                      // [32] fill cursors per class
  // [MAXB] non-empty destinations, longest class first: {value-row element offset of (y, x0), first record | records << 12 |
  // left pixel exists << 24 | right pixel exists << 25} -- written once per destination by the thread that scans its bin
  uint2* desc = reinterpret_cast<uint2*>(cur2 + 32);

  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int qi = t / TPQ, sub = t % TPQ;  // my query inside the tile, my share of its P points
  const int gid = t / LPG, lig = t % LPG;
  const int MD = M * D, LP = L * P;

  // ---- level / tile tables (once per CTA, from the device-resident level tensors) ----
  if (t == 0) {
    int pixels = 0;
    for (int l = 0; l < L; ++l) {
      const int H = __ldg(shapes + 2 * l), W = __ldg(shapes + 2 * l + 1);
      lv[LV_H * MAXL + l] = H; lv[LV_W * MAXL + l] = W; lv[LV_ST * MAXL + l] = __ldg(start + l);
      lv[LV_FIRST * MAXL + l] = pixels;
      pixels += H * W;
      const int nx = (W + T - 1) / T, ny = (H + T - 1) / T;
      const int tw = max(1, (W + max(nx, 1) - 1) / max(nx, 1)), th = max(1, (H + max(ny, 1) - 1) / max(ny, 1));  // balanced edges (<= T)
      lv[LV_NX * MAXL + l] = nx; lv[LV_TW * MAXL + l] = tw; lv[LV_TH * MAXL + l] = th;
      lv[LV_INVTW * MAXL + l] = (65536 + tw - 1) / tw;  // i / tw == (i * inv) >> 16 for i < 256, tw <= 16
    }
    const int tail_tiles = Lq > pixels ? (Lq - pixels + TQ - 1) / TQ : 0;
    int cum = tail_tiles;  // processing order: tail tiles, then level L-1 ... 0
    for (int l = L - 1; l >= 0; --l) {
      lv[LV_CUM * MAXL + l] = cum;  // first tile of level l
      cum += lv[LV_NX * MAXL + l] * ((lv[LV_H * MAXL + l] + T - 1) / T);
    }
    glob[0] = pixels; glob[1] = tail_tiles; glob[2] = cum;
    bb[BB_YMIN] = INT_MAX; bb[BB_YMAX] = INT_MIN; bb[BB_XMIN] = INT_MAX; bb[BB_XMAX] = INT_MIN;
    bb[BB_SUMY] = 0; bb[BB_SUMX] = 0; bb[BB_CNT] = 0; bb[BB_NDIR] = 0; bb[BB_NNE] = 0; bb[BB_NEXT] = 0;
    rec[NREC] = make_uint4(0u, 0u, 0u, (unsigned)(TQ * D * 4));  // padding record: zero weights, zero grad_out row
  }
  if (t < D / 4) go_s[TQ * (D / 4) + t] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  const int pixels = glob[0], tail_tiles = glob[1], ntiles = glob[2];
  const long long items = (long long)ntiles * N * M;
  bool fenced = false;

  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
    const int m = (int)(item % M);
    const long long r1 = item / M;
    const int b = (int)(r1 % N);
    const int tile = (int)(r1 / N);
    // ---- which queries ----
    int q_first, qW = 1, qH = 1, ty0 = 0, tx0 = 0, tw = TQ, th = 1, inv_tw = 0;
    const bool linear = tile < tail_tiles;
    if (linear) {
      q_first = pixels + tile * TQ;  // one row of TQ consecutive queries
    } else {
      // CUM[l] = first tile of level l (coarsest level first): level l owns [CUM[l], CUM[l-1]), level 0 the rest
      int l = 0;
      for (int k = L - 1; k >= 0; --k)
        if (tile >= lv[LV_CUM * MAXL + k]) l = k;
      const int local = tile - lv[LV_CUM * MAXL + l];
      const int nx = lv[LV_NX * MAXL + l];
      tw = lv[LV_TW * MAXL + l]; th = lv[LV_TH * MAXL + l]; inv_tw = lv[LV_INVTW * MAXL + l];
      qH = lv[LV_H * MAXL + l]; qW = lv[LV_W * MAXL + l]; q_first = lv[LV_FIRST * MAXL + l];
      const int tyi = local / nx;
      ty0 = tyi * th;
      tx0 = (local - tyi * nx) * tw;
    }
    auto query_of = [&](int i, bool& ok) -> int {
      if (linear) {
        const int q = q_first + i;
        ok = q < Lq;
        return q;
      }
      const int iy = (i * inv_tw) >> 16, ix = i - iy * tw;
      const int y = ty0 + iy, x = tx0 + ix;
      const int q = q_first + y * qW + x;
      ok = iy < th && y < qH && x < qW && q < Lq;
      return q;
    };
    bool q_ok;
    const int q = query_of(qi, q_ok);
    const long long unit = ((long long)b * Lq + (q_ok ? q : 0)) * M + m;

    // ---- stage the tile's grad_out rows (LPG lanes x 16 bytes per row); the previous item's readers are behind a barrier ----
    for (int r = gid; r < TQ; r += THREADS / LPG) {
      bool ok;
      const int rq = query_of(r, ok);
      if (ok) cp_async16(go_s + r * (D / 4) + lig, go + (((long long)b * Lq + rq) * M + m) * D + lig * 4);
    }

    for (int l = 0; l < L; ++l) {
      const int H = lv[LV_H * MAXL + l], W = lv[LV_W * MAXL + l], st = lv[LV_ST * MAXL + l];
      const float fH = (float)H, fW = (float)W;
      // ---- A: my samples of this level ----
      int y0[SPT], x0[SPT];
      float fy[SPT], fx[SPT], aw[SPT];  // fractional parts (ly, lx) and attention weight (0 for a dropped sample)
      bool ins[SPT];
      const long long sbase = unit * LP + l * P + sub * SPT;
#pragma unroll
      for (int j = 0; j < SPT; ++j) {
        float lx_ = 0.f, ly_ = 0.f, a_ = 0.f;
        if (q_ok) {
          const float2 xy = __ldg(reinterpret_cast<const float2*>(loc + 2 * (sbase + j)));
          lx_ = xy.x; ly_ = xy.y;
          a_ = __ldg(attn + sbase + j);
        }
        const float y = fmaf(ly_, fH, -0.5f), x = fmaf(lx_, fW, -0.5f);  // same roundings as make_geo
        const bool inside = q_ok && y > -1.f && x > -1.f && y < fH && x < fW;
        const float gy = floorf(y), gx = floorf(x);
        ins[j] = inside;
        y0[j] = inside ? (int)gy : 0;
        x0[j] = inside ? (int)gx : 0;
        fy[j] = inside ? y - gy : 0.f;
        fx[j] = inside ? x - gx : 0.f;
        aw[j] = inside ? a_ : 0.f;
      }
      // ---- B: bounding box / mean of the (tap row y, x0) destinations of the tile ----
      {
        int ymin = INT_MAX, ymax = INT_MIN, xmin = INT_MAX, xmax = INT_MIN, sy = 0, sx = 0, n = 0;
#pragma unroll
        for (int j = 0; j < SPT; ++j) {
          if (!ins[j]) continue;
          const bool r0 = y0[j] >= 0, r1ok = y0[j] + 1 <= H - 1;
          if (r0) { ymin = min(ymin, y0[j]); ymax = max(ymax, y0[j]); sy += y0[j]; sx += x0[j]; ++n; }
          if (r1ok) { ymin = min(ymin, y0[j] + 1); ymax = max(ymax, y0[j] + 1); sy += y0[j] + 1; sx += x0[j]; ++n; }
          if (r0 || r1ok) { xmin = min(xmin, x0[j]); xmax = max(xmax, x0[j]); }
        }
        ymin = __reduce_min_sync(0xffffffffu, ymin); ymax = __reduce_max_sync(0xffffffffu, ymax);
        xmin = __reduce_min_sync(0xffffffffu, xmin); xmax = __reduce_max_sync(0xffffffffu, xmax);
        sy = __reduce_add_sync(0xffffffffu, sy); sx = __reduce_add_sync(0xffffffffu, sx);
        n = __reduce_add_sync(0xffffffffu, n);
        if (lane == 0 && n > 0) {
          atomicMin(&bb[BB_YMIN], ymin); atomicMax(&bb[BB_YMAX], ymax);
          atomicMin(&bb[BB_XMIN], xmin); atomicMax(&bb[BB_XMAX], xmax);
          atomicAdd(&bb[BB_SUMY], sy); atomicAdd(&bb[BB_SUMX], sx); atomicAdd(&bb[BB_CNT], n);
        }
      }
      __syncthreads();  // #1
      // ---- C: the window (CTA-uniform) ----
      const int nrec_lvl = bb[BB_CNT];
      int wy0 = 0, wx0 = 0, WW = 1, WH = 0;
      if (nrec_lvl > 0) {
        const int by0 = bb[BB_YMIN], by1 = bb[BB_YMAX], bx0 = bb[BB_XMIN], bx1 = bb[BB_XMAX];
        const int BW = bx1 - bx0 + 1, BH = by1 - by0 + 1;
        WW = BW; WH = BH; wy0 = by0; wx0 = bx0;
        if ((long long)BW * BH > MAXB) {  // shrink around the mean destination
          WW = min(BW, max(C::SIDE, MAXB / BH));
          WH = min(BH, MAXB / WW);
          const int cy = bb[BB_SUMY] / nrec_lvl, cx = bb[BB_SUMX] / nrec_lvl;
          wy0 = min(max(cy - WH / 2, by0), by1 - WH + 1);
          wx0 = min(max(cx - WW / 2, bx0), bx1 - WW + 1);
        }
      }
      const int nbins = WW * WH;
      for (int i = t; i <= nbins; i += THREADS) cnt[i] = 0;
      if (t < 64) cls[t] = 0;  // cls[32] + cur2[32]
      const float inv_ww = 1.0f / (float)WW;
      __syncthreads();  // #2: window read by everyone, counters zero
      if (t == 0) {     // the next level's reductions start from neutral values (next use is behind barriers #3..)
        bb[BB_YMIN] = INT_MAX; bb[BB_YMAX] = INT_MIN; bb[BB_XMIN] = INT_MAX; bb[BB_XMAX] = INT_MIN;
        bb[BB_SUMY] = 0; bb[BB_SUMX] = 0; bb[BB_CNT] = 0;
      }
      // ---- D: count ----
      int binof[SPT][2];  // bin of (sample, tap row): >= 0 binned, -1 direct, -2 no record
#pragma unroll
      for (int j = 0; j < SPT; ++j) {
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int y = y0[j] + rr;
          const bool have = ins[j] && y >= 0 && y <= H - 1;
          int bn = -2;
          if (have) {
            const int dy = y - wy0, dx = x0[j] - wx0;
            bn = (dy >= 0 && dy < WH && dx >= 0 && dx < WW) ? dy * WW + dx : -1;
            if (bn >= 0) atomicAdd(&cnt[bn], 1);
          }
          binof[j][rr] = bn;
        }
      }
      __syncthreads();  // #3
      // ---- E: exclusive scan of cnt[0 .. nbins) ----
      {
        int v[BPT], sum = 0;
#pragma unroll
        for (int k = 0; k < BPT; ++k) {
          const int i = t * BPT + k;
          v[k] = i < nbins ? cnt[i] : 0;
          sum += v[k];
        }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int up = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += up;
        }
        if (lane == 31) wsum[warp] = incl;
        // destinations by class c = ceil(records / 4) (= steps of the destination pass): a round of GPW destinations of one
        // class has no padding steps; empty bins are not destinations at all
#pragma unroll
        for (int k = 0; k < BPT; ++k)
          if (v[k] > 0) atomicAdd(&cls[min((v[k] + 3) >> 2, NCLS - 1)], 1);
        __syncthreads();  // #4
        int run = incl - sum;
        for (int w2 = 0; w2 < warp; ++w2) run += wsum[w2];
        // first position of every class, longest class first: lane c holds cls[c]; suffix sums by warp shuffles
        const int cval = lane < NCLS ? cls[lane] : 0;  // (cls[0] stays 0: empty bins are counted nowhere)
        int suf = cval;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int dn = __shfl_down_sync(0xffffffffu, suf, o);
          if (lane + o < 32) suf += dn;
        }
        const int above = suf - cval;  // destinations in classes above mine
        if (t == 0) bb[BB_NNE] = suf;  // all non-empty bins
#pragma unroll
        for (int k = 0; k < BPT; ++k) {
          const int i = t * BPT + k;
          const int c = v[k] > 0 ? min((v[k] + 3) >> 2, NCLS - 1) : 0;
          const int cstart = __shfl_sync(0xffffffffu, above, c);
          if (i < nbins) cnt[i] = run;
          if (v[k] > 0) {
            const int pos = cstart + atomicAdd(&cur2[c], 1);
            const int dy = __float2int_rz(((float)i + 0.5f) * inv_ww);  // i / WW (exact: i, WW <= 2048)
            const int y = wy0 + dy, x = wx0 + (i - dy * WW);
            const unsigned fl = (x >= 0 ? 1u : 0u) | (x + 1 <= W - 1 ? 2u : 0u);
            desc[pos] = make_uint2((unsigned)((y * W + x) * MD), (unsigned)run | ((unsigned)v[k] << 12) | (fl << 24));
          }
          run += v[k];
        }
      }
      __syncthreads();  // #5
      // ---- F: fill (binned records from the front, direct records from the back of `rec`) ----
#pragma unroll
      for (int j = 0; j < SPT; ++j) {
        const float hy = 1.f - fy[j], hx = 1.f - fx[j];
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int bn = binof[j][rr];
          if (bn == -2) continue;
          const float wy = rr ? fy[j] : hy;
          const float wl = wy * hx * aw[j], wr = wy * fx[j] * aw[j];
          const int s = qi * P + sub * SPT + j;
          const int y = y0[j] + rr;
          int slot;
          unsigned meta = (unsigned)s | ((unsigned)rr << 10);
          if (bn >= 0) {
            slot = atomicAdd(&cnt[bn], 1);
          } else {
            slot = NREC - 1 - atomicAdd(&bb[BB_NDIR], 1);
            meta |= (unsigned)(y * (W + 1) + x0[j] + 1) << 11;  // < 2^21: the host checks S <= 2^19
          }
          rec[slot] = make_uint4(__float_as_uint(wl), __float_as_uint(wr), meta, (unsigned)(qi * D * 4));
        }
      }
      if (l == 0) cp_async_wait_all();  // my part of the grad_out tile has landed
      __syncthreads();  // #6: records, bin ends (cnt[b] = end of bin b) and the grad_out tile are visible
      // ---- G: destinations ----
      if (!fenced) {  // grad_value is zero-filled by the previous kernel in the stream (programmatic dependent launch)
        pdl_wait();
        fenced = true;
      }
      {
        const long long lvl_off = ((long long)b * S + st) * MD + m * D + lig * 4;
        const float* __restrict__ vlev = value + lvl_off;
        float* __restrict__ glev = gv + lvl_off;
        const int ndir = bb[BB_NDIR], nne = bb[BB_NNE];
        const int ndest = nne + ndir;  // destination k < nne: desc[k]; else direct record NREC - ndir + (k - nne)
        const int gw = lane / LPG;
        // The 8 lanes of a group sum 8 values per step -- <grad_out, left row> and <grad_out, right row> of 4 records -- with
        // a butterfly reduce-scatter (4 + 2 + 1 shuffles).  Lane `lig` = (b2 b1 b0) keeps value index m = b0*4 + b1*2 + b2 and
        // holds true value (k ^ m) in its slot k, i.e. it walks the 4 records in the order j ^ (2*b0 + b1) and swaps the two
        // value rows when b2 is set: every lane then keeps the low half of its slots in every step (no selects).
        const int pr = ((lig & 1) << 1) | ((lig >> 1) & 1), b2 = (lig >> 2) & 1;
        const uint32_t rec_a = smem_u32(rec), go_a = smem_u32(go_s) + (uint32_t)lig * 16u;
        // destination -> record range, value-row offset, which of its two pixels exist
        auto dest_of = [&](int k, int& beg, int& end, int& voff, unsigned& fl) {
          beg = 0; end = 0; voff = 0; fl = 0u;
          if (k < nne) {
            const uint2 d = desc[k];
            voff = (int)d.x;
            beg = (int)(d.y & 4095u);
            end = beg + (int)((d.y >> 12) & 4095u);
            fl = d.y >> 24;
          } else if (k < ndest) {
            beg = NREC - ndir + (k - nne);
            end = beg + 1;
            const int pix = (int)(rec[beg].z >> 11);
            const int y = pix / (W + 1), x = pix - y * (W + 1) - 1;
            voff = (y * W + x) * MD;
            fl = (x >= 0 ? 1u : 0u) | (x + 1 <= W - 1 ? 2u : 0u);
          }
        };
        auto fetch = [&](int voff, unsigned fl, float4& vl, float4& vr) {
          const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
          const float* p = vlev + voff;  // 32-bit element offset: the host checks S * M * D < 2^29
          vl = (fl & 1u) ? __ldg(reinterpret_cast<const float4*>(p)) : z;
          vr = (fl & 2u) ? __ldg(reinterpret_cast<const float4*>(p + MD)) : z;
        };
        // rounds of GPW destinations, strided over the warps: the destinations are ordered by class, so every warp gets the
        // same mix of long and short ones.  (Handing the rounds out dynamically through a shared-memory counter measured
        // 316 vs 309 us at ENC: the atomic + broadcast per round cost more than the barrier wait they save.)
        constexpr int RSTRIDE = (THREADS / 32) * GPW;
        int kb = warp * GPW;
        int beg, end, voff;
        unsigned fl;
        float4 vl = make_float4(0.f, 0.f, 0.f, 0.f), vr = vl;
        dest_of(kb + gw, beg, end, voff, fl);
        if (end > beg) fetch(voff, fl, vl, vr);
        while (kb < ndest) {  // warp-uniform
          // the next round's rows are in flight while this one's records are summed
          const int kn = kb + RSTRIDE;
          int nbeg, nend, nvoff;
          unsigned nfl;
          float4 nvl = make_float4(0.f, 0.f, 0.f, 0.f), nvr = nvl;
          dest_of(kn + gw, nbeg, nend, nvoff, nfl);
          if (nend > nbeg) fetch(nvoff, nfl, nvl, nvr);
          const int trips = __reduce_max_sync(0xffffffffu, end - beg);
          if (trips > 0) {
            const float2 va0 = b2 ? make_float2(vr.x, vr.y) : make_float2(vl.x, vl.y);
            const float2 va1 = b2 ? make_float2(vr.z, vr.w) : make_float2(vl.z, vl.w);
            const float2 vb0 = b2 ? make_float2(vl.x, vl.y) : make_float2(vr.x, vr.y);
            const float2 vb1 = b2 ? make_float2(vl.z, vl.w) : make_float2(vr.z, vr.w);
            float2 al0 = make_float2(0.f, 0.f), al1 = al0, ar0 = al0, ar1 = al0;
            for (int i0 = 0; i0 < trips; i0 += 4) {  // warp-uniform trip count; short destinations read the padding record
              float d[8];
              unsigned meta0 = 0u;
              bool act0 = false;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int r = beg + i0 + (j ^ pr);
                const bool act = r < end;
                const uint4 rc = lds128_u32(rec_a + (uint32_t)(act ? r : NREC) * 16u);
                const uint4 gu = lds128_u32(go_a + rc.w);
                const float2 g0 = make_float2(__uint_as_float(gu.x), __uint_as_float(gu.y));
                const float2 g1 = make_float2(__uint_as_float(gu.z), __uint_as_float(gu.w));
                const float wl = __uint_as_float(rc.x), wr = __uint_as_float(rc.y);
                const float2 wl2 = make_float2(wl, wl), wr2 = make_float2(wr, wr);
                al0 = __ffma2_rn(wl2, g0, al0); al1 = __ffma2_rn(wl2, g1, al1);
                ar0 = __ffma2_rn(wr2, g0, ar0); ar1 = __ffma2_rn(wr2, g1, ar1);
                const float2 pa = __ffma2_rn(g1, va1, __fmul2_rn(g0, va0));
                const float2 pb = __ffma2_rn(g1, vb1, __fmul2_rn(g0, vb0));
                d[2 * j] = pa.x + pa.y;
                d[2 * j + 1] = pb.x + pb.y;
                if (j == 0) { meta0 = rc.z; act0 = act; }
              }
              float k4[4], k2[2];
#pragma unroll
              for (int j = 0; j < 4; ++j) k4[j] = d[j] + __shfl_xor_sync(0xffffffffu, d[4 + j], 1);
#pragma unroll
              for (int j = 0; j < 2; ++j) k2[j] = k4[j] + __shfl_xor_sync(0xffffffffu, k4[2 + j], 2);
              const float k1 = k2[0] + __shfl_xor_sync(0xffffffffu, k2[1], 4);
              // my slot 0 = record `pr` of this step, tap side b2
              if (act0) dots[(int)((((meta0 >> 10) & 1u) * 2u + (unsigned)b2) * NS + (meta0 & 1023u))] = k1;
            }
            if (end > beg) {
              float* gp = glev + voff;
              if (fl & 1u) red_add_v4(gp, al0.x, al0.y, al1.x, al1.y);
              if (fl & 2u) red_add_v4(gp + MD, ar0.x, ar0.y, ar1.x, ar1.y);
            }
          }
          kb = kn; beg = nbeg; end = nend; voff = nvoff; fl = nfl; vl = nvl; vr = nvr;
        }
      }
      __syncthreads();  // #7: dots complete; records / counters free
      if (t == 0) { bb[BB_NDIR] = 0; bb[BB_NEXT] = 0; }  // (next use is behind barriers #1.. of the next level)
      // ---- H: grad_attn / grad_loc of my samples ----
      if (q_ok) {
        float ga[SPT], glx[SPT], gly[SPT];
#pragma unroll
        for (int j = 0; j < SPT; ++j) {
          const int s = qi * P + sub * SPT + j;
          const bool row0 = ins[j] && y0[j] >= 0, row1 = ins[j] && y0[j] + 1 <= H - 1;
          const bool c0 = x0[j] >= 0, c1 = x0[j] + 1 <= W - 1;
          const float r00 = (row0 && c0) ? dots[0 * NS + s] : 0.f;
          const float r01 = (row0 && c1) ? dots[1 * NS + s] : 0.f;
          const float r10 = (row1 && c0) ? dots[2 * NS + s] : 0.f;
          const float r11 = (row1 && c1) ? dots[3 * NS + s] : 0.f;
          const float ly = fy[j], lx = fx[j];
          const float hy = ins[j] ? 1.f - ly : 0.f, hx = ins[j] ? 1.f - lx : 0.f;
          const float top = hx * r00 + lx * r01;
          const float bot = hx * r10 + lx * r11;
          ga[j] = hy * top + ly * bot;
          glx[j] = fW * aw[j] * (hy * (r01 - r00) + ly * (r11 - r10));
          gly[j] = fH * aw[j] * (bot - top);
        }
        if constexpr (SPT == 2) {
          *reinterpret_cast<float2*>(gattn + sbase) = make_float2(ga[0], ga[1]);
          *reinterpret_cast<float4*>(gloc + 2 * sbase) = make_float4(glx[0], gly[0], glx[1], gly[1]);
        } else if constexpr (SPT == 4) {
          *reinterpret_cast<float4*>(gattn + sbase) = make_float4(ga[0], ga[1], ga[2], ga[3]);
          *reinterpret_cast<float4*>(gloc + 2 * sbase) = make_float4(glx[0], gly[0], glx[1], gly[1]);
          *reinterpret_cast<float4*>(gloc + 2 * sbase + 4) = make_float4(glx[2], gly[2], glx[3], gly[3]);
        } else {
#pragma unroll
          for (int j = 0; j < SPT; ++j) {
            gattn[sbase + j] = ga[j];
            *reinterpret_cast<float2*>(gloc + 2 * (sbase + j)) = make_float2(glx[j], gly[j]);
          }
        }
      }
      // no barrier here: the next level's phases A..C touch neither `dots` nor `rec`; `cnt` is re-zeroed in C, after
      // barrier #1, which every thread reaches only after finishing this epilogue's reads
    }
    __syncthreads();  // the next item's staging overwrites the grad_out tile (read in phase G of the last level)
  }
}

}  // namespace msda
