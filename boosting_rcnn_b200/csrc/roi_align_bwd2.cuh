// K4 (v2): multi-level RoIAlign backward as a deterministic, table-driven
// gather (no float atomics, no per-RoI block synchronisation).
// Reference: autograd of mmcv RoIAlign (roi_align_backward, 4 atomicAdd per
// sample) reached from single_level_roi_extractor.py:79,103; levels without
// RoIs still receive a zero gradient like :105-114.
//
// The gradient a RoI sends to feature pixel (y,x) is separable,
//     G[y][x][c] = sum_ph Wy[y][ph] * sum_pw Wx[x][pw] * g[c][ph][pw],
// with the same Wy (1/count folded in) / Wx tables as the forward kernel.
//
//   roi_bwd_prep_kernel     one CTA per RoI: geometry, footprint box, (image,
//                           level) key and the Wy / Wx tables -> global memory,
//                           each table row = 7 weights + its non-zero band
//   transpose_multi_kernel  grad_out (R,C,49) -> (R,49,C): bins become
//                           128-bit channel-quad loads
//   roi_bwd_gather2_kernel  one CTA per 8x8-pixel tile x 128-channel slab.
//                           Phase A builds the index-ordered list of RoIs that
//                           touch the tile (2-byte key prefilter, ballot scan).
//                           Phase B: warp = one tile row (uniform y, uniform ph
//                           band), lane = (x, quad group), 8 channel quads per
//                           thread; warps walk the list independently, weights
//                           come from the (L1-resident) tables.  Every output
//                           element is written exactly once with a fixed
//                           summation order -> bit-reproducible.
#pragma once
#include "common.cuh"
#include "roi_align.cuh"

namespace brcnn {

constexpr int B2_TS = 8;            // tile side in pixels
constexpr int B2_QPT = 8;           // channel quads per thread
constexpr int B2_CCH = 16 * B2_QPT;  // channels per CTA (4 quad groups x B2_QPT quads)
constexpr int B2_THREADS = 256;     // 8 rows x 8 px x 4 quad groups
constexpr int B2_LIST = 1024;       // RoIs gathered per round
constexpr int B2_P = 7;             // max pooled side on this path
constexpr int B2_PREP_THREADS = 128;   // 16 table rows per pass (a RoI has ~20)

struct RoiBwdRec {      // 16 B, read by every tile CTA
  int ylo, yhi, xlo, xhi;   // inclusive footprint; empty if ylo > yhi
};

// optional outputs of roi_bwd_prep_kernel for the v3 gather (roi_align_bwd3.cuh)
struct RoiBwdBuckets {
  int32_t* bucket;       // [B*L][R] RoI ids per (image, level), arrival order
  RoiBwdRec* bucket_rec; // [B*L][R] their footprint boxes (same order)
  int32_t* bucket_cnt;   // [B*L]
  int32_t* tile_cnt;     // [tiles] number of RoIs whose footprint box touches the tile
  int32_t* tile_r;       // [tiles][tile_cap] their ids (arrival order; only the first tile_cap)
  RoiBwdRec* tile_rec;   // [tiles][tile_cap] and footprint boxes
  int tile_cap;
  int tile_side;
  int tiles_x[BRCNN_MAX_LEVELS], tiles_y[BRCNN_MAX_LEVELS], tile_first[BRCNN_MAX_LEVELS];
};

struct RoiBwd2Args {
  RoiArgs a;
  float* grad[BRCNN_MAX_LEVELS];  // NHWC (B,H,W,C)
  int tiles_x[BRCNN_MAX_LEVELS], tiles_y[BRCNN_MAX_LEVELS];
  int tile_base[BRCNN_MAX_LEVELS + 1];
  int TR;               // table rows per RoI = max_h + max_w
};

// grid R, block B2_PREP_THREADS.  tab: [R][TR][8] floats, key: uint16 (b*L + lvl, 0xFFFF = none)
__global__ void __launch_bounds__(B2_PREP_THREADS)
roi_bwd_prep_kernel(const __grid_constant__ RoiArgs a, const float* __restrict__ rois, int R,
                    int TR, RoiBwdRec* __restrict__ recs, unsigned short* __restrict__ keys,
                    float* __restrict__ tab, const RoiBwdBuckets bk) {
  const int r = blockIdx.x;
  const int tid = threadIdx.x;
  const float* roi = rois + (size_t)r * 5;
  RoiBwdRec rec;
  rec.ylo = rec.xlo = 1; rec.yhi = rec.xhi = 0;
  unsigned short key = 0xFFFFu;
  RoiGeom g;
  bool ok = false;
  if (!(roi[0] < 0.f)) {
    g = roi_geometry(a, roi);
    if (g.b >= 0 && g.b < a.B) {
      roi_axis_range(g.start_h, g.bin_h, a.PH, g.gh, g.H, rec.ylo, rec.yhi);
      roi_axis_range(g.start_w, g.bin_w, a.PW, g.gw, g.W, rec.xlo, rec.xhi);
      ok = (rec.ylo <= rec.yhi) && (rec.xlo <= rec.xhi);
      if (ok) key = (unsigned short)(g.b * a.L + g.lvl);
      else { rec.ylo = rec.xlo = 1; rec.yhi = rec.xhi = 0; }
    }
  }
  if (tid == 0) {
    recs[r] = rec; keys[r] = key;
    // v3 gather: append to the (image, level) bucket; arrival order is arbitrary, the
    // gather kernel sorts each tile's list by RoI index
    if (ok && bk.bucket != nullptr) {
      const size_t slot = (size_t)key * R + atomicAdd(bk.bucket_cnt + key, 1);
      bk.bucket[slot] = r;
      bk.bucket_rec[slot] = rec;
    }
  }
  if (!ok) return;
  if (bk.tile_cnt != nullptr) {
    const int ts = bk.tile_side;
    const int tx0 = rec.xlo / ts, ty0 = rec.ylo / ts;
    const int ntx = rec.xhi / ts - tx0 + 1, nty = rec.yhi / ts - ty0 + 1;
    int32_t* tc = bk.tile_cnt + bk.tile_first[g.lvl] +
                  (size_t)g.b * bk.tiles_x[g.lvl] * bk.tiles_y[g.lvl];
    for (int i = tid; i < ntx * nty; i += B2_PREP_THREADS) {
      const int dy = i / ntx, dx = i - dy * ntx;
      int32_t* cell = tc + (ty0 + dy) * bk.tiles_x[g.lvl] + tx0 + dx;
      const int pos = atomicAdd(cell, 1);
      if (bk.tile_r != nullptr && pos < bk.tile_cap) {
        const size_t slot = (size_t)(cell - bk.tile_cnt) * bk.tile_cap + pos;
        bk.tile_r[slot] = r;
        bk.tile_rec[slot] = rec;
      }
    }
  }
  // tables: thread = (table row, pooled index); 8 consecutive lanes own one row, the row's
  // non-zero band comes from a ballot over them
  const int fh = rec.yhi - rec.ylo + 1, fw = rec.xhi - rec.xlo + 1;
  float* t = tab + (size_t)r * TR * 8;
  const int lane = tid & 31, p = tid & 7;
  for (int row0 = 0; row0 < fh + fw; row0 += B2_PREP_THREADS / 8) {
    const int row = row0 + (tid >> 3);
    const bool live = row < fh + fw;
    const bool isy = row < fh;
    const int pos = isy ? rec.ylo + row : rec.xlo + (row - fh);
    float v = 0.f;
    if (live && p < (isy ? a.PH : a.PW))
      v = isy ? roi_axis_weight(g.start_h, g.bin_h, g.gh, g.H, p, pos) * g.inv_count
              : roi_axis_weight(g.start_w, g.bin_w, g.gw, g.W, p, pos);
    const unsigned nz = (__ballot_sync(0xffffffffu, v != 0.f) >> (lane & 24)) & 0x7fu;
    if (p == 7) v = __int_as_float(nz ? ((__ffs(nz) - 1) | ((31 - __clz(nz)) << 8)) : 1);
    if (live) t[(size_t)(isy ? row : a.max_h + (row - fh)) * 8 + p] = v;
  }
}

// grid (total_tiles, ceil(C / B2_CCH))
__global__ void __launch_bounds__(B2_THREADS)
roi_bwd_gather2_kernel(const __grid_constant__ RoiBwd2Args ba,
                       const RoiBwdRec* __restrict__ recs,
                       const unsigned short* __restrict__ keys, int R,
                       const float* __restrict__ tab,
                       const float* __restrict__ gt /* (R, nbins, C) */) {
  __shared__ int s_list[B2_LIST];
  __shared__ int s_warp[B2_THREADS / 32];
  __shared__ int s_n;

  const RoiArgs& a = ba.a;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  int lvl = 0;
  while (lvl + 1 < a.L && (int)blockIdx.x >= ba.tile_base[lvl + 1]) ++lvl;
  int t = blockIdx.x - ba.tile_base[lvl];
  const int tpi = ba.tiles_x[lvl] * ba.tiles_y[lvl];
  const int b = t / tpi; t -= b * tpi;
  const int ty = t / ba.tiles_x[lvl], tx = t - ty * ba.tiles_x[lvl];
  const int y0 = ty * B2_TS, x0 = tx * B2_TS;
  const int H = a.H[lvl], W = a.W[lvl], C = a.C;
  const int c0 = blockIdx.y * B2_CCH;
  const int PW = a.PW, nbins = a.PH * a.PW;
  const unsigned short mykey = (unsigned short)(b * a.L + lvl);
  // warp = tile row, lane = (x, quad group); thread owns quads qg + 4*k, k < B2_QPT
  const int y = y0 + wid;
  const int x = x0 + (lane >> 2);
  const int qg = lane & 3;
  const int cbase = c0 + qg * 4;            // channel of quad k: cbase + 16*k
  const int nq = max(0, min(B2_QPT, (C - cbase + 15) / 16));  // valid quads of this thread
  const int TR = ba.TR;

  float4 acc[B2_QPT];
#pragma unroll
  for (int k = 0; k < B2_QPT; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);

  int r_next = 0;
  while (r_next < R) {
    // ---- phase A: index-ordered list of the RoIs touching this tile ----
    if (tid == 0) s_n = 0;
    __syncthreads();
    while (r_next < R) {
      const int n_now = s_n;
      if (n_now + B2_THREADS > B2_LIST) break;
      const int r = r_next + tid;
      bool f = false;
      if (r < R && keys[r] == mykey) {
        const int4 q = *reinterpret_cast<const int4*>(recs + r);
        f = (q.x <= y0 + B2_TS - 1) && (q.y >= y0) && (q.z <= x0 + B2_TS - 1) && (q.w >= x0);
      }
      const unsigned bm = __ballot_sync(0xffffffffu, f);
      if (lane == 0) s_warp[wid] = __popc(bm);
      __syncthreads();
      int wbase = 0, tot = 0;
#pragma unroll
      for (int w = 0; w < B2_THREADS / 32; ++w) {
        const int c = s_warp[w];
        if (w < wid) wbase += c;
        tot += c;
      }
      if (f) s_list[n_now + wbase + __popc(bm & ((1u << lane) - 1u))] = r;
      __syncthreads();
      if (tid == 0) s_n = n_now + tot;
      r_next += B2_THREADS;
      __syncthreads();
    }
    const int n = s_n;
    // ---- phase B: every warp walks the list on its own (no block syncs) ----
    if (y < H) {
      for (int li = 0; li < n; ++li) {
        const int r = s_list[li];
        const int4 rec = __ldg(reinterpret_cast<const int4*>(recs + r));
        if (y < rec.x || y > rec.y) continue;            // warp-uniform
        const float* ty_row = tab + ((size_t)r * TR + (y - rec.x)) * 8;
        const int pk = __float_as_int(__ldg(ty_row + 7));
        const int pa = pk & 0xff, pb = pk >> 8;
        if (pa > pb) continue;                           // warp-uniform
        const bool xin = (x >= rec.z) && (x <= rec.w) && (x < W) && (nq > 0);
        int qa = 1, qb = 0;
        const float* tx_row = tab;
        if (xin) {
          tx_row = tab + ((size_t)r * TR + a.max_h + (x - rec.z)) * 8;
          const int qk = __float_as_int(__ldg(tx_row + 7));
          qa = qk & 0xff; qb = qk >> 8;
        }
        const float* gr = gt + (size_t)r * nbins * C + cbase;
        for (int ph = pa; ph <= pb; ++ph) {
          const float wy = __ldg(ty_row + ph);
          if (wy == 0.f) continue;                       // warp-uniform
          for (int pw = qa; pw <= qb; ++pw) {
            const float w = wy * __ldg(tx_row + pw);
            const float* gp = gr + (size_t)(ph * PW + pw) * C;
#pragma unroll
            for (int k = 0; k < B2_QPT; ++k) {
              if (k < nq) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(gp + 16 * k));
                acc[k].x = fmaf(w, v.x, acc[k].x);
                acc[k].y = fmaf(w, v.y, acc[k].y);
                acc[k].z = fmaf(w, v.z, acc[k].z);
                acc[k].w = fmaf(w, v.w, acc[k].w);
              }
            }
          }
        }
      }
    }
    __syncthreads();
  }
  if (y < H && x < W) {
    float* gout = ba.grad[lvl] + (((size_t)b * H + y) * W + x) * C + cbase;
#pragma unroll
    for (int k = 0; k < B2_QPT; ++k)
      if (k < nq) *reinterpret_cast<float4*>(gout + 16 * k) = acc[k];
  }
}

struct RoiBwd2Ws {
  size_t recs, keys, tab, gt, total;
};
static inline RoiBwd2Ws roi_bwd2_ws(const RoiArgs& a, int R) {
  RoiBwd2Ws w;
  size_t o = 0;
  const size_t Rn = (size_t)(R > 0 ? R : 1);
  int max_h = 0, max_w = 0;
  for (int l = 0; l < a.L; ++l) { max_h = a.H[l] > max_h ? a.H[l] : max_h; max_w = a.W[l] > max_w ? a.W[l] : max_w; }
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  w.recs = o; o = al(o + Rn * sizeof(RoiBwdRec));
  w.keys = o; o = al(o + Rn * 2);
  w.tab = o;  o = al(o + Rn * (size_t)(max_h + max_w) * 8 * 4);
  w.gt = o;   o = al(o + Rn * (size_t)a.PH * a.PW * a.C * 4);
  w.total = o;
  return w;
}

}  // namespace brcnn
