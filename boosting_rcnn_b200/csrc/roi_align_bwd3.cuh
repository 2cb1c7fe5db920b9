// Shared pieces of the tiled RoIAlign backward gathers: argument / workspace layout of the
// per-tile path (roi_bwd_prep_kernel in roi_align_bwd2.cuh fills the tables, the (image, level)
// buckets and the per-tile RoI lists; roi_bwd_gather5_kernel in roi_align_bwd5.cuh consumes
// them) and the small asynchronous-copy / packed-FMA helpers.
//
// Reference: autograd of mmcv RoIAlign (roi_align_backward, 4 atomicAdd per sample) reached from
// single_level_roi_extractor.py:79,103; levels without RoIs still receive a zero gradient like
// :105-114.  The gradient one RoI sends to feature pixel (y,x) is separable,
//     G[y][x][c] = sum_ph Wy[y][ph] * ( sum_pw Wx[x][pw] * g[ph][pw][c] ),
// with the Wy (1/count folded in) / Wx tables of the prep kernel; grad_out is consumed
// bin-major, (R, PH*PW, C): the layout the (R,7,7,C) feature hand-off produces (no transpose),
// where a (bin, 128-channel) piece is one contiguous 512 B run.
//
// History (measured on B200, 1024 RoIs): v3 = one CTA per (tile, slab) with a producer warp
// staging gradient blocks for 8 consumer warps through an mbarrier ring, 0.090 ms; v4 = the same
// with persistent CTAs and per-tile lists, 0.084 ms; both were bound by per-RoI latency chains
// through the single producer.  v5 (roi_align_bwd5.cuh), 0.059 ms, replaced them.
#pragma once
#include "common.cuh"
#include "roi_align.cuh"
#include "roi_align_bwd2.cuh"
#include "roi_align_tma.cuh"

namespace brcnn {

constexpr int B3_TS = 8;                         // tile side in pixels
constexpr int B3_CS = 128;                       // channels per CTA (one quad per lane)

struct RoiBwd3Args {
  RoiArgs a;
  float* grad[BRCNN_MAX_LEVELS];  // NHWC (B,H,W,C)
  int tiles_x[BRCNN_MAX_LEVELS], tiles_y[BRCNN_MAX_LEVELS];
  int tile_first[BRCNN_MAX_LEVELS];  // first tile (CTA x) of each level; coarse levels first
  int TR;                            // table rows per RoI = max_h + max_w
  int bucket_cap;                    // entries per (image, level) bucket = R
};

__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<const unsigned long long&>(a)),
        "l"(reinterpret_cast<const unsigned long long&>(b)));
}

struct RoiBwd3Ws {
  size_t recs, keys, tab, gt, bucket_cnt, tile_cnt, work_counter, zero_bytes, bucket, bucket_rec,
      tile_r, tile_rec, total;
};
constexpr int B4_TILE_CAP = 128;   // entries per tile list (== B5_WIN)
static inline long long roi_bwd3_tiles(const RoiArgs& a) {
  long long t = 0;
  for (int l = 0; l < a.L; ++l)
    t += (long long)((a.W[l] + B3_TS - 1) / B3_TS) * ((a.H[l] + B3_TS - 1) / B3_TS) * a.B;
  return t;
}
static inline RoiBwd3Ws roi_bwd3_ws(const RoiArgs& a, int R, bool need_transpose) {
  RoiBwd3Ws w;
  size_t o = 0;
  const size_t Rn = (size_t)(R > 0 ? R : 1);
  int max_h = 0, max_w = 0;
  for (int l = 0; l < a.L; ++l) { max_h = a.H[l] > max_h ? a.H[l] : max_h; max_w = a.W[l] > max_w ? a.W[l] : max_w; }
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  w.recs = o; o = al(o + Rn * sizeof(RoiBwdRec));
  w.keys = o; o = al(o + Rn * 2);
  w.tab = o;  o = al(o + Rn * (size_t)(max_h + max_w) * 8 * 4);
  w.gt = o;   o = al(o + (need_transpose ? Rn * (size_t)a.PH * a.PW * a.C * 4 : 0));
  w.bucket_cnt = o; o = al(o + (size_t)a.B * a.L * 4);
  w.tile_cnt = o;   o = al(o + (size_t)roi_bwd3_tiles(a) * 4);
  w.work_counter = o; o = al(o + 4);
  w.zero_bytes = o - w.bucket_cnt;                 // bucket_cnt + tile_cnt + counter: one memset
  w.bucket = o;     o = al(o + (size_t)a.B * a.L * Rn * 4);
  w.bucket_rec = o; o = al(o + (size_t)a.B * a.L * Rn * sizeof(RoiBwdRec));
  w.tile_r = o;     o = al(o + (size_t)roi_bwd3_tiles(a) * B4_TILE_CAP * 4);
  w.tile_rec = o;   o = al(o + (size_t)roi_bwd3_tiles(a) * B4_TILE_CAP * sizeof(RoiBwdRec));
  w.total = o;
  return w;
}

}  // namespace brcnn
