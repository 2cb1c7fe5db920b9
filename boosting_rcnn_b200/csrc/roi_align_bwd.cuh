// K4 (v1, kept for pooled sizes > 7; roi_align_bwd2.cuh is the main path):
// multi-level RoIAlign backward as a deterministic gather (no float atomics).
// Reference: autograd of mmcv RoIAlign (roi_align_backward, 4 atomicAdd per sample) reached from single_level_roi_extractor.py:79,103;
// empty levels still receive a (zero) gradient like :105-114.
//
// Because bilinear weights are separable, the gradient a RoI sends to the
// feature pixel (y,x) is
//     (1/count) * sum_ph sum_pw Wy[y][ph] * Wx[x][pw] * g[ph][pw]
// with Wy[y][ph] = sum over the bin's sample rows of the tap weight that
// lands on row y (same for Wx).  Each 8x8-pixel tile of every (image, level)
// map is owned by one CTA per 64-channel slab; it scans the RoIs in index
// order, so every output element is written exactly once with a fixed
// summation order -> bit-reproducible.
//
//   roi_bwd_geom_kernel    per-RoI geometry (level, footprint box)
//   transpose_multi_kernel grad_out (R,C,49) -> (R,49,C) so taps are 128-bit
//   roi_bwd_gather_kernel  the tile-owner gather
#pragma once
#include "common.cuh"
#include "roi_align.cuh"

namespace brcnn {

constexpr int BWD_TS = 8;          // tile side in pixels
constexpr int BWD_CCH = 64;        // channels per CTA
constexpr int BWD_THREADS = 256;
constexpr int BWD_LIST = 1024;     // RoIs gathered per round
constexpr int BWD_MAXP = 14;       // max pooled side supported by the tables

struct RoiBwdGeom {
  RoiGeom g;
  int ylo, yhi, xlo, xhi;  // conservative footprint (inclusive), empty if ylo>yhi
};

struct RoiBwdArgs {
  RoiArgs a;
  float* grad[BRCNN_MAX_LEVELS];  // NHWC (B,H,W,C)
  int tiles_x[BRCNN_MAX_LEVELS], tiles_y[BRCNN_MAX_LEVELS];
  int tile_base[BRCNN_MAX_LEVELS + 1];  // first tile id of each level
};

__global__ void roi_bwd_geom_kernel(const __grid_constant__ RoiArgs a,
                                    const float* __restrict__ rois, int R,
                                    RoiBwdGeom* __restrict__ geom) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float* roi = rois + (size_t)r * 5;
  RoiBwdGeom o;
  if (roi[0] < 0.f) {
    o.g.b = -1; o.g.lvl = -1; o.g.H = o.g.W = o.g.gh = o.g.gw = 0;
    o.g.start_w = o.g.start_h = o.g.bin_w = o.g.bin_h = o.g.inv_count = 0.f;
    o.ylo = o.xlo = 1; o.yhi = o.xhi = 0;
    geom[r] = o;
    return;
  }
  o.g = roi_geometry(a, roi);
  const float y0 = o.g.start_h, y1 = o.g.start_h + o.g.bin_h * (float)a.PH;
  const float x0 = o.g.start_w, x1 = o.g.start_w + o.g.bin_w * (float)a.PW;
  // samples lie strictly inside (y0,y1); taps reach floor(y) and floor(y)+1
  bool ok = (o.g.gh > 0) && (o.g.gw > 0) && (y1 >= -1.0f) && (y0 <= (float)o.g.H) &&
            (x1 >= -1.0f) && (x0 <= (float)o.g.W) && (o.g.b >= 0) && (o.g.b < a.B);
  if (ok) {
    o.ylo = max(0, (int)floorf(fmaxf(y0, 0.f)));
    o.yhi = min(o.g.H - 1, (int)floorf(fminf(fmaxf(y1, 0.f), (float)o.g.H)) + 1);
    o.xlo = max(0, (int)floorf(fmaxf(x0, 0.f)));
    o.xhi = min(o.g.W - 1, (int)floorf(fminf(fmaxf(x1, 0.f), (float)o.g.W)) + 1);
  } else {
    o.ylo = o.xlo = 1; o.yhi = o.xhi = 0;
  }
  geom[r] = o;
}

// grid (total_tiles, C / BWD_CCH)
__global__ void __launch_bounds__(BWD_THREADS)
roi_bwd_gather_kernel(const __grid_constant__ RoiBwdArgs ba,
                      const RoiBwdGeom* __restrict__ geom, int R,
                      const float* __restrict__ gt /* (R, nbins, C) */) {
  __shared__ int s_list[BWD_LIST];
  __shared__ int s_warp[BWD_THREADS / 32];
  __shared__ int s_n;
  __shared__ float s_wy[BWD_TS][BWD_MAXP];
  __shared__ float s_wx[BWD_TS][BWD_MAXP];
  __shared__ int s_rowany[BWD_TS], s_colany[BWD_TS];

  const RoiArgs& a = ba.a;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // locate tile
  int lvl = 0;
  while (lvl + 1 < a.L && (int)blockIdx.x >= ba.tile_base[lvl + 1]) ++lvl;
  int t = blockIdx.x - ba.tile_base[lvl];
  const int tpi = ba.tiles_x[lvl] * ba.tiles_y[lvl];
  const int b = t / tpi; t -= b * tpi;
  const int ty = t / ba.tiles_x[lvl], tx = t - ty * ba.tiles_x[lvl];
  const int y0 = ty * BWD_TS, x0 = tx * BWD_TS;
  const int H = a.H[lvl], W = a.W[lvl], C = a.C;
  const int c0 = blockIdx.y * BWD_CCH;
  const int nbins = a.PH * a.PW;
  const int cq = tid & 15;    // channel quad within the slab
  const int pg = tid >> 4;    // pixel group 0..15; pixels pg + 16*k
  const bool cq_ok = (c0 + cq * 4) < C;

  float4 acc[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);

  int r_next = 0;
  while (r_next < R) {
    // ---- phase A: ordered list of RoIs touching this tile ----
    if (tid == 0) s_n = 0;
    __syncthreads();
    while (r_next < R) {
      const int n_now = s_n;
      if (n_now + BWD_THREADS > BWD_LIST) break;
      const int r = r_next + tid;
      bool f = false;
      if (r < R) {
        const RoiBwdGeom* q = geom + r;
        f = (q->g.lvl == lvl) && (q->g.b == b) && (q->ylo <= y0 + BWD_TS - 1) &&
            (q->yhi >= y0) && (q->xlo <= x0 + BWD_TS - 1) && (q->xhi >= x0);
      }
      const unsigned bm = __ballot_sync(0xffffffffu, f);
      if (lane == 0) s_warp[wid] = __popc(bm);
      __syncthreads();
      int wbase = 0, tot = 0;
#pragma unroll
      for (int w = 0; w < BWD_THREADS / 32; ++w) {
        const int c = s_warp[w];
        if (w < wid) wbase += c;
        tot += c;
      }
      if (f) s_list[n_now + wbase + __popc(bm & ((1u << lane) - 1u))] = r;
      __syncthreads();
      if (tid == 0) s_n = n_now + tot;
      r_next += BWD_THREADS;
      __syncthreads();
    }
    const int n = s_n;
    // ---- phase B: accumulate the listed RoIs in order ----
    for (int li = 0; li < n; ++li) {
      const int r = s_list[li];
      const RoiGeom g = geom[r].g;
      __syncthreads();  // previous iteration's table reads are done
      if (tid < BWD_TS * a.PH) {
        const int i = tid / a.PH, ph = tid - i * a.PH;
        const int y = y0 + i;
        float wsum = 0.f;
        const float ybase = g.start_h + (float)ph * g.bin_h;
        for (int iy = 0; iy < g.gh; ++iy) {
          const float yy = ybase + ((float)iy + 0.5f) * g.bin_h / (float)g.gh;
          int yl, yh; float hy, ly;
          if (!bilinear_axis(yy, H, yl, yh, hy, ly)) continue;
          if (yl == y) wsum += hy;
          if (yh == y) wsum += ly;
        }
        s_wy[i][ph] = wsum * g.inv_count;
      } else if (tid >= 128 && tid < 128 + BWD_TS * a.PW) {
        const int tt = tid - 128;
        const int i = tt / a.PW, pw = tt - i * a.PW;
        const int x = x0 + i;
        float wsum = 0.f;
        const float xbase = g.start_w + (float)pw * g.bin_w;
        for (int ix = 0; ix < g.gw; ++ix) {
          const float xx = xbase + ((float)ix + 0.5f) * g.bin_w / (float)g.gw;
          int xl, xh; float hx, lx;
          if (!bilinear_axis(xx, W, xl, xh, hx, lx)) continue;
          if (xl == x) wsum += hx;
          if (xh == x) wsum += lx;
        }
        s_wx[i][pw] = wsum;
      }
      __syncthreads();
      if (tid < BWD_TS) {
        int any = 0;
        for (int ph = 0; ph < a.PH; ++ph) any |= (s_wy[tid][ph] != 0.f);
        s_rowany[tid] = any;
      } else if (tid >= 32 && tid < 32 + BWD_TS) {
        const int i = tid - 32;
        int any = 0;
        for (int pw = 0; pw < a.PW; ++pw) any |= (s_wx[i][pw] != 0.f);
        s_colany[i] = any;
      }
      __syncthreads();
      if (!cq_ok) continue;
      const float* gr = gt + (size_t)r * nbins * C + c0 + cq * 4;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int p = pg + 16 * k;
        const int i = p >> 3, j = p & 7;
        if (!s_rowany[i] || !s_colany[j]) continue;
        for (int ph = 0; ph < a.PH; ++ph) {
          const float wy = s_wy[i][ph];
          if (wy == 0.f) continue;
          for (int pw = 0; pw < a.PW; ++pw) {
            const float wx = s_wx[j][pw];
            if (wx == 0.f) continue;
            const float w = wy * wx;
            const float4 v = __ldg(reinterpret_cast<const float4*>(
                gr + (size_t)(ph * a.PW + pw) * C));
            acc[k].x = fmaf(w, v.x, acc[k].x);
            acc[k].y = fmaf(w, v.y, acc[k].y);
            acc[k].z = fmaf(w, v.z, acc[k].z);
            acc[k].w = fmaf(w, v.w, acc[k].w);
          }
        }
      }
    }
    __syncthreads();
  }
  if (!cq_ok) return;
  float* gout = ba.grad[lvl] + (size_t)b * H * W * C + c0 + cq * 4;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int p = pg + 16 * k;
    const int y = y0 + (p >> 3), x = x0 + (p & 7);
    if (y < H && x < W)
      *reinterpret_cast<float4*>(gout + ((size_t)y * W + x) * C) = acc[k];
  }
}

struct RoiBwdWs {
  size_t geom, gt, total;
};
static inline RoiBwdWs roi_bwd_ws(const RoiArgs& a, int R) {
  RoiBwdWs w;
  size_t o = 0;
  const size_t nb = (size_t)a.PH * a.PW;
  w.geom = o; o = (o + (size_t)(R > 0 ? R : 1) * sizeof(RoiBwdGeom) + 255) & ~(size_t)255;
  w.gt = o;   o = (o + (size_t)(R > 0 ? R : 1) * nb * a.C * 4 + 255) & ~(size_t)255;
  w.total = o;
  return w;
}

static inline int roi_bwd_launch(RoiArgs a, const float* grad_out, const float* rois,
                                 int R, float* const* grad_feats, void* workspace,
                                 size_t workspace_bytes, cudaStream_t stream) {
  if (a.PH > BWD_MAXP || a.PW > BWD_MAXP || BWD_TS * a.PH > 128 || BWD_TS * a.PW > 128)
    return BRCNN_ERR_UNSUPPORTED;
  RoiBwdWs w = roi_bwd_ws(a, R);
  if (!workspace || workspace_bytes < w.total) return BRCNN_ERR_WORKSPACE;
  RoiBwdArgs ba;
  memset(&ba, 0, sizeof(ba));
  ba.a = a;
  int base = 0;
  for (int l = 0; l < a.L; ++l) {
    if (!grad_feats[l]) return BRCNN_ERR_ARG;
    ba.grad[l] = grad_feats[l];
    ba.tiles_x[l] = (a.W[l] + BWD_TS - 1) / BWD_TS;
    ba.tiles_y[l] = (a.H[l] + BWD_TS - 1) / BWD_TS;
    ba.tile_base[l] = base;
    base += ba.tiles_x[l] * ba.tiles_y[l] * a.B;
  }
  ba.tile_base[a.L] = base;
  char* ws = (char*)workspace;
  RoiBwdGeom* geom = (RoiBwdGeom*)(ws + w.geom);
  float* gt = (float*)(ws + w.gt);
  const int nbins = a.PH * a.PW;
  if (R > 0) {
    roi_bwd_geom_kernel<<<(R + 255) / 256, 256, 0, stream>>>(a, rois, R, geom);
    g_launch_count_add(1);
    BRCNN_CUDA_CHECK_LAST();
    const float* tin[1] = {grad_out};
    float* tout[1] = {gt};
    const int trows[1] = {a.C}, tcols[1] = {nbins};
    const int rc = transpose_launch_multi(tin, tout, 1, R, trows, tcols, stream);
    if (rc) return rc;
  }
  dim3 grid(base, (a.C + BWD_CCH - 1) / BWD_CCH);
  if (grid.y > 65535) return BRCNN_ERR_UNSUPPORTED;
  roi_bwd_gather_kernel<<<grid, BWD_THREADS, 0, stream>>>(ba, geom, R, gt);
  g_launch_count_add(1);
  BRCNN_CUDA_CHECK_LAST();
  return BRCNN_OK;
}

}  // namespace brcnn
