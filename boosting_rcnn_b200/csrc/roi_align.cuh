// K3 (v1 register-tile kernel, kept for pooled sizes > 7; roi_align_tma.cuh is the main
// path) + the shared geometry helpers, bbox2roi, level map and pyramid transposes.
// Fused FPN level mapping + multi-level RoIAlign forward (one launch for all levels),
// NHWC feature maps, NCHW (R,C,ph,pw) output.
//
// Reference behaviour: single_level_roi_extractor.py:36-115 + mmcv RoIAlign
// (aligned=True, pool_mode='avg', sampling_ratio=0), SURVEY.md App. A5/A6.
//
// Formulation.  mmcv averages gh x gw bilinear samples per bin; bilinear
// weights are separable, so for one RoI
//     out[c][ph][pw] = sum_y sum_x Wy[ph][y] * Wx[pw][x] * F[y][x][c]
// with Wy[ph][y] = (1/count) * sum over the bin's sample rows of the tap weight
// landing on feature row y (Wx likewise, without the 1/count).  Both tables
// are tiny (7 x footprint side) and banded; they are built once per RoI in
// shared memory.  The kernel then reads every footprint pixel of a bin row
// ONCE per bin row (not 4 taps x samples), as 128-bit channel-quad loads:
//     t[x]  = sum_{y in band(ph)} Wy[ph][y] * F[y][x]     (registers, 8 px chunk)
//     u[pw] += Wx[pw][x] * t[x]                            (7 accumulators)
// Mapping: one CTA per (RoI, <=256-channel slab); thread = (channel quad,
// bin-row slot), so a warp reads 512 contiguous bytes per pixel.  The slab's
// outputs are staged in shared memory in (c, bin) order and streamed out with
// coalesced 128-bit stores.
#pragma once
#include <cstring>

#include "common.cuh"

namespace brcnn {

extern int64_t g_launch_count_add(int n);

constexpr int ROI_THREADS = 256;
constexpr int ROI_XCH = 8;       // pixels per register chunk
constexpr int ROI_MAXP = 14;     // max pooled side (table / accumulator bound)

struct RoiArgs {
  const float* feat[BRCNN_MAX_LEVELS];  // NHWC (B,H,W,C)
  int H[BRCNN_MAX_LEVELS], W[BRCNN_MAX_LEVELS];
  float scale[BRCNN_MAX_LEVELS];
  int B, C, L, PH, PW, sampling_ratio, aligned;
  float finest_scale;
  int chunk_c;   // channels per CTA (multiple of 4, <= 256)
  int max_h, max_w;  // largest level map (table capacity)
};

struct RoiGeom {
  int b, lvl, H, W, gh, gw;
  float start_w, start_h, bin_w, bin_h, inv_count;
};

__device__ __forceinline__ RoiGeom roi_geometry(const RoiArgs& a,
                                                const float* __restrict__ roi) {
  RoiGeom g;
  g.b = (int)roi[0];
  const float x1 = roi[1], y1 = roi[2], x2 = roi[3], y2 = roi[4];
  g.lvl = map_roi_level(x1, y1, x2, y2, a.finest_scale, a.L);
  g.H = a.H[g.lvl];
  g.W = a.W[g.lvl];
  const float sc = a.scale[g.lvl];
  const float off = a.aligned ? 0.5f : 0.0f;
  g.start_w = x1 * sc - off;
  g.start_h = y1 * sc - off;
  const float end_w = x2 * sc - off;
  const float end_h = y2 * sc - off;
  float roi_w = end_w - g.start_w;
  float roi_h = end_h - g.start_h;
  if (!a.aligned) {
    roi_w = fmaxf(roi_w, 1.0f);
    roi_h = fmaxf(roi_h, 1.0f);
  }
  g.bin_h = roi_h / (float)a.PH;
  g.bin_w = roi_w / (float)a.PW;
  g.gh = a.sampling_ratio > 0 ? a.sampling_ratio : (int)ceilf(roi_h / (float)a.PH);
  g.gw = a.sampling_ratio > 0 ? a.sampling_ratio : (int)ceilf(roi_w / (float)a.PW);
  const float count = fmaxf((float)(g.gh * g.gw), 1.0f);
  g.inv_count = 1.0f / count;
  return g;
}

// One axis of mmcv's bilinear_interpolate: returns false when the sample is
// outside [-1, size]; otherwise low/high indices and weights.
__device__ __forceinline__ bool bilinear_axis(float v, int size, int& lo, int& hi,
                                              float& wl, float& wh) {
  if (v < -1.0f || v > (float)size) return false;
  if (v <= 0.f) v = 0.f;
  lo = (int)v;
  if (lo >= size - 1) {
    hi = lo = size - 1;
    v = (float)lo;
  } else {
    hi = lo + 1;
  }
  wh = v - (float)lo;  // weight of the high tap (ly / lx)
  wl = 1.0f - wh;      // weight of the low tap  (hy / hx)
  return true;
}

// Conservative inclusive range of feature rows (or columns) any sample of the
// RoI can touch along one axis; empty (lo > hi) when nothing is valid.
__device__ __forceinline__ void roi_axis_range(float start, float bin, int nbin, int grid,
                                               int size, int& lo, int& hi) {
  const float v0 = start, v1 = start + bin * (float)nbin;
  if (!(grid > 0) || !(v1 >= -1.0f) || !(v0 <= (float)size)) { lo = 1; hi = 0; return; }
  lo = max(0, (int)floorf(fmaxf(v0, 0.f)));
  hi = min(size - 1, (int)floorf(fminf(fmaxf(v1, 0.f), (float)size)) + 1);
}

// Weight of feature index `pos` for pooled bin `pb` along one axis:
// sum over the bin's `grid` samples of the tap weight that lands on `pos`.
__device__ __forceinline__ float roi_axis_weight(float start, float bin, int grid, int size,
                                                 int pb, int pos) {
  float w = 0.f;
  const float base = start + (float)pb * bin;
  for (int i = 0; i < grid; ++i) {
    const float v = base + ((float)i + 0.5f) * bin / (float)grid;
    int lo, hi; float wl, wh;
    if (!bilinear_axis(v, size, lo, hi, wl, wh)) continue;
    if (lo == pos) w += wl;
    if (hi == pos) w += wh;
  }
  return w;
}

// dynamic smem: chunk_c*nbins floats (stage) + PH*max_h + PW*max_w floats
// NPW: compile-time bound on pooled_w (number of register accumulators)
template <int NPW>
__global__ void __launch_bounds__(ROI_THREADS, NPW <= 7 ? 2 : 1)
roi_align_fwd_kernel(const __grid_constant__ RoiArgs a,
                     const float* __restrict__ rois, int R,
                     float* __restrict__ out, int32_t* __restrict__ roi_levels) {
  extern __shared__ __align__(16) float roi_smem[];
  const int nbins = a.PH * a.PW;
  float* stage = roi_smem;                       // [chunk_c][nbins]
  float* wy = stage + (size_t)a.chunk_c * nbins; // [PH][fh]   (includes 1/count)
  float* wx = wy + (size_t)a.PH * a.max_h;       // [fw][PW]   (x-major)
  __shared__ int s_ys[ROI_MAXP], s_ye[ROI_MAXP];  // row band of every bin row

  const int r = blockIdx.x;
  const int c0 = blockIdx.y * a.chunk_c;
  const int cc = min(a.chunk_c, a.C - c0);
  const float* roi = rois + (size_t)r * 5;
  float* dst = out + ((size_t)r * a.C + c0) * nbins;
  const int tid = threadIdx.x;
  const int total = cc * nbins;

  if (roi[0] < 0.f) {  // padding row
    for (int i = tid; i < total; i += ROI_THREADS) dst[i] = 0.f;
    if (roi_levels != nullptr && blockIdx.y == 0 && tid == 0) roi_levels[r] = -1;
    return;
  }
  const RoiGeom g = roi_geometry(a, roi);
  if (roi_levels != nullptr && blockIdx.y == 0 && tid == 0) roi_levels[r] = g.lvl;
  int ylo, yhi, xlo, xhi;
  roi_axis_range(g.start_h, g.bin_h, a.PH, g.gh, g.H, ylo, yhi);
  roi_axis_range(g.start_w, g.bin_w, a.PW, g.gw, g.W, xlo, xhi);
  const int fh = yhi - ylo + 1, fw = xhi - xlo + 1;
  // an image index outside [0, B) (hand-built rois, batch mismatch) reads nothing: zeros,
  // like the TMA kernels (block-uniform)
  const bool empty = (fh <= 0) || (fw <= 0) || g.b < 0 || g.b >= a.B;

  // ---- pull the whole footprint towards L2 in one shot ----
  // Every footprint row is a contiguous run of fw*C floats; issuing all the
  // prefetches up front turns the ~dozen dependent DRAM round trips of the
  // main loop into one DRAM latency plus L2 hits.
  if (!empty) {
    const char* fb = reinterpret_cast<const char*>(
        a.feat[g.lvl] + (size_t)g.b * g.H * g.W * a.C + c0);
    const int row_bytes = (fw - 1) * a.C * 4 + cc * 4;
    const int lines = (row_bytes + 127) >> 7;
    for (int i = tid; i < fh * lines; i += ROI_THREADS) {
      const int dy = i / lines, ln = i - dy * lines;
      const char* p = fb + ((size_t)(ylo + dy) * g.W + xlo) * a.C * 4 + (size_t)ln * 128;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
    }
  }

  // ---- separable weight tables ----
  if (!empty) {
    for (int i = tid; i < a.PH * fh; i += ROI_THREADS) {
      const int ph = i / fh, dy = i - ph * fh;
      wy[ph * fh + dy] =
          roi_axis_weight(g.start_h, g.bin_h, g.gh, g.H, ph, ylo + dy) * g.inv_count;
    }
    for (int i = tid; i < a.PW * fw; i += ROI_THREADS) {
      const int dx = i / a.PW, pw = i - dx * a.PW;
      wx[dx * a.PW + pw] = roi_axis_weight(g.start_w, g.bin_w, g.gw, g.W, pw, xlo + dx);
    }
  }
  __syncthreads();
  if (!empty && tid < a.PH) {  // non-zero row band of bin row `tid`
    int ys = fh, ye = -1;
    for (int dy = 0; dy < fh; ++dy)
      if (wy[tid * fh + dy] != 0.f) { ys = min(ys, dy); ye = dy; }
    s_ys[tid] = ys;
    s_ye[tid] = ye;
  }
  __syncthreads();

  // ---- main loop: thread = (channel quad, bin-row slot) ----
  const int ncq = cc >> 2;
  const int nslot = ROI_THREADS / 64;          // 4 slots of 64 quads
  const int cq = tid & 63, slot = tid >> 6;
  const int C = a.C;
  for (int cqb = 0; cqb < ncq; cqb += 64) {
    const int q = cqb + cq;
    const bool q_ok = q < ncq;
    const float* fbase = a.feat[g.lvl] + (size_t)g.b * g.H * g.W * C + c0 + q * 4;
    for (int ph = slot; ph < a.PH; ph += nslot) {
      float4 u[NPW];
#pragma unroll
      for (int pw = 0; pw < NPW; ++pw) u[pw] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!empty && q_ok) {
        const int ys = s_ys[ph], ye = s_ye[ph];
        for (int x0 = 0; x0 < fw; x0 += ROI_XCH) {
          float4 t[ROI_XCH];
#pragma unroll
          for (int i = 0; i < ROI_XCH; ++i) t[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int dy = ys; dy <= ye; ++dy) {
            const float w = wy[ph * fh + dy];
            const float* row = fbase + ((size_t)(ylo + dy) * g.W + xlo + x0) * C;
#pragma unroll
            for (int i = 0; i < ROI_XCH; ++i) {
              if (x0 + i < fw) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(row + (size_t)i * C));
                t[i].x = fmaf(w, v.x, t[i].x);
                t[i].y = fmaf(w, v.y, t[i].y);
                t[i].z = fmaf(w, v.z, t[i].z);
                t[i].w = fmaf(w, v.w, t[i].w);
              }
            }
          }
#pragma unroll
          for (int i = 0; i < ROI_XCH; ++i) {
            if (x0 + i < fw) {
              const float* wrow = wx + (x0 + i) * a.PW;
#pragma unroll
              for (int pw = 0; pw < NPW; ++pw) {
                if (pw < a.PW) {
                  const float w = wrow[pw];
                  if (w != 0.f) {
                    u[pw].x = fmaf(w, t[i].x, u[pw].x);
                    u[pw].y = fmaf(w, t[i].y, u[pw].y);
                    u[pw].z = fmaf(w, t[i].z, u[pw].z);
                    u[pw].w = fmaf(w, t[i].w, u[pw].w);
                  }
                }
              }
            }
          }
        }
      }
      if (q_ok) {
        float* s = stage + (size_t)(q * 4) * nbins + ph * a.PW;
#pragma unroll
        for (int pw = 0; pw < NPW; ++pw) {
          if (pw < a.PW) {
            s[pw] = u[pw].x;
            s[nbins + pw] = u[pw].y;
            s[2 * nbins + pw] = u[pw].z;
            s[3 * nbins + pw] = u[pw].w;
          }
        }
      }
    }
  }
  __syncthreads();
  if ((total & 3) == 0 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
    const float4* s4 = reinterpret_cast<const float4*>(stage);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i = tid; i < (total >> 2); i += ROI_THREADS) __stcs(d4 + i, s4[i]);
  } else {
    for (int i = tid; i < total; i += ROI_THREADS) dst[i] = stage[i];
  }
}

// map_roi_levels as a standalone operator (int64 output like the reference)
__global__ void map_roi_levels_kernel(const float* __restrict__ rois, int R,
                                      float finest_scale, int L,
                                      int64_t* __restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float* roi = rois + (size_t)r * 5;
  out[r] = (int64_t)map_roi_level(roi[1], roi[2], roi[3], roi[4], finest_scale, L);
}

// bbox2roi (mmdet/core/bbox/transforms.py:59-78) on the padded proposal layout
// of brcnn_rpn_get_bboxes: (B,cap,5)[x1,y1,x2,y2,score] + num[B] ->
// rois (B*cap,5)[b,x1,y1,x2,y2] (b = -1 on padding rows) and prior (B*cap)
// = proposals[:, -1] (prob_roi_head.py:214).
__global__ void bbox2roi_padded_kernel(const float* __restrict__ props,
                                       const int32_t* __restrict__ num, int B, int cap,
                                       float* __restrict__ rois, float* __restrict__ prior) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * cap) return;
  const int b = i / cap, j = i - b * cap;
  const float* p = props + (size_t)i * 5;
  float* r = rois + (size_t)i * 5;
  const bool live = j < num[b];
  r[0] = live ? (float)b : -1.0f;
  r[1] = p[0]; r[2] = p[1]; r[3] = p[2]; r[4] = p[3];
  if (prior != nullptr) prior[i] = p[4];
}

// (B, rows, cols) -> (B, cols, rows) fp32 tile transpose of up to
// BRCNN_MAX_LEVELS maps in ONE launch.  NCHW->NHWC is rows=C, cols=HW;
// NHWC->NCHW is rows=HW, cols=C.  64x64 tiles, 256 threads: 128-bit global
// loads along cols and 128-bit global stores along rows whenever the map's
// geometry allows it (cols % 4 == 0 / rows % 4 == 0), scalar otherwise; the
// tile is staged in shared memory with a 65-float pitch (<= 2-way conflicts).
// 16 KB of loads in flight per CTA, up to 8 CTAs per SM.
constexpr int TR_TILE = 64;
constexpr int TR_PITCH = 65;

struct TransposeMap {
  const float* in;
  float* out;
  int rows, cols;
  int tiles_x, tiles_y;  // tiles along cols / rows
  int tile_base;         // first CTA of this map
  int pad;
};
struct TransposeArgs {
  TransposeMap m[BRCNN_MAX_LEVELS];
  int n, batch;
};

__global__ void __launch_bounds__(256)
transpose_multi_kernel(const __grid_constant__ TransposeArgs a) {
  __shared__ float tile[TR_TILE * TR_PITCH];
  int mi = 0;
  while (mi + 1 < a.n && (int)blockIdx.x >= a.m[mi + 1].tile_base) ++mi;
  const TransposeMap& m = a.m[mi];
  int t = blockIdx.x - m.tile_base;
  const int tpb = m.tiles_x * m.tiles_y;
  const int b = t / tpb; t -= b * tpb;
  const int tyi = t / m.tiles_x, txi = t - tyi * m.tiles_x;
  const int r0 = tyi * TR_TILE, c0 = txi * TR_TILE;
  const int rows = m.rows, cols = m.cols;
  const size_t boff = (size_t)b * rows * cols;
  const float* __restrict__ in = m.in + boff;
  float* __restrict__ out = m.out + boff;
  const int q = threadIdx.x & 15, s = threadIdx.x >> 4;
  const bool vin = ((cols & 3) == 0) && ((reinterpret_cast<uintptr_t>(m.in) & 15) == 0);
  const bool vout = ((rows & 3) == 0) && ((reinterpret_cast<uintptr_t>(m.out) & 15) == 0);
  // ---- load: thread (s, q) reads rows r0+s+16k, cols c0+4q..4q+3 ----
  float4 v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = r0 + s + 16 * k, c = c0 + 4 * q;
    v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < rows) {
      const float* p = in + (size_t)r * cols + c;
      if (vin) {
        if (c < cols) v[k] = __ldcs(reinterpret_cast<const float4*>(p));
      } else {
        if (c < cols) v[k].x = __ldcs(p);
        if (c + 1 < cols) v[k].y = __ldcs(p + 1);
        if (c + 2 < cols) v[k].z = __ldcs(p + 2);
        if (c + 3 < cols) v[k].w = __ldcs(p + 3);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float* d = tile + (s + 16 * k) * TR_PITCH + 4 * q;
    d[0] = v[k].x; d[1] = v[k].y; d[2] = v[k].z; d[3] = v[k].w;
  }
  __syncthreads();
  // ---- store: thread (s, q) writes out rows (= in cols) c0+s+16k, in-rows r0+4q.. ----
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + s + 16 * k, r = r0 + 4 * q;
    if (c >= cols || r >= rows) continue;
    const float* sp = tile + (4 * q) * TR_PITCH + s + 16 * k;
    const float4 o = make_float4(sp[0], sp[TR_PITCH], sp[2 * TR_PITCH], sp[3 * TR_PITCH]);
    float* p = out + (size_t)c * rows + r;
    if (vout) {
      *reinterpret_cast<float4*>(p) = o;
    } else {
      p[0] = o.x;
      if (r + 1 < rows) p[1] = o.y;
      if (r + 2 < rows) p[2] = o.z;
      if (r + 3 < rows) p[3] = o.w;
    }
  }
}

// host launcher: n maps sharing `batch`; map i is (batch, rows[i], cols[i])
inline int transpose_launch_multi(const float* const* in, float* const* out, int n,
                                  int batch, const int* rows, const int* cols,
                                  cudaStream_t stream) {
  if (n <= 0 || n > BRCNN_MAX_LEVELS || batch <= 0 || !in || !out) return BRCNN_ERR_ARG;
  TransposeArgs a;
  memset(&a, 0, sizeof(a));
  a.n = n; a.batch = batch;
  long long base = 0;
  for (int i = 0; i < n; ++i) {
    if (!in[i] || !out[i] || rows[i] <= 0 || cols[i] <= 0) return BRCNN_ERR_ARG;
    TransposeMap& m = a.m[i];
    m.in = in[i]; m.out = out[i]; m.rows = rows[i]; m.cols = cols[i];
    m.tiles_x = (cols[i] + TR_TILE - 1) / TR_TILE;
    m.tiles_y = (rows[i] + TR_TILE - 1) / TR_TILE;
    m.tile_base = (int)base;
    base += (long long)m.tiles_x * m.tiles_y * batch;
    if (base > 0x7fffffffLL) return BRCNN_ERR_UNSUPPORTED;
  }
  transpose_multi_kernel<<<(unsigned)base, 256, 0, stream>>>(a);
  g_launch_count_add(1);
  BRCNN_CUDA_CHECK_LAST();
  return BRCNN_OK;
}

}  // namespace brcnn
