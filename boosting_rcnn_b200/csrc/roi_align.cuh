// K3: fused FPN level mapping + multi-level RoIAlign forward (one launch for
// all levels), NHWC feature maps, NCHW (R,C,ph,pw) output.
//
// Reference behaviour: single_level_roi_extractor.py:36-115 + mmcv RoIAlign
// (aligned=True, pool_mode='avg', sampling_ratio=0), SURVEY.md App. A5/A6.
//
// Mapping: one CTA per (RoI, channel chunk <= 256).  A warp covers 8 channel
// quads x 4 bins: every tap is a 128-byte line read as 8 x LDG.128, the four
// taps of a sample come from at most four lines.  The chunk's ph*pw*C outputs
// are staged in shared memory in (c, bin) order (bank-conflict free for 7x7)
// and streamed out as 128-bit coalesced stores.
#pragma once
#include "common.cuh"

namespace brcnn {

struct RoiArgs {
  const float* feat[BRCNN_MAX_LEVELS];  // NHWC (B,H,W,C)
  int H[BRCNN_MAX_LEVELS], W[BRCNN_MAX_LEVELS];
  float scale[BRCNN_MAX_LEVELS];
  int B, C, L, PH, PW, sampling_ratio, aligned;
  float finest_scale;
  int chunk_c;  // channels per CTA (multiple of 32)
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

__global__ void __launch_bounds__(256)
roi_align_fwd_kernel(const __grid_constant__ RoiArgs a,
                     const float* __restrict__ rois, int R,
                     float* __restrict__ out, int32_t* __restrict__ roi_levels) {
  extern __shared__ __align__(16) float stage[];  // [chunk_c][nbins]
  const int r = blockIdx.x;
  const int c0 = blockIdx.y * a.chunk_c;
  const int cc = min(a.chunk_c, a.C - c0);
  const int nbins = a.PH * a.PW;
  const float* roi = rois + (size_t)r * 5;
  float* dst = out + ((size_t)r * a.C + c0) * nbins;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int total = cc * nbins;

  if (roi[0] < 0.f) {  // padding row
    for (int i = tid; i < total; i += blockDim.x) dst[i] = 0.f;
    if (roi_levels != nullptr && blockIdx.y == 0 && tid == 0) roi_levels[r] = -1;
    return;
  }
  const RoiGeom g = roi_geometry(a, roi);
  if (roi_levels != nullptr && blockIdx.y == 0 && tid == 0) roi_levels[r] = g.lvl;
  const float* feat = a.feat[g.lvl] + (size_t)g.b * g.H * g.W * a.C + c0;
  const int C = a.C;

  const int ncg = cc >> 2;             // channel quads in this chunk
  const int ncgt = (ncg + 7) >> 3;     // tiles of 8 quads
  const int nbt = (nbins + 3) >> 2;    // tiles of 4 bins
  const int cg_sub = lane & 7, bin_sub = lane >> 3;
  const int nwarps = blockDim.x >> 5;
  for (int it = wid; it < ncgt * nbt; it += nwarps) {
    const int bt = it / ncgt, cgt = it - bt * ncgt;
    const int bin = bt * 4 + bin_sub;
    const int cg = cgt * 8 + cg_sub;
    if (bin >= nbins || cg >= ncg) continue;
    const int ph = bin / a.PW, pw = bin - ph * a.PW;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float ybase = g.start_h + (float)ph * g.bin_h;
    const float xbase = g.start_w + (float)pw * g.bin_w;
    for (int iy = 0; iy < g.gh; ++iy) {
      const float y = ybase + ((float)iy + 0.5f) * g.bin_h / (float)g.gh;
      int yl, yh; float hy, ly;
      if (!bilinear_axis(y, g.H, yl, yh, hy, ly)) continue;
      const float* rowl = feat + (size_t)yl * g.W * C + cg * 4;
      const float* rowh = feat + (size_t)yh * g.W * C + cg * 4;
      for (int ix = 0; ix < g.gw; ++ix) {
        const float x = xbase + ((float)ix + 0.5f) * g.bin_w / (float)g.gw;
        int xl, xh; float hx, lx;
        if (!bilinear_axis(x, g.W, xl, xh, hx, lx)) continue;
        const float4 v1 = __ldg(reinterpret_cast<const float4*>(rowl + (size_t)xl * C));
        const float4 v2 = __ldg(reinterpret_cast<const float4*>(rowl + (size_t)xh * C));
        const float4 v3 = __ldg(reinterpret_cast<const float4*>(rowh + (size_t)xl * C));
        const float4 v4 = __ldg(reinterpret_cast<const float4*>(rowh + (size_t)xh * C));
        const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
        acc.x = fmaf(w1, v1.x, fmaf(w2, v2.x, fmaf(w3, v3.x, fmaf(w4, v4.x, acc.x))));
        acc.y = fmaf(w1, v1.y, fmaf(w2, v2.y, fmaf(w3, v3.y, fmaf(w4, v4.y, acc.y))));
        acc.z = fmaf(w1, v1.z, fmaf(w2, v2.z, fmaf(w3, v3.z, fmaf(w4, v4.z, acc.z))));
        acc.w = fmaf(w1, v1.w, fmaf(w2, v2.w, fmaf(w3, v3.w, fmaf(w4, v4.w, acc.w))));
      }
    }
    float* s = stage + (size_t)(cg * 4) * nbins + bin;
    s[0] = acc.x * g.inv_count;
    s[nbins] = acc.y * g.inv_count;
    s[2 * nbins] = acc.z * g.inv_count;
    s[3 * nbins] = acc.w * g.inv_count;
  }
  __syncthreads();
  if ((total & 3) == 0 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
    const float4* s4 = reinterpret_cast<const float4*>(stage);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i = tid; i < (total >> 2); i += blockDim.x) __stcs(d4 + i, s4[i]);
  } else {
    for (int i = tid; i < total; i += blockDim.x) dst[i] = stage[i];
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

// (B, rows, cols) -> (B, cols, rows) fp32 tile transpose.  NCHW->NHWC is
// rows=C, cols=HW; NHWC->NCHW is rows=HW, cols=C.
__global__ void __launch_bounds__(256)
transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int rows,
                 int cols) {
  __shared__ float tile[32][33];
  const size_t boff = (size_t)blockIdx.z * rows * cols;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int r = r0 + ty + k, c = c0 + tx;
    if (r < rows && c < cols) tile[ty + k][tx] = in[boff + (size_t)r * cols + c];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int c = c0 + ty + k, r = r0 + tx;
    if (r < rows && c < cols) out[boff + (size_t)c * rows + r] = tile[tx][ty + k];
  }
}

}  // namespace brcnn
