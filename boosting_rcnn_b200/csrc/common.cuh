// Shared device helpers for libbrcnn (sm_100a only).
//
// Arithmetic policy (see DESIGN.md "Pinned arithmetic"): the library is built
// with -fmad=false so that every `a*b+c` written below is two separately
// rounded IEEE operations, exactly like the chained torch ops of the
// reference (e.g. mmdet/core/bbox/coder/delta_xywh_bbox_coder.py:210,225-247).
// Where a fused multiply-add is wanted it is spelled fmaf(). exp() is the
// pinned polynomial below (pinned_expf) so that CPU oracle and GPU agree bit
// for bit; division and sqrtf are IEEE-correct in CUDA by default.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define BRCNN_OK 0
#define BRCNN_ERR_ARG (-1)
#define BRCNN_ERR_WORKSPACE (-2)
#define BRCNN_ERR_UNSUPPORTED (-3)

#define BRCNN_MAX_LEVELS 8
#define BRCNN_MAX_ANCHORS 32

#define BRCNN_CUDA_CHECK_LAST()                        \
  do {                                                 \
    cudaError_t _e = cudaGetLastError();               \
    if (_e != cudaSuccess) return (int)_e;             \
  } while (0)

namespace brcnn {

// cached cudaFuncSetAttribute(MaxDynamicSharedMemorySize) — defined in libbrcnn.cu
cudaError_t ensure_dyn_smem(const void* fn, size_t bytes, bool max_carveout = false);

// ---------------------------------------------------------------------------
// pinned_expf: Cephes-style single precision exp, written with explicit
// fmaf so that gcc (-ffp-contract=off) and nvcc (-fmad=false) produce the same
// bits.  Accuracy ~1 ulp over the whole range.  The identical routine lives
// in oracle/brcnn_oracle.c (oracle_expf).
// ---------------------------------------------------------------------------
__device__ __forceinline__ float pinned_expf(float x) {
  if (x != x) return x;                       // NaN
  if (x > 88.72283905206835f) return __int_as_float(0x7f800000);   // +inf
  if (x < -103.972076416f) return 0.0f;
  const float LOG2E = 1.44269504088896341f;
  const float C1 = 0.693359375f;
  const float C2 = -2.12194440e-4f;
  float t = x * LOG2E;
  int n = __float2int_rn(t);
  float fn = (float)n;
  float r = fmaf(fn, -C1, x);
  r = fmaf(fn, -C2, r);
  float p = 1.9875691500E-4f;
  p = fmaf(p, r, 1.3981999507E-3f);
  p = fmaf(p, r, 8.3334519073E-3f);
  p = fmaf(p, r, 4.1665795894E-2f);
  p = fmaf(p, r, 1.6666665459E-1f);
  p = fmaf(p, r, 5.0000001201E-1f);
  float r2 = r * r;
  float y = fmaf(p, r2, r);
  y = y + 1.0f;
  // scale by 2^n in two exact steps (handles the denormal tail)
  int n1 = n / 2;
  int n2 = n - n1;
  float s1 = __int_as_float((n1 + 127) << 23);
  float s2 = __int_as_float((n2 + 127) << 23);
  return (y * s1) * s2;
}

// sigmoid as 1/(1+exp(-x)) (ATSSRPNHead scores: atss_rpn_head.py:712-716).
__device__ __forceinline__ float pinned_sigmoid(float x) {
  return 1.0f / (1.0f + pinned_expf(-x));
}

// ---------------------------------------------------------------------------
// delta2bbox for one box (delta_xywh_bbox_coder.py:206-270), op for op.
// max_ratio = |log(wh_ratio_clip)| is computed by the host in float64 and
// rounded to fp32 exactly like `dw.clamp(min=-max_ratio, max=max_ratio)`.
// ---------------------------------------------------------------------------
struct Box4 {
  float x1, y1, x2, y2;
};

__device__ __forceinline__ float clampf(float v, float lo, float hi) {
  // torch.clamp semantics: NaN propagates
  if (v != v) return v;
  return v < lo ? lo : (v > hi ? hi : v);
}

__device__ __forceinline__ Box4 delta2bbox_one(
    Box4 roi, float d0, float d1, float d2, float d3, const float* means,
    const float* stds, float max_ratio, int clip, float max_w, float max_h) {
  float dx = d0 * stds[0] + means[0];
  float dy = d1 * stds[1] + means[1];
  float dw = d2 * stds[2] + means[2];
  float dh = d3 * stds[3] + means[3];
  float px = (roi.x1 + roi.x2) * 0.5f;
  float py = (roi.y1 + roi.y2) * 0.5f;
  float pw = roi.x2 - roi.x1;
  float ph = roi.y2 - roi.y1;
  float dx_width = pw * dx;
  float dy_height = ph * dy;
  dw = clampf(dw, -max_ratio, max_ratio);
  dh = clampf(dh, -max_ratio, max_ratio);
  float gw = pw * pinned_expf(dw);
  float gh = ph * pinned_expf(dh);
  float gx = px + dx_width;
  float gy = py + dy_height;
  Box4 o;
  o.x1 = gx - gw * 0.5f;
  o.y1 = gy - gh * 0.5f;
  o.x2 = gx + gw * 0.5f;
  o.y2 = gy + gh * 0.5f;
  if (clip) {
    // torch.where(b < 0, 0, b); torch.where(b > max, max, b): NaN stays NaN
    o.x1 = o.x1 < 0.f ? 0.f : o.x1;
    o.y1 = o.y1 < 0.f ? 0.f : o.y1;
    o.x2 = o.x2 < 0.f ? 0.f : o.x2;
    o.y2 = o.y2 < 0.f ? 0.f : o.y2;
    o.x1 = o.x1 > max_w ? max_w : o.x1;
    o.y1 = o.y1 > max_h ? max_h : o.y1;
    o.x2 = o.x2 > max_w ? max_w : o.x2;
    o.y2 = o.y2 > max_h ? max_h : o.y2;
  }
  return o;
}

// ---------------------------------------------------------------------------
// mmcv nms_cpu suppression test (division form, strict >), offset = 0:
//   inter / (area_i + area_j - inter) > thr       
// ---------------------------------------------------------------------------
__device__ __forceinline__ float box_area(float4 b) {
  return (b.z - b.x) * (b.w - b.y);
}

__device__ __forceinline__ bool iou_suppresses(float4 a, float area_a,
                                                        float4 b, float area_b,
                                                        float thr) {
  float xx1 = a.x > b.x ? a.x : b.x;
  float yy1 = a.y > b.y ? a.y : b.y;
  float xx2 = a.z < b.z ? a.z : b.z;
  float yy2 = a.w < b.w ? a.w : b.w;
  float w = xx2 - xx1;
  float h = yy2 - yy1;
  w = w > 0.f ? w : 0.f;
  h = h > 0.f ? h : 0.f;
  float inter = w * h;
  float ovr = inter / (area_a + area_b - inter);
  return ovr > thr;
}

// ---------------------------------------------------------------------------
// FPN level of one RoI (single_level_roi_extractor.py:51-54):
//   floor(log2(sqrt(w*h)/finest + 1e-6)).clamp(0, L-1)
// evaluated by comparing the fp32 argument of log2 with 2,4,8,...
// ---------------------------------------------------------------------------
__device__ __forceinline__ int map_roi_level(float x1, float y1,
                                                      float x2, float y2,
                                                      float finest_scale,
                                                      int num_levels) {
  float scale = sqrtf((x2 - x1) * (y2 - y1));
  float v = scale / finest_scale + 1e-6f;
  // floor(log2f(v)) >= k  <=>  v >= T[k], with log2f rounded to nearest: near
  // k >= 3 the fp32 grid of the RESULT is coarser than log2's slope, so the
  // 1 (k = 3,4) or 2 (k >= 5) floats just below 2^k already round up to k.
  // The table holds exactly those thresholds (checked against torch.log2 on a
  // +-50-float sweep around every power of two, tests/test_oracle_golden.py).
  const uint32_t T[8] = {0u,          0x40000000u, 0x40800000u, 0x40FFFFFFu,
                         0x417FFFFFu, 0x41FFFFFEu, 0x427FFFFEu, 0x42FFFFFEu};
  int lvl = 0;
  // NaN compares false everywhere -> level 0 (torch: floor(nan).clamp.long()
  // is implementation defined; documented in DESIGN.md)
  while (lvl < num_levels - 1 && lvl < 7 && v >= __uint_as_float(T[lvl + 1])) ++lvl;
  return lvl;
}

#ifdef __CUDACC__
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned warp_id() { return threadIdx.x >> 5; }

// In-place bitonic sort (descending) of n_pow2 u64 keys in shared memory by
// the whole CTA.  n_pow2 must be a power of two; pad with 0.
//
// Hybrid network: every warp owns 64-key blocks (lane holds keys lane and
// lane+32 of the block); all compare-exchange stages with partner distance
// j <= 32 run in registers (j == 32 inside the lane, j < 32 through warp
// shuffles) without any block barrier, only the stages with j >= 64 go through
// shared memory.  For 2048 keys that is 21 block barriers instead of 66.
__device__ __forceinline__ void bitonic_cmpx_lane(unsigned long long& x, int gi, int k, int j) {
  const unsigned long long y = __shfl_xor_sync(0xffffffffu, x, j);
  const bool desc = ((gi & k) == 0);
  const bool lower = ((gi & j) == 0);
  const bool take_max = (desc == lower);
  const unsigned long long mx = x > y ? x : y, mn = x > y ? y : x;
  x = take_max ? mx : mn;
}

__device__ __forceinline__ void bitonic_block_tail(unsigned long long& x0, unsigned long long& x1,
                                                   int gi0, int k) {
  // j = 32: the pair lives in one lane
  {
    const bool desc = ((gi0 & k) == 0);
    const unsigned long long mx = x0 > x1 ? x0 : x1, mn = x0 > x1 ? x1 : x0;
    x0 = desc ? mx : mn;
    x1 = desc ? mn : mx;
  }
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) {
    bitonic_cmpx_lane(x0, gi0, k, j);
    bitonic_cmpx_lane(x1, gi0 + 32, k, j);
  }
}

__device__ __forceinline__ void bitonic_sort_desc_u64(unsigned long long* keys,
                                                      int n_pow2) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  if (n_pow2 < 64) {   // tiny: plain shared-memory network
    for (int k = 2; k <= n_pow2; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        __syncthreads();
        for (int t = threadIdx.x; t < (n_pow2 >> 1); t += blockDim.x) {
          int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
          int ixj = i + j;
          unsigned long long a = keys[i];
          unsigned long long b = keys[ixj];
          bool desc = ((i & k) == 0);
          bool swap = desc ? (a < b) : (a > b);
          if (swap) {
            keys[i] = b;
            keys[ixj] = a;
          }
        }
      }
    }
    __syncthreads();
    return;
  }
  const int nblk = n_pow2 >> 6;
  __syncthreads();
  // merge sizes 2..64: entirely inside the 64-key blocks
  for (int blk = warp; blk < nblk; blk += nwarps) {
    const int gi0 = blk * 64 + lane;
    unsigned long long x0 = keys[gi0], x1 = keys[gi0 + 32];
    for (int k = 2; k <= 32; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        bitonic_cmpx_lane(x0, gi0, k, j);
        bitonic_cmpx_lane(x1, gi0 + 32, k, j);
      }
    }
    bitonic_block_tail(x0, x1, gi0, 64);
    keys[gi0] = x0;
    keys[gi0 + 32] = x1;
  }
  for (int k = 128; k <= n_pow2; k <<= 1) {
    for (int j = k >> 1; j >= 64; j >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (n_pow2 >> 1); t += blockDim.x) {
        // index of the lower element of the t-th compare-exchange pair
        // (j is a power of two: t/j*2j + t%j without integer division)
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int ixj = i + j;
        const unsigned long long a = keys[i];
        const unsigned long long b = keys[ixj];
        const bool desc = ((i & k) == 0);
        const bool swap = desc ? (a < b) : (a > b);
        if (swap) {
          keys[i] = b;
          keys[ixj] = a;
        }
      }
    }
    __syncthreads();
    for (int blk = warp; blk < nblk; blk += nwarps) {
      const int gi0 = blk * 64 + lane;
      unsigned long long x0 = keys[gi0], x1 = keys[gi0 + 32];
      bitonic_block_tail(x0, x1, gi0, k);
      keys[gi0] = x0;
      keys[gi0 + 32] = x1;
    }
  }
  __syncthreads();
}
#endif

}  // namespace brcnn
