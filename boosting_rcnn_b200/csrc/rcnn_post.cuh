// K6: probabilistic score fusion + per-class delta decode + class-wise NMS.
// Reference: prob_roi_head.py:213-246 (fusion :232-240),
// convfc_bbox_head.py:294-330, bbox_nms.py:8-95.
//
//   rcnn_fuse_decode_kernel  warp per RoI row: softmax (pinned exp, class sum
//       in index order) * prior -> sqrt; every class box decoded against the
//       RoI, clipped, optionally / scale_factor; image-wide max coordinate of
//       the above-threshold candidates (the batched_nms offset base).
//   rcnn_class_sort_kernel   CTA per (image, class): gather candidates with
//       score > score_thr, bitonic sort by (score desc, flat index asc).
//   nms_mask / nms_sweep / nms_merge  (nms_kernels.cuh), segments = image x
//       class, offset id*(max+1) added in fp32 like mmcv.batched_nms.
#pragma once
#include "common.cuh"
#include "nms_kernels.cuh"

namespace brcnn {

struct RcnnArgs {
  int B, Rc, C, agnostic, prob, rescale;
  float means[4], stds[4];
  float max_ratio, score_thr;
};

constexpr int RCNN_FUSE_WARPS = 4;

__global__ void __launch_bounds__(RCNN_FUSE_WARPS * 32)
rcnn_fuse_decode_kernel(const __grid_constant__ RcnnArgs a,
                        const float* __restrict__ rois,
                        const float* __restrict__ prior,
                        const int32_t* __restrict__ num_rois,
                        const float* __restrict__ cls_score,
                        const float* __restrict__ bbox_pred,
                        const float* __restrict__ img_hw,
                        const float* __restrict__ scale_factor,
                        float* __restrict__ scores, float4* __restrict__ bboxes,
                        int* __restrict__ img_maxc_bits) {
  extern __shared__ float s_e[];  // [RCNN_FUSE_WARPS][C+1]
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int row = blockIdx.x * RCNN_FUSE_WARPS + wid;
  const int C = a.C, C1 = C + 1;
  if (row >= a.B * a.Rc) return;
  const int b = row / a.Rc, j = row - b * a.Rc;
  if (j >= num_rois[b]) return;
  float* e = s_e + wid * C1;
  const float* x = cls_score + (size_t)row * C1;
  float* srow = scores + (size_t)row * C1;

  if (a.prob) {
    float m = -INFINITY;
    for (int c = lane; c < C1; c += 32) m = fmaxf(m, x[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    for (int c = lane; c < C1; c += 32) e[c] = pinned_expf(x[c] - m);
    __syncwarp();
    float sum = 0.f;
    if (lane == 0)
      for (int c = 0; c < C1; ++c) sum += e[c];   // pinned order: c ascending
    sum = __shfl_sync(0xffffffffu, sum, 0);
    const float pr = prior[row];
    for (int c = lane; c < C1; c += 32) {
      const float p = e[c] / sum;
      const float f = sqrtf(p * pr);
      e[c] = f;
      srow[c] = f;
    }
  } else {
    for (int c = lane; c < C1; c += 32) { const float f = x[c]; e[c] = f; srow[c] = f; }
  }
  __syncwarp();

  const float* roi = rois + (size_t)row * 5;
  Box4 rb; rb.x1 = roi[1]; rb.y1 = roi[2]; rb.x2 = roi[3]; rb.y2 = roi[4];
  const float max_h = img_hw[b * 2], max_w = img_hw[b * 2 + 1];
  const float4 sf = a.rescale ? reinterpret_cast<const float4*>(scale_factor)[b]
                              : make_float4(1.f, 1.f, 1.f, 1.f);
  float local_max = 0.f;
  const int nbox = a.agnostic ? 1 : C;
  for (int c = lane; c < nbox; c += 32) {
    const float* d = bbox_pred + ((size_t)row * nbox + c) * 4;
    Box4 o = delta2bbox_one(rb, d[0], d[1], d[2], d[3], a.means, a.stds,
                            a.max_ratio, 1, max_w, max_h);
    if (a.rescale) { o.x1 = o.x1 / sf.x; o.y1 = o.y1 / sf.y; o.x2 = o.x2 / sf.z; o.y2 = o.y2 / sf.w; }
    bboxes[(size_t)row * nbox + c] = make_float4(o.x1, o.y1, o.x2, o.y2);
    const float bm = fmaxf(fmaxf(o.x1, o.y1), fmaxf(o.x2, o.y2));
    if (a.agnostic) {
      bool any = false;
      for (int cc = 0; cc < C; ++cc) any |= (e[cc] > a.score_thr);
      if (any) local_max = fmaxf(local_max, bm);
    } else if (e[c] > a.score_thr) {
      local_max = fmaxf(local_max, bm);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    local_max = fmaxf(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
  if (lane == 0 && local_max > 0.f) atomicMax(img_maxc_bits + b, __float_as_int(local_max));
}

// grid (C, B); dynamic smem: rpow2 u64
__global__ void __launch_bounds__(256)
rcnn_class_sort_kernel(const __grid_constant__ RcnnArgs a,
                       const int32_t* __restrict__ num_rois,
                       const float* __restrict__ scores,
                       const float4* __restrict__ bboxes, int rpow2,
                       u64* __restrict__ seg_key, float4* __restrict__ seg_boxes,
                       int32_t* __restrict__ seg_count) {
  extern __shared__ __align__(16) u64 s_keys[];
  __shared__ int s_n;
  const int c = blockIdx.x, b = blockIdx.y;
  const int C = a.C, C1 = C + 1;
  const int nr = min(num_rois[b], a.Rc);
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid == 0) s_n = 0;
  __syncthreads();
  const int nround = (nr + 31) & ~31;
  for (int j = tid; j < nround; j += blockDim.x) {
    float f = 0.f;
    bool pick = false;
    if (j < nr) {
      f = scores[((size_t)b * a.Rc + j) * C1 + c];
      pick = f > a.score_thr;
    }
    const unsigned bm = __ballot_sync(0xffffffffu, pick);
    if (bm) {
      int base = 0;
      if (lane == 0) base = atomicAdd(&s_n, __popc(bm));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (pick) {
        const uint32_t flat = (uint32_t)(j * C + c);
        s_keys[base + __popc(bm & ((1u << lane) - 1u))] =
            ((u64)__float_as_uint(f) << 32) | (u64)(0xFFFFFFFFu - flat);
      }
    }
  }
  __syncthreads();
  const int n = s_n;
  int np = 1;
  while (np < n) np <<= 1;
  for (int i = n + tid; i < np; i += blockDim.x) s_keys[i] = 0ull;
  if (n > 1) bitonic_sort_desc_u64(s_keys, np);
  else __syncthreads();
  const size_t seg = (size_t)b * C + c;
  const int nbox = a.agnostic ? 1 : C;
  for (int i = tid; i < n; i += blockDim.x) {
    const u64 k = s_keys[i];
    const uint32_t flat = 0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull);
    const int j = (int)(flat / (uint32_t)C);
    seg_key[seg * a.Rc + i] = k;
    seg_boxes[seg * a.Rc + i] =
        bboxes[((size_t)b * a.Rc + j) * nbox + (a.agnostic ? 0 : c)];
  }
  if (tid == 0) seg_count[seg] = n;
}

struct RcnnMergeEpilogue {
  static constexpr bool kNeedsPos = false;   // operator() uses (seg, pos)
  const float4* bboxes;  // [B*Rc][nbox]
  float* det_bboxes;     // (B, max_out, 5)
  int64_t* det_labels;   // (B, max_out)
  int Rc, C, nbox, max_out;
  __device__ void operator()(int b, int rank, int seg, int pos, u64 key) const {
    const uint32_t flat = 0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull);
    const int j = (int)(flat / (uint32_t)C), c = (int)(flat - (uint32_t)j * C);
    const float4 bx = bboxes[((size_t)b * Rc + j) * nbox + (nbox == 1 ? 0 : c)];
    float* o = det_bboxes + ((size_t)b * max_out + rank) * 5;
    o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w;
    o[4] = __uint_as_float((uint32_t)(key >> 32));
    det_labels[(size_t)b * max_out + rank] = c;
  }
  __device__ void pad(int b, int rank) const {
    float* o = det_bboxes + ((size_t)b * max_out + rank) * 5;
    o[0] = 0.f; o[1] = 0.f; o[2] = 0.f; o[3] = 0.f; o[4] = 0.f;
    det_labels[(size_t)b * max_out + rank] = -1;
  }
};

}  // namespace brcnn
