// RPN loss path: anchor targets + classification / regression / IoU
// losses of ATSSRPNHead with atss=False, forward value and gradients in three launches.
//
// Reference (file:line relative to the reference tree):
//   ATSSRPNHead.loss / loss_single / get_targets  mmdet/models/dense_heads/atss_rpn_head.py:299-464,505-603
//   AnchorHead.get_anchors / _get_targets_single  mmdet/models/dense_heads/anchor_head.py:126-265
//   AnchorGenerator.grid_anchors / valid_flags    mmdet/core/anchor/anchor_generator.py:338-434
//   MaxIoUAssigner (match_low_quality=True, gt_max_assign_all=True)  assigners/max_iou_assigner.py:61-212
//   PseudoSampler                                 samplers/pseudo_sampler.py:24-42
//   bbox_overlaps (eps 1e-6)                      iou_calculators/iou2d_calculator.py:75-260
//   delta2bbox / bbox2delta                       coder/delta_xywh_bbox_coder.py:98-272
//   FocalLoss (py_sigmoid_focal_loss == mmcv sigmoid_focal_loss)  losses/focal_loss.py:13-58,86
//   IoULoss / iou_loss (mode 'log', eps 1e-6)     losses/iou_loss.py:15-52,457-535
//   MSELoss                                       losses/mse_loss.py:9-57
//   CrossEntropyLoss(use_sigmoid) / binary_cross_entropy  losses/cross_entropy_loss.py:73-111
//   weight_reduce_loss                            losses/utils.py:28-55
//
//   rpn_loss_gtmax_kernel   per (image, GT): max IoU over the image's valid anchors
//                           (gt_max_overlaps, :170): block-level max in shared memory, one
//                           global atomicMax on the IoU's bit pattern per (block, GT).
//   rpn_loss_main_kernel    one thread per anchor, plane order (coalesced reads of the NCHW head
//                           outputs): re-evaluates the anchor's IoUs against the image's GTs
//                           (bit-identical to the first pass), applies the assigner rules incl.
//                           the sequential `overlaps[i] == gt_max_overlaps[i]` override of
//                           match_low_quality, then focal loss + gradient for every weighted
//                           anchor and, for positives, decode -> aligned IoU -> IoU-log loss,
//                           MSE "aug" loss on the encoded target, BCE on the IoU logit, with
//                           hand-derived gradients (torch.max / clamp tie rules included).
//                           Un-normalised gradients are written for EVERY element (zeros where
//                           nothing flows); 5 sums per block go to a partials array.
//   rpn_loss_reduce_kernel  fixed-order reduction of the partials -> per-level sums + the two
//                           normalisers (num_total_pos, sum of iou_target).  The division by
//                           reduce_mean(...) (atss_rpn_head.py:441-444,458-460) happens on the
//                           device after ONE fused all-reduce; no .item() host sync.
//   rpn_loss_scale_kernel   backward: raw gradients x (upstream / normaliser) per (loss, level).
#pragma once
#include <cstring>

#include "common.cuh"
#include "rcnn_train_prep.cuh"

namespace brcnn {

extern int64_t g_launch_count_add(int n);

constexpr int RL_THREADS = 256;
constexpr int RL_SUMS = 5;     // per block: cls, bbox (0.5*(iou-log + aug)), bce, iou_target, num_pos

struct RpnLossLevel {
  const float* cls;    // (B, A, H, W)
  const float* bbox;   // (B, 4A, H, W)
  const float* iou;    // (B, A, H, W)
  float* g_cls;        // same shapes, un-normalised gradients
  float* g_bbox;
  float* g_iou;
  int H, W, stride_w, stride_h;
  int n;               // H * W * A
  int block_base;      // first block (x) of this level
};

struct RpnLossArgs {
  RpnLossLevel lv[BRCNN_MAX_LEVELS];
  int A, L, B, Gmax, blocks_per_img;
  float pos_iou_thr, neg_iou_thr, min_pos_iou;
  float gamma;                   // bbox weight = iou_target ** gamma
  float focal_gamma, focal_alpha;
  int cls_loss_type;             // 0 sigmoid focal loss, 1 varifocal loss (iou_weighted)
  float w_cls, w_bbox, w_iou, w_aug;
  float max_ratio;
};

__device__ __forceinline__ int rl_level_of_block(const RpnLossArgs& a, int bx) {
  int l = 0;
  while (l + 1 < a.L && bx >= a.lv[l + 1].block_base) ++l;
  return l;
}

// anchor of plane-order element t of level lv: a = t / (H*W), (y, x) = t % (H*W)
__device__ __forceinline__ float4 rl_anchor(const RpnLossLevel& lv, const float* __restrict__ base,
                                            int A, int l, int a, int y, int x) {
  const float4 b = reinterpret_cast<const float4*>(base)[l * A + a];
  const float sx = (float)(x * lv.stride_w), sy = (float)(y * lv.stride_h);
  return make_float4(b.x + sx, b.y + sy, b.z + sx, b.w + sy);
}

__device__ __forceinline__ bool rl_valid(const RpnLossLevel& lv, const float* __restrict__ pad_hw,
                                         int b, int y, int x) {
  // AnchorGenerator.valid_flags: valid_h = min(ceil(pad_h / stride_h), H)
  const int vh = min((int)ceilf(pad_hw[b * 2 + 0] / (float)lv.stride_h), lv.H);
  const int vw = min((int)ceilf(pad_hw[b * 2 + 1] / (float)lv.stride_w), lv.W);
  return y < vh && x < vw;
}

// grid (blocks_per_img, B); gt_max: uint32 (B, Gmax) zeroed by the caller (IoU >= 0: the bit
// pattern orders like the value)
__global__ void __launch_bounds__(RL_THREADS)
rpn_loss_gtmax_kernel(const __grid_constant__ RpnLossArgs a, const float* __restrict__ base,
                      const float* __restrict__ gt_boxes, const int32_t* __restrict__ num_gt,
                      const float* __restrict__ pad_hw, unsigned int* __restrict__ gt_max) {
  extern __shared__ __align__(16) unsigned char rl_smem[];
  float4* s_gt = reinterpret_cast<float4*>(rl_smem);                          // [Gmax]
  unsigned int* s_max = reinterpret_cast<unsigned int*>(s_gt + a.Gmax);       // [Gmax]
  const int b = blockIdx.y;
  const int G = min(num_gt[b], a.Gmax);
  if (G == 0) return;
  for (int i = threadIdx.x; i < G; i += RL_THREADS) {
    s_gt[i] = reinterpret_cast<const float4*>(gt_boxes)[(size_t)b * a.Gmax + i];
    s_max[i] = 0u;
  }
  __syncthreads();
  const int l = rl_level_of_block(a, blockIdx.x);
  const RpnLossLevel& lv = a.lv[l];
  const int t = (blockIdx.x - lv.block_base) * RL_THREADS + threadIdx.x;
  const int hw = lv.H * lv.W;
  bool live = false;
  float4 anc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (t < lv.n) {
    const int an = t / hw, pos = t - an * hw;
    const int y = pos / lv.W, x = pos - y * lv.W;
    live = rl_valid(lv, pad_hw, b, y, x);
    anc = rl_anchor(lv, base, a.A, l, an, y, x);
  }
  for (int i = 0; i < G; ++i) {
    const float v = live ? iou_bbox_overlaps(s_gt[i], anc) : 0.f;
    const unsigned int m = __reduce_max_sync(0xffffffffu, __float_as_uint(v));
    if ((threadIdx.x & 31) == 0 && m > 0u) atomicMax(&s_max[i], m);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < G; i += RL_THREADS)
    if (s_max[i] > 0u) atomicMax(&gt_max[(size_t)b * a.Gmax + i], s_max[i]);
}

// max(a, b) derivative w.r.t. a as autograd defines it: 1 if a > b, 1/2 on ties, else 0
__device__ __forceinline__ float rl_dmax(float a, float b) {
  return a > b ? 1.f : (a == b ? 0.5f : 0.f);
}

// grid (blocks_per_img, B); partials: (B, blocks_per_img, RL_SUMS)
__global__ void __launch_bounds__(RL_THREADS)
rpn_loss_main_kernel(const __grid_constant__ RpnLossArgs a, const float* __restrict__ base,
                     const float* __restrict__ gt_boxes, const int32_t* __restrict__ num_gt,
                     const float* __restrict__ pad_hw, const unsigned int* __restrict__ gt_max,
                     float* __restrict__ partials) {
  extern __shared__ __align__(16) unsigned char rl_smem[];
  float4* s_gt = reinterpret_cast<float4*>(rl_smem);                   // [Gmax]
  float* s_gmax = reinterpret_cast<float*>(s_gt + a.Gmax);             // [Gmax]
  __shared__ float s_red[RL_THREADS / 32][RL_SUMS];
  const int b = blockIdx.y;
  const int G = min(num_gt[b], a.Gmax);
  for (int i = threadIdx.x; i < G; i += RL_THREADS) {
    s_gt[i] = reinterpret_cast<const float4*>(gt_boxes)[(size_t)b * a.Gmax + i];
    s_gmax[i] = __uint_as_float(gt_max[(size_t)b * a.Gmax + i]);
  }
  __syncthreads();
  const int l = rl_level_of_block(a, blockIdx.x);
  const RpnLossLevel& lv = a.lv[l];
  const int t = (blockIdx.x - lv.block_base) * RL_THREADS + threadIdx.x;
  const int hw = lv.H * lv.W;
  float s_cls = 0.f, s_bbox = 0.f, s_bce = 0.f, s_iou = 0.f, s_pos = 0.f;
  if (t < lv.n) {
    const int an = t / hw, pos = t - an * hw;
    const int y = pos / lv.W, x = pos - y * lv.W;
    const bool valid = rl_valid(lv, pad_hw, b, y, x);
    const float4 anc = rl_anchor(lv, base, a.A, l, an, y, x);
    // ---- MaxIoUAssigner.assign_wrt_overlaps ----
    int gt_ind = -1;          // -1 ignore / invalid, 0 negative, k + 1 positive
    if (valid) {
      if (G == 0) {
        gt_ind = 0;
      } else {
        float best = iou_bbox_overlaps(s_gt[0], anc);
        int arg = 0;
        int lowq = (s_gmax[0] >= a.min_pos_iou && best == s_gmax[0]) ? 0 : -1;
        for (int i = 1; i < G; ++i) {
          const float v = iou_bbox_overlaps(s_gt[i], anc);
          if (v > best) { best = v; arg = i; }            // first maximum
          if (s_gmax[i] >= a.min_pos_iou && v == s_gmax[i]) lowq = i;   // later GTs override
        }
        if (best >= 0.f && best < a.neg_iou_thr) gt_ind = 0;
        if (best >= a.pos_iou_thr) gt_ind = arg + 1;
        if (lowq >= 0) gt_ind = lowq + 1;
      }
    }
    const bool is_pos = gt_ind > 0;
    const float lw = gt_ind >= 0 ? 1.f : 0.f;             // pos_weight <= 0 -> 1 (anchor_head.py:247-252)
    const size_t plane = ((size_t)b * a.A + an) * hw + pos;
    // ---- positives: regression + IoU branch ----
    float gd0 = 0.f, gd1 = 0.f, gd2 = 0.f, gd3 = 0.f, gu = 0.f;
    float iou_q = 0.f;                                      // varifocal target (0 off positives)
    const size_t bplane = ((size_t)b * a.A * 4 + an * 4) * hw + pos;
    if (is_pos) {
      const float4 gt = s_gt[gt_ind - 1];
      const float d0 = lv.bbox[bplane], d1 = lv.bbox[bplane + hw];
      const float d2 = lv.bbox[bplane + 2 * (size_t)hw], d3 = lv.bbox[bplane + 3 * (size_t)hw];
      const float px = (anc.x + anc.z) * 0.5f, py = (anc.y + anc.w) * 0.5f;
      const float pw = anc.z - anc.x, ph = anc.w - anc.y;
      const float M = a.max_ratio;
      const float dwc = fminf(fmaxf(d2, -M), M), dhc = fminf(fmaxf(d3, -M), M);
      const float gw = pw * expf(dwc), gh = ph * expf(dhc);
      const float gx = px + pw * d0, gy = py + ph * d1;
      const float x1 = gx - gw * 0.5f, y1 = gy - gh * 0.5f, x2 = gx + gw * 0.5f, y2 = gy + gh * 0.5f;
      // aligned IoU with eps 1e-6 (iou2d_calculator.py:214-253)
      const float ltx = fmaxf(x1, gt.x), lty = fmaxf(y1, gt.y);
      const float rbx = fminf(x2, gt.z), rby = fminf(y2, gt.w);
      const float iwr = rbx - ltx, ihr = rby - lty;
      const float iw = fmaxf(iwr, 0.f), ih = fmaxf(ihr, 0.f);
      const float inter = iw * ih;
      const float ap = (x2 - x1) * (y2 - y1), ag = (gt.z - gt.x) * (gt.w - gt.y);
      const float uni = ap + ag - inter;
      const float uc = fmaxf(uni, 1e-6f);
      const float iou_t = inter / uc;                      // iou_target (detached) == loss IoU
      const float wgt = fmaxf(a.gamma == 0.5f ? sqrtf(iou_t) : powf(iou_t, a.gamma), 1e-12f);
      // encoded target (bbox2delta) and the MSE "aug" loss
      const float tx = (gt.x + gt.z) * 0.5f, ty = (gt.y + gt.w) * 0.5f;
      const float tw = gt.z - gt.x, th = gt.w - gt.y;
      const float e0 = (tx - px) / pw, e1 = (ty - py) / ph;
      const float e2 = logf(tw / pw), e3 = logf(th / ph);
      const float f0 = d0 - e0, f1 = d1 - e1, f2 = d2 - e2, f3 = d3 - e3;
      const float l_aug = a.w_aug * wgt * (f0 * f0 + f1 * f1 + f2 * f2 + f3 * f3);
      const float iou_c = fmaxf(iou_t, 1e-6f);
      const float l_iou = a.w_bbox * wgt * -logf(iou_c);
      s_bbox = 0.5f * (l_iou + l_aug);
      s_iou = iou_t;
      s_pos = 1.f;
      iou_q = iou_t;
      // d(-log iou) / d(box), then through the decode
      const float mw = iwr >= 0.f ? 1.f : 0.f, mh = ihr >= 0.f ? 1.f : 0.f;
      const float di0 = -ih * mw * rl_dmax(x1, gt.x), di1 = -iw * mh * rl_dmax(y1, gt.y);
      const float di2 = ih * mw * rl_dmax(gt.z, x2), di3 = iw * mh * rl_dmax(gt.w, y2);
      const float hgt = y2 - y1, wid = x2 - x1;
      const float um = rl_dmax(uni, 1e-6f);
      const float coef = (iou_t >= 1e-6f ? -1.f : 0.f) / iou_c * wgt * a.w_bbox * 0.5f;
      const float uc2 = uc * uc;
      const float gb0 = coef * (di0 / uc - inter * um * (-hgt - di0) / uc2);
      const float gb1 = coef * (di1 / uc - inter * um * (-wid - di1) / uc2);
      const float gb2 = coef * (di2 / uc - inter * um * (hgt - di2) / uc2);
      const float gb3 = coef * (di3 / uc - inter * um * (wid - di3) / uc2);
      const float cw = fabsf(d2) <= M ? 1.f : 0.f, ch = fabsf(d3) <= M ? 1.f : 0.f;
      const float s2 = a.w_aug * wgt;                      // 0.5 * 2 * w
      gd0 = (gb0 + gb2) * pw + s2 * f0;
      gd1 = (gb1 + gb3) * ph + s2 * f1;
      gd2 = (gb2 - gb0) * 0.5f * gw * cw + s2 * f2;
      gd3 = (gb3 - gb1) * 0.5f * gh * ch + s2 * f3;
      // IoU branch: BCE with logits against iou_target
      const float xu = lv.iou[plane];
      s_bce = fmaxf(xu, 0.f) - xu * iou_t + log1pf(expf(-fabsf(xu)));
      gu = a.w_iou * (1.f / (1.f + expf(-xu)) - iou_t);
    }
    // ---- classification loss on the objectness logit ----
    {
      const float xv = lv.cls[plane];
      const float p = 1.f / (1.f + expf(-xv));
      float g = 0.f;
      if (a.cls_loss_type == 0) {
        // sigmoid focal loss (focal_loss.py:13-58 == mmcv sigmoid_focal_loss), weighted by the
        // label weight: invalid / ignored anchors contribute nothing
        if (lw > 0.f) {
          const float tt = is_pos ? 1.f : 0.f;
          const float bce = fmaxf(xv, 0.f) - xv * tt + log1pf(expf(-fabsf(xv)));
          const float pt = is_pos ? 1.f - p : p;
          const float aw = is_pos ? a.focal_alpha : 1.f - a.focal_alpha;
          const float ptg1 = a.focal_gamma == 2.f ? pt : powf(pt, a.focal_gamma - 1.f);
          const float fw = aw * ptg1 * pt;
          s_cls = bce * fw;
          const float dfw = aw * a.focal_gamma * ptg1 * (is_pos ? -1.f : 1.f) * p * (1.f - p);
          g = a.w_cls * (fw * (p - tt) + bce * dfw);
        }
      } else {
        // VarifocalLoss(iou_weighted=True) against q = iou_target on positives, 0 elsewhere;
        // the reference passes no weights here, so EVERY anchor counts
        // (atss_rpn_head.py:393-397, varifocal_loss.py:45-57)
        const float q = iou_q;
        const float bce = fmaxf(xv, 0.f) - xv * q + log1pf(expf(-fabsf(xv)));
        if (q > 0.f) {
          s_cls = bce * q;
          g = a.w_cls * q * (p - q);
        } else {
          const float dq = p - q, ad = fabsf(dq);
          const float adg1 = a.focal_gamma == 2.f ? ad : powf(ad, a.focal_gamma - 1.f);
          const float fw = a.focal_alpha * adg1 * ad;
          s_cls = bce * fw;
          const float dfw = a.focal_alpha * a.focal_gamma * adg1 * (dq > 0.f ? 1.f : (dq < 0.f ? -1.f : 0.f))
                            * p * (1.f - p);
          g = a.w_cls * (fw * (p - q) + bce * dfw);
        }
      }
      lv.g_cls[plane] = g;
    }
    lv.g_bbox[bplane] = gd0;
    lv.g_bbox[bplane + hw] = gd1;
    lv.g_bbox[bplane + 2 * (size_t)hw] = gd2;
    lv.g_bbox[bplane + 3 * (size_t)hw] = gd3;
    lv.g_iou[plane] = gu;
  }
  // ---- block sums (fixed order: warp shuffle tree, then warp 0 over the 8 warps) ----
  float v[RL_SUMS] = {s_cls, s_bbox, s_bce, s_iou, s_pos};
#pragma unroll
  for (int k = 0; k < RL_SUMS; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < RL_SUMS; ++k) s_red[threadIdx.x >> 5][k] = v[k];
  }
  __syncthreads();
  if (threadIdx.x < RL_SUMS) {
    float s = 0.f;
    for (int w = 0; w < RL_THREADS / 32; ++w) s += s_red[w][threadIdx.x];
    partials[((size_t)b * a.blocks_per_img + blockIdx.x) * RL_SUMS + threadIdx.x] = s;
  }
}

// one block.  sums: [0,L) cls, [L,2L) bbox (raw), [2L,3L) bce, 3L num_pos, 3L+1 sum iou_target
__global__ void __launch_bounds__(RL_THREADS)
rpn_loss_reduce_kernel(const __grid_constant__ RpnLossArgs a, const float* __restrict__ partials,
                       float* __restrict__ sums) {
  __shared__ double s_acc[RL_THREADS];
  const int nq = 3 * a.L + 2;
  for (int q = 0; q < nq; ++q) {
    // quantity q: (kind k, level l) or a normaliser over every level
    const int k = q < 3 * a.L ? q / a.L : (q == 3 * a.L ? 4 : 3);
    const int l0 = q < 3 * a.L ? q % a.L : 0, l1 = q < 3 * a.L ? l0 + 1 : a.L;
    const int kk = k == 0 ? 0 : (k == 1 ? 1 : (k == 2 ? 2 : k));
    double acc = 0.0;
    for (int l = l0; l < l1; ++l) {
      const int bb = a.lv[l].block_base;
      const int be = l + 1 < a.L ? a.lv[l + 1].block_base : a.blocks_per_img;
      const int nblk = be - bb;
      for (int i = threadIdx.x; i < nblk * a.B; i += RL_THREADS) {
        const int b = i / nblk, j = i - b * nblk;
        acc += (double)partials[((size_t)b * a.blocks_per_img + bb + j) * RL_SUMS + kk];
      }
    }
    s_acc[threadIdx.x] = acc;
    __syncthreads();
    for (int o = RL_THREADS / 2; o > 0; o >>= 1) {
      if (threadIdx.x < o) s_acc[threadIdx.x] += s_acc[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      const float w = k == 0 ? a.w_cls : (k == 2 ? a.w_iou : 1.f);
      sums[q] = (float)(s_acc[0] * (double)w);
    }
    __syncthreads();
  }
}

struct RpnLossScaleArgs {
  const float* raw[3][BRCNN_MAX_LEVELS];
  float* out[3][BRCNN_MAX_LEVELS];
  int n[3][BRCNN_MAX_LEVELS];          // elements
  int block_base[3 * BRCNN_MAX_LEVELS + 1];
  int L;
};
// out = raw * scale[kind * L + level] (scale on the device: upstream grad / normaliser)
__global__ void __launch_bounds__(RL_THREADS)
rpn_loss_scale_kernel(const __grid_constant__ RpnLossScaleArgs s, const float* __restrict__ scale) {
  int seg = 0;
  while (seg + 1 < 3 * s.L && (int)blockIdx.x >= s.block_base[seg + 1]) ++seg;
  const int kind = seg / s.L, l = seg - kind * s.L;
  const float f = scale[seg];
  const int n = s.n[kind][l];
  const float* __restrict__ in = s.raw[kind][l];
  float* __restrict__ out = s.out[kind][l];
  const int i0 = ((int)blockIdx.x - s.block_base[seg]) * RL_THREADS * 4 + threadIdx.x * 4;
  if (i0 + 3 < n && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    float4 v = *reinterpret_cast<const float4*>(in + i0);
    v.x *= f; v.y *= f; v.z *= f; v.w *= f;
    *reinterpret_cast<float4*>(out + i0) = v;
  } else {
    for (int i = i0; i < min(i0 + 4, n); ++i) out[i] = in[i] * f;
  }
}

}  // namespace brcnn
