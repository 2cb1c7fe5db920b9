// K5: boosting-reweighted R-CNN loss, forward value and gradients in one
// launch.  Reference: prob_roi_head.py:107-154 (norm_loss :151-154),
// convfc_bbox_head.py:332-418, cross_entropy_loss.py:10-50, losses/utils.py:
// 28-55, smooth_l1_loss.py:35-52, accuracy.py:6-51  (SURVEY.md App. A8).
//
//   l_i   = loss_cls_weight * (CE(cls_score_i, label_i) * label_weight_i)
//   w_i   = (1 - prior_i)^gamma (* alpha if alpha != 0)
//   s     = sum(l) / sum(w*l)
//   loss_cls = sum(l_i * (w_i*s)) / N          (w*s is detached)
//   d loss_cls / d x_ic = (w_i*s/N) * loss_cls_weight * label_weight_i
//                          * (softmax_ic - [c == label_i])
//   loss_bbox = sum(loss_bbox_weight * |pred - tgt| * bbox_w)[pos] / N
//               (or .mean() when reg_norm == 'mean')
//   acc = 100 * mean(argmax_c x_ic == label_i)
//
// Two launches over up to one CTA per SM (the problem needs two global sums
// before any gradient can be written): part 1 = softmax / CE / L1 per row and
// one partial-sum record per CTA, part 2 = fixed-order reduction of the records
// (identical in every CTA) + gradients.  Row -> (CTA, warp) mapping and both
// reduction trees are fixed, so the result is bit-reproducible run to run.
#pragma once
#include "common.cuh"

namespace brcnn {

struct LossArgs {
  int N, C, agnostic, reg_norm_mean;
  float gamma, alpha, wcls, wbbox;
};

__device__ __forceinline__ float boost_weight(float prior, float gamma, float alpha) {
  const float base = 1.0f - prior;
  float w;
  if (gamma == 0.5f) w = sqrtf(base);
  else if (gamma == 1.0f) w = base;
  else if (gamma == 0.0f) w = 1.0f;
  else w = powf(base, gamma);
  if (alpha != 0.f) w = w * alpha;
  return w;
}

constexpr int BL_THREADS = 256;
constexpr int BL_NW = BL_THREADS / 32;
constexpr int BL_REC = 8;   // floats per partial record: sl, swl, slb, corr, npos

// grid G; CTA g owns rows [g*rpc, min(N, (g+1)*rpc))
__global__ void __launch_bounds__(BL_THREADS)
boost_loss_part1_kernel(const LossArgs a, int rpc, const float* __restrict__ cls_score,
                        const int64_t* __restrict__ labels,
                        const float* __restrict__ label_weights,
                        const float* __restrict__ prior,
                        const float* __restrict__ bbox_pred,
                        const float* __restrict__ bbox_targets,
                        const float* __restrict__ bbox_weights,
                        float* __restrict__ partials, float* __restrict__ grad_cls) {
  __shared__ float s_l[BL_NW], s_wl[BL_NW], s_lb[BL_NW];
  __shared__ int s_correct[BL_NW], s_npos[BL_NW];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int C1 = a.C + 1;
  const int r0 = blockIdx.x * rpc, r1 = min(a.N, r0 + rpc);

  float acc_l = 0.f, acc_wl = 0.f;
  int acc_correct = 0;
  for (int i = r0 + wid; i < r1; i += BL_NW) {
    const float* x = cls_score + (size_t)i * C1;
    float m = -INFINITY;
    int am = 0x7fffffff;
    for (int c = lane; c < C1; c += 32) {
      const float v = x[c];
      if (v > m) { m = v; am = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, m, o);
      const int oa = __shfl_xor_sync(0xffffffffu, am, o);
      if (om > m || (om == m && oa < am)) { m = om; am = oa; }
    }
    float sum = 0.f;
    for (int c = lane; c < C1; c += 32) sum += expf(x[c] - m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
    float* g = grad_cls + (size_t)i * C1;
    for (int c = lane; c < C1; c += 32) g[c] = expf(x[c] - m) * inv;   // softmax, scaled in part 2
    if (lane == 0) {
      const int64_t lab = labels[i];
      const float lw = label_weights ? label_weights[i] : 1.0f;
      float l = 0.f;
      if (lab >= 0 && lab < C1) {
        const float ce = (m + logf(sum)) - x[lab];
        l = a.wcls * (ce * lw);
      }
      const float w = boost_weight(prior[i], a.gamma, a.alpha);
      acc_l += l;
      acc_wl += w * l;
      acc_correct += ((int64_t)am == lab);
    }
  }
  if (lane == 0) { s_l[wid] = acc_l; s_wl[wid] = acc_wl; s_correct[wid] = acc_correct; }

  float acc_lb = 0.f;
  int acc_npos = 0;
  for (int i = r0 + tid; i < r1; i += BL_THREADS) {
    const int64_t lab = labels[i];
    if (lab >= 0 && lab < a.C) {
      ++acc_npos;
      const float* pr = bbox_pred + (a.agnostic ? (size_t)i * 4 : ((size_t)i * a.C + lab) * 4);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float d = fabsf(pr[j] - bbox_targets[(size_t)i * 4 + j]);
        acc_lb += a.wbbox * (d * bbox_weights[(size_t)i * 4 + j]);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc_lb += __shfl_xor_sync(0xffffffffu, acc_lb, o);
    acc_npos += __shfl_xor_sync(0xffffffffu, acc_npos, o);
  }
  if (lane == 0) { s_lb[wid] = acc_lb; s_npos[wid] = acc_npos; }
  __syncthreads();
  if (tid == 0) {
    float sl = 0.f, swl = 0.f, slb = 0.f;
    int corr = 0, npos = 0;
    for (int w = 0; w < BL_NW; ++w) {
      sl += s_l[w]; swl += s_wl[w]; slb += s_lb[w];
      corr += s_correct[w]; npos += s_npos[w];
    }
    float* rec = partials + (size_t)blockIdx.x * BL_REC;
    rec[0] = sl; rec[1] = swl; rec[2] = slb;
    rec[3] = __int_as_float(corr); rec[4] = __int_as_float(npos);
  }
}

// grid G (same as part 1)
__global__ void __launch_bounds__(BL_THREADS)
boost_loss_part2_kernel(const LossArgs a, int rpc, int G, const float* __restrict__ partials,
                        const int64_t* __restrict__ labels,
                        const float* __restrict__ label_weights,
                        const float* __restrict__ prior,
                        const float* __restrict__ bbox_pred,
                        const float* __restrict__ bbox_targets,
                        const float* __restrict__ bbox_weights,
                        float* __restrict__ out, float* __restrict__ grad_cls,
                        float* __restrict__ grad_bbox) {
  __shared__ float s_scale, s_bden;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int C1 = a.C + 1, N = a.N;
  if (wid == 0) {
    // fixed-order reduction of the G records: lane-strided serial sums, then a
    // fixed shuffle tree (same in every CTA -> same bits everywhere)
    float sl = 0.f, swl = 0.f, slb = 0.f;
    int corr = 0, npos = 0;
    for (int g = lane; g < G; g += 32) {
      const float* rec = partials + (size_t)g * BL_REC;
      sl += rec[0]; swl += rec[1]; slb += rec[2];
      corr += __float_as_int(rec[3]); npos += __float_as_int(rec[4]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sl += __shfl_xor_sync(0xffffffffu, sl, o);
      swl += __shfl_xor_sync(0xffffffffu, swl, o);
      slb += __shfl_xor_sync(0xffffffffu, slb, o);
      corr += __shfl_xor_sync(0xffffffffu, corr, o);
      npos += __shfl_xor_sync(0xffffffffu, npos, o);
    }
    if (lane == 0) {
      const float s = sl / swl;
      const float bden = a.reg_norm_mean ? (npos > 0 ? (float)(npos * 4) : 1.0f) : (float)N;
      if (blockIdx.x == 0) {
        // loss_cls = sum(l * (w*s)) / N  ==  (s * swl) / N
        out[0] = N > 0 ? (s * swl) / (float)N : 0.f;
        out[1] = npos > 0 ? slb / bden : 0.f;
        out[2] = N > 0 ? 100.0f * (float)corr / (float)N : 0.f;
        out[3] = sl; out[4] = swl; out[5] = (float)npos; out[6] = s; out[7] = 0.f;
      }
      s_scale = s; s_bden = bden;
    }
  }
  __syncthreads();
  const float s = s_scale, bden = s_bden;
  const int r0 = blockIdx.x * rpc, r1 = min(N, r0 + rpc);
  for (int i = r0 + wid; i < r1; i += BL_NW) {
    const int64_t lab = labels[i];
    const float lw = label_weights ? label_weights[i] : 1.0f;
    const float w = boost_weight(prior[i], a.gamma, a.alpha);
    float coef = (w * s) / (float)N * a.wcls * lw;
    if (!(lab >= 0 && lab < C1)) coef = 0.f;
    float* g = grad_cls + (size_t)i * C1;
    for (int c = lane; c < C1; c += 32) {
      const float p = g[c];
      g[c] = coef * (p - ((int64_t)c == lab ? 1.0f : 0.0f));
    }
  }
  if (grad_bbox != nullptr) {
    for (int i = r0 + tid; i < r1; i += BL_THREADS) {
      const int64_t lab = labels[i];
      if (lab >= 0 && lab < a.C) {
        const size_t o = a.agnostic ? (size_t)i * 4 : ((size_t)i * a.C + lab) * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float d = bbox_pred[o + j] - bbox_targets[(size_t)i * 4 + j];
          const float sg = d > 0.f ? 1.0f : (d < 0.f ? -1.0f : 0.f);
          grad_bbox[o + j] = a.wbbox * bbox_weights[(size_t)i * 4 + j] * sg / bden;
        }
      }
    }
  }
}

// launch geometry shared by the workspace query and the launcher
inline void boost_loss_grid(int N, int* G, int* rpc) {
  int g = (N + BL_NW - 1) / BL_NW;
  if (g > 148) g = 148;
  if (g < 1) g = 1;
  *rpc = (N + g - 1) / g;
  if (*rpc < 1) *rpc = 1;
  *G = N > 0 ? (N + *rpc - 1) / *rpc : 1;
}

}  // namespace brcnn
