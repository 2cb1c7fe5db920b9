// R-CNN training front-end (the caller side of the RoI
// kernels): MaxIoUAssigner.assign + add_gt_ + RandomSampler ordering +
// BBoxHead.get_targets + ProbRoIHead prior extraction, batched over images.
//
// Reference: mmdet/core/bbox/assigners/max_iou_assigner.py:61-212,
//   assigners/assign_result.py:191-205 (add_gt_), iou_calculators/
//   iou2d_calculator.py (bbox_overlaps, mode='iou', eps=1e-6),
//   samplers/base_sampler.py:35-102, random_sampler.py:32-82,
//   sampling_result.py:26-55, roi_heads/bbox_heads/bbox_head.py:122-253,
//   coder/delta_xywh_bbox_coder.py:98-141 (bbox2delta),
//   roi_heads/prob_roi_head.py:51-64 (prior vector).
//
//   rcnn_assign_kernel         thread per candidate (GTs first, then proposals,
//       the sampler's concatenation order): max / first-argmax IoU over the
//       image's GTs (held in shared memory), thresholds -> gt_inds, and the
//       number of positive / negative candidates per image.
//   (host: the reference draws torch.randperm on the CPU RNG; the two counts
//    are the only values that cross to the host, once per step)
//   rcnn_sample_target_kernel  one CTA per image: ordered compaction of the
//       positive / negative candidate lists, selection through the host's
//       permutations, index sort (RandomSampler's .unique()), and emission of
//       the final rows: rois, labels, label weights, bbox targets (bbox2delta)
//       and weights, prior -- positives then negatives, as SamplingResult.bboxes.
//
// match_low_quality=True (unused by the R-CNN stage of the named configs) is not
// handled here; the Python path in sampling.py remains the fallback.
#pragma once
#include "common.cuh"

namespace brcnn {

struct AssignArgs {
  int B, M, Gmax;              // images, proposal capacity, GT capacity
  float pos_iou_thr, neg_iou_thr;
};

// IoU exactly as bbox_overlaps(mode='iou', eps=1e-6) evaluates it in fp32
__device__ __forceinline__ float iou_bbox_overlaps(float4 g, float4 p) {
  const float area1 = (g.z - g.x) * (g.w - g.y);
  const float area2 = (p.z - p.x) * (p.w - p.y);
  const float ltx = fmaxf(g.x, p.x), lty = fmaxf(g.y, p.y);
  const float rbx = fminf(g.z, p.z), rby = fminf(g.w, p.w);
  float w = rbx - ltx, h = rby - lty;
  w = w < 0.f ? 0.f : w;       // clamp(min=0); NaN stays NaN like torch.clamp
  h = h < 0.f ? 0.f : h;
  const float overlap = w * h;
  float uni = area1 + area2 - overlap;
  uni = uni < 1e-6f ? 1e-6f : uni;   // torch.max(union, eps)
  return overlap / uni;
}

// grid (ceil((Gmax + M) / 256), B); gt_inds: (B, Gmax + M) int32 in the sampler's
// index space (GT g -> slot g, proposal j -> slot num_gt[b] + j); counts: (B, 2)
__global__ void __launch_bounds__(256)
rcnn_assign_kernel(const AssignArgs a, const float* __restrict__ proposals /* (B,M,5) */,
                   const int32_t* __restrict__ num_props, const float* __restrict__ gt_boxes
                   /* (B,Gmax,4) */, const int32_t* __restrict__ num_gt,
                   int32_t* __restrict__ gt_inds, int32_t* __restrict__ counts) {
  extern __shared__ __align__(16) float4 s_gt[];
  const int b = blockIdx.y;
  const int G = min(num_gt[b], a.Gmax), n = min(num_props[b], a.M);
  for (int i = threadIdx.x; i < G; i += blockDim.x)
    s_gt[i] = reinterpret_cast<const float4*>(gt_boxes)[(size_t)b * a.Gmax + i];
  __syncthreads();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  int assigned = -2;   // -2: slot not in use
  if (c < G) {
    assigned = c + 1;                      // AssignResult.add_gt_
  } else if (c < G + n) {
    const float* pr = proposals + ((size_t)b * a.M + (c - G)) * 5;
    const float4 p = make_float4(pr[0], pr[1], pr[2], pr[3]);
    if (G == 0) {
      assigned = 0;                        // no GT: everything is background (:113-124)
    } else {
      float best = iou_bbox_overlaps(s_gt[0], p);
      int arg = 0;
      for (int i = 1; i < G; ++i) {
        const float v = iou_bbox_overlaps(s_gt[i], p);
        if (v > best) { best = v; arg = i; }   // first maximum, like torch.max on the CPU
      }
      assigned = -1;
      if (best >= 0.f && best < a.neg_iou_thr) assigned = 0;
      if (best >= a.pos_iou_thr) assigned = arg + 1;
    }
  }
  if (c < a.Gmax + a.M) gt_inds[(size_t)b * (a.Gmax + a.M) + c] = assigned;
  const unsigned mp = __ballot_sync(0xffffffffu, assigned > 0);
  const unsigned mn = __ballot_sync(0xffffffffu, assigned == 0);
  if ((threadIdx.x & 31) == 0) {
    if (mp) atomicAdd(counts + b * 2 + 0, __popc(mp));
    if (mn) atomicAdd(counts + b * 2 + 1, __popc(mn));
  }
}

struct SampleArgs {
  int B, M, Gmax, num_classes;
  int perm_cap;                 // columns of perm_pos / perm_neg
  float means[4], stds[4];
  float pos_weight;             // label weight of positives (1 when cfg.pos_weight <= 0)
};

constexpr int ST_THREADS = 1024;

// grid B, block ST_THREADS.  Dynamic smem: (Gmax + M) ints (list) + sel_cap u64 (sort).
// plan (host, per image): [n_pos_sel, n_neg_sel, row_base, use_perm_pos, use_perm_neg]
__global__ void __launch_bounds__(ST_THREADS)
rcnn_sample_target_kernel(const SampleArgs a, const float* __restrict__ proposals,
                          const int32_t* __restrict__ num_props,
                          const float* __restrict__ gt_boxes,
                          const int64_t* __restrict__ gt_labels,
                          const int32_t* __restrict__ num_gt,
                          const int32_t* __restrict__ gt_inds,
                          const int32_t* __restrict__ plan,
                          const int32_t* __restrict__ perm_pos,
                          const int32_t* __restrict__ perm_neg, int sel_cap,
                          float* __restrict__ rois, int64_t* __restrict__ labels,
                          float* __restrict__ label_weights, float* __restrict__ bbox_targets,
                          float* __restrict__ bbox_weights, float* __restrict__ prior) {
  extern __shared__ __align__(16) unsigned char st_smem[];
  unsigned long long* s_sel = reinterpret_cast<unsigned long long*>(st_smem);   // [sel_cap]
  int* s_list = reinterpret_cast<int*>(st_smem + (size_t)sel_cap * 8);          // [Gmax + M]
  __shared__ int s_wsum[ST_THREADS / 32];
  __shared__ int s_total;
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int G = min(num_gt[b], a.Gmax), n = min(num_props[b], a.M);
  const int C = G + n;
  const int32_t* gi = gt_inds + (size_t)b * (a.Gmax + a.M);
  const int n_pos_sel = plan[b * 5 + 0], n_neg_sel = plan[b * 5 + 1], row_base = plan[b * 5 + 2];
  const int use_pp = plan[b * 5 + 3], use_pn = plan[b * 5 + 4];
  const float4* gtb = reinterpret_cast<const float4*>(gt_boxes) + (size_t)b * a.Gmax;
  const int64_t* gtl = gt_labels + (size_t)b * a.Gmax;

  for (int pass = 0; pass < 2; ++pass) {       // 0: positives, 1: negatives
    const int n_sel = pass == 0 ? n_pos_sel : n_neg_sel;
    const int use_perm = pass == 0 ? use_pp : use_pn;
    const int32_t* perm = (pass == 0 ? perm_pos : perm_neg) + (size_t)b * a.perm_cap;
    // ---- ordered compaction of the candidate list (index-ascending) ----
    if (tid == 0) s_total = 0;
    __syncthreads();
    for (int c0 = 0; c0 < C; c0 += ST_THREADS) {
      const int c = c0 + tid;
      bool f = false;
      if (c < C) f = pass == 0 ? (gi[c] > 0) : (gi[c] == 0);
      const unsigned bm = __ballot_sync(0xffffffffu, f);
      if (lane == 0) s_wsum[wid] = __popc(bm);
      __syncthreads();
      int wbase = 0, tot = 0;
      for (int w = 0; w < ST_THREADS / 32; ++w) {
        const int v = s_wsum[w];
        if (w < wid) wbase += v;
        tot += v;
      }
      const int base = s_total;
      if (f) s_list[base + wbase + __popc(bm & ((1u << lane) - 1u))] = c;
      __syncthreads();
      if (tid == 0) s_total = base + tot;
      __syncthreads();
    }
    // ---- selection: all of them, or gallery[perm[:n_sel]] then .unique() (= sort) ----
    int np2 = 1;
    while (np2 < n_sel) np2 <<= 1;
    for (int i = tid; i < np2; i += ST_THREADS) {
      unsigned long long key = 0ull;            // pads sort last (descending order)
      if (i < n_sel) {
        const int c = use_perm ? s_list[perm[i]] : s_list[i];
        key = (unsigned long long)(0xFFFFFFFFu - (unsigned)c) + 1ull;   // ascending c
      }
      s_sel[i] = key;
    }
    if (use_perm && n_sel > 1) bitonic_sort_desc_u64(s_sel, np2);
    else __syncthreads();
    // ---- emit rows: positives at [row_base, +n_pos_sel), negatives after them ----
    const int out0 = row_base + (pass == 0 ? 0 : n_pos_sel);
    for (int i = tid; i < n_sel; i += ST_THREADS) {
      const int c = (int)(0xFFFFFFFFu - (unsigned)(s_sel[i] - 1ull));
      const int row = out0 + i;
      float4 box;
      float score = 0.f;
      if (c < G) {
        box = gtb[c];
      } else {
        const float* pr = proposals + ((size_t)b * a.M + (c - G)) * 5;
        box = make_float4(pr[0], pr[1], pr[2], pr[3]);
        score = pr[4];
      }
      float* ro = rois + (size_t)row * 5;
      ro[0] = (float)b; ro[1] = box.x; ro[2] = box.y; ro[3] = box.z; ro[4] = box.w;
      float4 tgt = make_float4(0.f, 0.f, 0.f, 0.f);
      float bw = 0.f, lw = 1.0f, pri;
      int64_t lab = a.num_classes;
      if (pass == 0) {
        const int g = gi[c] - 1;
        const float4 gb = gtb[g];
        lab = gtl[g];
        lw = a.pos_weight;
        bw = 1.0f;
        // bbox2delta, op for op
        const float px = (box.x + box.z) * 0.5f, py = (box.y + box.w) * 0.5f;
        const float pw = box.z - box.x, ph = box.w - box.y;
        const float gx = (gb.x + gb.z) * 0.5f, gy = (gb.y + gb.w) * 0.5f;
        const float gw = gb.z - gb.x, gh = gb.w - gb.y;
        tgt.x = ((gx - px) / pw - a.means[0]) / a.stds[0];
        tgt.y = ((gy - py) / ph - a.means[1]) / a.stds[1];
        tgt.z = (logf(gw / pw) - a.means[2]) / a.stds[2];
        tgt.w = (logf(gh / ph) - a.means[3]) / a.stds[3];
        // prob_roi_head.py:52-57: the first num_gts rows get prior 0, the others the
        // proposal score at pos_inds - num_gts
        pri = (i < G || c < G) ? 0.f : score;
      } else {
        pri = 1.0f - score;                   // prob_roi_head.py:58-59 (negatives are never GTs)
      }
      labels[row] = lab;
      label_weights[row] = lw;
      reinterpret_cast<float4*>(bbox_targets)[row] = tgt;
      reinterpret_cast<float4*>(bbox_weights)[row] = make_float4(bw, bw, bw, bw);
      prior[row] = pri;
    }
    __syncthreads();
  }
}

}  // namespace brcnn
