/*
 * oracle/brcnn_oracle.c — CPU restatement of the reference algorithm for the
 * proposal-to-RoI hot path.  TEST INFRASTRUCTURE ONLY: nothing under
 * boosting_rcnn_b200/ may import, link or execute this file; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * use it, as the checker or the timed CPU baseline.
 *
 * Reference: mousecpn/Boosting-R-CNN (mmdet 2.17.0 fork) + mmcv-full 1.4.0.
 * mmcv's source is NOT vendored in /root/reference (mmdet/__init__.py:19-20
 * pins mmcv-full>=1.3.8,<=1.4.0); its native ops nms / roi_align are
 * restated here from their published algorithm and pinned
 * against torchvision.ops.{nms,roi_align} (the same algorithm mmcv delegates
 * to under use_torchvision=True) in tests/test_oracle_*.py.  Parity status of
 * mmcv-native pieces vs mmcv itself: UNPINNED (mmcv cannot be installed here);
 * pieces restated from reference Python are pinned by tests/golden/.
 *
 * Pinned arithmetic: fp32 throughout, no FMA contraction
 * (-ffp-contract=off), exp() = oracle_expf below (same operations as
 * csrc/common.cuh pinned_expf), ties in every sort broken by lower index
 * (stable descending sort; the reference's torch.sort is unstable, F5).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fPIC -shared
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define API __attribute__((visibility("default")))

/* ---------------------------------------------------------------- math --- */
static inline float as_float(int32_t i) { float f; memcpy(&f, &i, 4); return f; }
static inline uint32_t as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

API float oracle_expf(float x) {
  if (x != x) return x;
  if (x > 88.72283905206835f) return as_float(0x7f800000);
  if (x < -103.972076416f) return 0.0f;
  const float LOG2E = 1.44269504088896341f;
  const float C1 = 0.693359375f;
  const float C2 = -2.12194440e-4f;
  float t = x * LOG2E;
  int n = (int)nearbyintf(t); /* round-half-even, default rounding mode */
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
  int n1 = n / 2;
  int n2 = n - n1;
  float s1 = as_float((n1 + 127) << 23);
  float s2 = as_float((n2 + 127) << 23);
  return (y * s1) * s2;
}

API float oracle_sigmoid(float x) { return 1.0f / (1.0f + oracle_expf(-x)); }

API void oracle_expf_array(const float* x, int64_t n, float* out) {
  for (int64_t i = 0; i < n; ++i) out[i] = oracle_expf(x[i]);
}

/* ------------------------------------------------------------- sorting --- */
typedef struct { uint64_t key; } u64key;
static int cmp_desc_u64(const void* a, const void* b) {
  uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
  return x > y ? -1 : (x < y ? 1 : 0);
}
static inline uint32_t ordered_bits(float f) {
  uint32_t u = as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
/* order[] = indices sorted by (score desc, index asc) */
static void stable_argsort_desc(const float* scores, int n, int32_t* order) {
  uint64_t* k = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; ++i)
    k[i] = ((uint64_t)ordered_bits(scores[i]) << 32) | (uint64_t)(0xFFFFFFFFu - (uint32_t)i);
  qsort(k, (size_t)n, sizeof(uint64_t), cmp_desc_u64);
  for (int i = 0; i < n; ++i) order[i] = (int32_t)(0xFFFFFFFFu - (uint32_t)(k[i] & 0xFFFFFFFFu));
  free(k);
}

/* ---------------------------------------------------------- delta2bbox --- */
/* delta_xywh_bbox_coder.py:206-270.  rois (n,4); deltas (n, 4*ncls);
 * out (n, 4*ncls).  max_h/max_w < 0 -> no clipping (max_shape=None). */
static inline float clampf(float v, float lo, float hi) {
  if (v != v) return v;
  return v < lo ? lo : (v > hi ? hi : v);
}
API void oracle_delta2bbox(const float* rois, const float* deltas, int n, int ncls,
                           const float* means, const float* stds, float max_ratio,
                           float max_h, float max_w, float* out) {
  for (int i = 0; i < n; ++i) {
    const float x1 = rois[i * 4], y1 = rois[i * 4 + 1], x2 = rois[i * 4 + 2], y2 = rois[i * 4 + 3];
    const float px = (x1 + x2) * 0.5f, py = (y1 + y2) * 0.5f; /* :218-219 */
    const float pw = x2 - x1, ph = y2 - y1;                   /* :221-222 */
    for (int c = 0; c < ncls; ++c) {
      const float* d = deltas + ((size_t)i * ncls + c) * 4;
      float dx = d[0] * stds[0] + means[0]; /* :210 denorm_deltas */
      float dy = d[1] * stds[1] + means[1];
      float dw = d[2] * stds[2] + means[2];
      float dh = d[3] * stds[3] + means[3];
      float dx_width = pw * dx, dy_height = ph * dy; /* :224-225 */
      dw = clampf(dw, -max_ratio, max_ratio);        /* :234-235 */
      dh = clampf(dh, -max_ratio, max_ratio);
      float gw = pw * oracle_expf(dw), gh = ph * oracle_expf(dh); /* :237-238 */
      float gx = px + dx_width, gy = py + dy_height;              /* :240-241 */
      float b[4];
      b[0] = gx - gw * 0.5f; b[1] = gy - gh * 0.5f; /* :243-246 */
      b[2] = gx + gw * 0.5f; b[3] = gy + gh * 0.5f;
      if (max_h >= 0.f) { /* :268-269 where(<0,0) ; where(>max,max) */
        const float mx[4] = {max_w, max_h, max_w, max_h};
        for (int j = 0; j < 4; ++j) {
          if (b[j] < 0.f) b[j] = 0.f;
          if (b[j] > mx[j]) b[j] = mx[j];
        }
      }
      float* o = out + ((size_t)i * ncls + c) * 4;
      o[0] = b[0]; o[1] = b[1]; o[2] = b[2]; o[3] = b[3];
    }
  }
}

/* ------------------------------------------------------------ mmcv nms --- */
/* mmcv 1.4.0 nms_cpu: boxes (n,4); returns #keep, keep[]
 * = order.masked_select(select), i.e. kept indices in score-desc order. */
API int oracle_nms_cpu(const float* boxes, const float* scores, int n, float thr,
                       int offset, int64_t* keep) {
  if (n == 0) return 0;
  int32_t* order = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
  float* areas = (float*)malloc(sizeof(float) * (size_t)n);
  uint8_t* select = (uint8_t*)malloc((size_t)n);
  const float off = (float)offset;
  for (int i = 0; i < n; ++i) {
    areas[i] = (boxes[i * 4 + 2] - boxes[i * 4] + off) * (boxes[i * 4 + 3] - boxes[i * 4 + 1] + off);
    select[i] = 1;
  }
  stable_argsort_desc(scores, n, order);
  for (int _i = 0; _i < n; ++_i) {
    if (!select[_i]) continue;
    const int i = order[_i];
    const float ix1 = boxes[i * 4], iy1 = boxes[i * 4 + 1], ix2 = boxes[i * 4 + 2], iy2 = boxes[i * 4 + 3];
    const float iarea = areas[i];
    for (int _j = _i + 1; _j < n; ++_j) {
      if (!select[_j]) continue;
      const int j = order[_j];
      const float xx1 = fmaxf(ix1, boxes[j * 4]), yy1 = fmaxf(iy1, boxes[j * 4 + 1]);
      const float xx2 = fminf(ix2, boxes[j * 4 + 2]), yy2 = fminf(iy2, boxes[j * 4 + 3]);
      const float w = fmaxf(0.f, xx2 - xx1 + off), h = fmaxf(0.f, yy2 - yy1 + off);
      const float inter = w * h;
      const float ovr = inter / (iarea + areas[j] - inter);
      if (ovr > thr) select[_j] = 0;
    }
  }
  int nk = 0;
  for (int _i = 0; _i < n; ++_i)
    if (select[_i]) keep[nk++] = order[_i];
  free(order); free(areas); free(select);
  return nk;
}

/* mmcv.ops.batched_nms (App. B): offsets = idxs.to(boxes)*(boxes.max()+1);
 * below split_thr one nms over offset boxes, else per-id nms + re-sort.
 * idxs == NULL -> class agnostic.  Returns #keep; keep[] are input indices in
 * output order; dets (optional) rows cat(boxes[keep], scores[keep]).       */
API int oracle_batched_nms(const float* boxes, const float* scores, const int64_t* idxs,
                           int n, float thr, int split_thr, int64_t* keep, float* dets) {
  if (n == 0) return 0;
  float* bn = (float*)malloc(sizeof(float) * 4 * (size_t)n);
  if (idxs) {
    float maxc = boxes[0];
    for (int i = 1; i < 4 * n; ++i) if (boxes[i] > maxc) maxc = boxes[i];
    const float m1 = maxc + 1.0f;
    for (int i = 0; i < n; ++i) {
      const float o = (float)idxs[i] * m1;
      for (int j = 0; j < 4; ++j) bn[i * 4 + j] = boxes[i * 4 + j] + o;
    }
  } else {
    memcpy(bn, boxes, sizeof(float) * 4 * (size_t)n);
  }
  int nk = 0;
  if (n < split_thr) {
    nk = oracle_nms_cpu(bn, scores, n, thr, 0, keep);
  } else {
    /* per unique id (ascending), then keep = nonzero(mask) re-sorted by score */
    uint8_t* total = (uint8_t*)calloc((size_t)n, 1);
    int64_t maxid = 0, minid = 0;
    if (idxs) { maxid = minid = idxs[0]; for (int i = 1; i < n; ++i) { if (idxs[i] > maxid) maxid = idxs[i]; if (idxs[i] < minid) minid = idxs[i]; } }
    int32_t* sub = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    float* sb = (float*)malloc(sizeof(float) * 4 * (size_t)n);
    float* ss = (float*)malloc(sizeof(float) * (size_t)n);
    int64_t* sk = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
    for (int64_t id = minid; id <= maxid; ++id) {
      int m = 0;
      for (int i = 0; i < n; ++i)
        if (!idxs || idxs[i] == id) { sub[m] = i; memcpy(sb + 4 * m, bn + 4 * i, 16); ss[m] = scores[i]; ++m; }
      if (m == 0) continue;
      const int k = oracle_nms_cpu(sb, ss, m, thr, 0, sk);
      for (int t = 0; t < k; ++t) total[sub[sk[t]]] = 1;
    }
    int m = 0;
    for (int i = 0; i < n; ++i) if (total[i]) { sub[m] = i; ss[m] = scores[i]; ++m; }
    int32_t* ord = (int32_t*)malloc(sizeof(int32_t) * (size_t)(m > 0 ? m : 1));
    stable_argsort_desc(ss, m, ord);
    for (int t = 0; t < m; ++t) keep[t] = sub[ord[t]];
    nk = m;
    free(total); free(sub); free(sb); free(ss); free(sk); free(ord);
  }
  if (dets)
    for (int t = 0; t < nk; ++t) {
      memcpy(dets + 5 * t, boxes + 4 * keep[t], 16);
      dets[5 * t + 4] = scores[keep[t]];
    }
  free(bn);
  return nk;
}

/* ------------------------------------------------------- RPN get_bboxes --- */
/* ATSSRPNHead._get_bboxes_single (atss_rpn_head.py:688-760) for one image.
 * cls[l]: (A,H,W), bbox[l]: (4A,H,W), iou[l]: (A,H,W); base_anchors (L,A,4).
 * Outputs: proposals (max_per_img,5) + return count; optional debug:
 *   topk_idx (K) int32 anchor index within its level, levels concatenated in
 *   rank order; cand_boxes (K,4) decoded; cand_scores (K); cand_n[L] per-level
 *   counts (K = sum cand_n), all BEFORE the min-size filter.                 */
API int oracle_rpn_get_bboxes_single(
    int L, int A, const int32_t* feat_h, const int32_t* feat_w, const int32_t* stride_w,
    const int32_t* stride_h, const float* const* cls, const float* const* bbox,
    const float* const* iou, const float* base_anchors, float img_h, float img_w,
    int nms_pre, int max_per_img, float iou_thr, float min_bbox_size, const float* means,
    const float* stds, float max_ratio, int split_thr, float* proposals, int32_t* topk_idx,
    float* cand_boxes_out, float* cand_scores_out, int32_t* cand_n) {
  int total_cap = 0;
  for (int l = 0; l < L; ++l) {
    int n = feat_h[l] * feat_w[l] * A;
    total_cap += (nms_pre > 0 && n > nms_pre) ? nms_pre : n;
  }
  float* boxes = (float*)malloc(sizeof(float) * 4 * (size_t)total_cap);
  float* scores = (float*)malloc(sizeof(float) * (size_t)total_cap);
  int64_t* ids = (int64_t*)malloc(sizeof(int64_t) * (size_t)total_cap);
  int K = 0;
  for (int l = 0; l < L; ++l) {
    const int H = feat_h[l], W = feat_w[l], P = H * W, n = P * A;
    float* s = (float*)malloc(sizeof(float) * (size_t)n);
    /* :711-725 permute(1,2,0).reshape(-1): k = (y*W+x)*A + a */
    for (int p = 0; p < P; ++p)
      for (int a = 0; a < A; ++a)
        s[p * A + a] = sqrtf(oracle_sigmoid(cls[l][(size_t)a * P + p]) *
                             oracle_sigmoid(iou[l][(size_t)a * P + p]));
    int k = n;
    int32_t* order = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    if (nms_pre > 0 && n > nms_pre) { /* :726-733 sort desc, first nms_pre */
      stable_argsort_desc(s, n, order);
      k = nms_pre;
    } else {
      for (int i = 0; i < n; ++i) order[i] = i;
    }
    float* anchors = (float*)malloc(sizeof(float) * 4 * (size_t)k);
    float* deltas = (float*)malloc(sizeof(float) * 4 * (size_t)k);
    for (int j = 0; j < k; ++j) {
      const int idx = order[j], p = idx / A, a = idx % A, y = p / W, x = p % W;
      const float sx = (float)(x * stride_w[l]), sy = (float)(y * stride_h[l]);
      const float* ba = base_anchors + ((size_t)l * A + a) * 4;
      /* anchor_generator.py:366-377 base_anchors + shifts */
      anchors[j * 4] = ba[0] + sx; anchors[j * 4 + 1] = ba[1] + sy;
      anchors[j * 4 + 2] = ba[2] + sx; anchors[j * 4 + 3] = ba[3] + sy;
      for (int c = 0; c < 4; ++c) deltas[j * 4 + c] = bbox[l][(size_t)(a * 4 + c) * P + p];
      scores[K + j] = s[idx];
      ids[K + j] = l;
      if (topk_idx) topk_idx[K + j] = idx;
    }
    oracle_delta2bbox(anchors, deltas, k, 1, means, stds, max_ratio, img_h, img_w, boxes + 4 * (size_t)K);
    if (cand_n) cand_n[l] = k;
    K += k;
    free(s); free(order); free(anchors); free(deltas);
  }
  if (cand_boxes_out) memcpy(cand_boxes_out, boxes, sizeof(float) * 4 * (size_t)K);
  if (cand_scores_out) memcpy(cand_scores_out, scores, sizeof(float) * (size_t)K);
  /* :747-754 min size filter */
  int Kv = K;
  if (min_bbox_size >= 0.f) {
    Kv = 0;
    for (int i = 0; i < K; ++i) {
      const float w = boxes[i * 4 + 2] - boxes[i * 4], h = boxes[i * 4 + 3] - boxes[i * 4 + 1];
      if (w > min_bbox_size && h > min_bbox_size) {
        if (Kv != i) { memcpy(boxes + 4 * Kv, boxes + 4 * i, 16); scores[Kv] = scores[i]; ids[Kv] = ids[i]; }
        ++Kv;
      }
    }
  }
  int nout = 0;
  if (Kv > 0) { /* :755-760 */
    int64_t* keep = (int64_t*)malloc(sizeof(int64_t) * (size_t)Kv);
    float* dets = (float*)malloc(sizeof(float) * 5 * (size_t)Kv);
    const int nk = oracle_batched_nms(boxes, scores, ids, Kv, iou_thr, split_thr, keep, dets);
    nout = nk < max_per_img ? nk : max_per_img;
    memcpy(proposals, dets, sizeof(float) * 5 * (size_t)nout);
    free(keep); free(dets);
  }
  free(boxes); free(scores); free(ids);
  return nout;
}

/* ------------------------------------------------------- RoI extractor --- */
/* map_roi_levels (single_level_roi_extractor.py:51-54), compare form */
API void oracle_map_roi_levels(const float* rois, int R, float finest_scale, int L, int64_t* out) {
  for (int r = 0; r < R; ++r) {
    const float* q = rois + (size_t)r * 5;
    const float scale = sqrtf((q[3] - q[1]) * (q[4] - q[2]));
    const float v = scale / finest_scale + 1e-6f;
    /* floor(log2f(v)) >= k  <=>  v >= T[k] for a round-to-nearest log2f: the
     * 1-2 floats just below 2^k (k >= 3) already round UP to k in fp32.
     * Pinned against torch.log2 (CPU) in tests/test_oracle_golden.py. */
    static const uint32_t T[8] = {0u,          0x40000000u, 0x40800000u, 0x40FFFFFFu,
                                  0x417FFFFFu, 0x41FFFFFEu, 0x427FFFFEu, 0x42FFFFFEu};
    int lvl = 0;
    while (lvl < L - 1 && lvl < 7 && v >= as_float((int32_t)T[lvl + 1])) ++lvl;
    out[r] = lvl;
  }
}

/* mmcv roi_align bilinear_interpolate (one axis; the 2-D version is the
 * product): returns 0 if outside [-1,size] */
static inline int axis_taps(float v, int size, int* lo, int* hi, float* wl, float* wh) {
  if (v < -1.0f || v > (float)size) return 0;
  if (v <= 0.f) v = 0.f;
  *lo = (int)v;
  if (*lo >= size - 1) { *hi = *lo = size - 1; v = (float)*lo; } else { *hi = *lo + 1; }
  *wh = v - (float)*lo;
  *wl = 1.0f - *wh;
  return 1;
}

/* mmcv roi_align_forward, pool_mode avg (App. A6).  input (N,C,H,W) NCHW;
 * rois (R,5); output (R,C,ph,pw). */
API void oracle_roi_align_forward(const float* input, int N, int C, int H, int W,
                                  const float* rois, int R, int PH, int PW,
                                  float spatial_scale, int sampling_ratio, int aligned,
                                  float* output) {
  (void)N;
  for (int r = 0; r < R; ++r) {
    const float* q = rois + (size_t)r * 5;
    const int b = (int)q[0];
    const float off = aligned ? 0.5f : 0.f;
    const float sw = q[1] * spatial_scale - off, sh = q[2] * spatial_scale - off;
    const float ew = q[3] * spatial_scale - off, eh = q[4] * spatial_scale - off;
    float rw = ew - sw, rh = eh - sh;
    if (!aligned) { rw = fmaxf(rw, 1.f); rh = fmaxf(rh, 1.f); }
    const float bin_h = rh / (float)PH, bin_w = rw / (float)PW;
    const int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rh / (float)PH);
    const int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rw / (float)PW);
    const float count = fmaxf((float)(gh * gw), 1.0f);
    for (int c = 0; c < C; ++c) {
      const float* f = input + ((size_t)b * C + c) * H * W;
      for (int ph = 0; ph < PH; ++ph)
        for (int pw = 0; pw < PW; ++pw) {
          float acc = 0.f;
          for (int iy = 0; iy < gh; ++iy) {
            const float y = sh + (float)ph * bin_h + ((float)iy + .5f) * bin_h / (float)gh;
            int yl, yh; float hy, ly;
            if (!axis_taps(y, H, &yl, &yh, &hy, &ly)) continue;
            for (int ix = 0; ix < gw; ++ix) {
              const float x = sw + (float)pw * bin_w + ((float)ix + .5f) * bin_w / (float)gw;
              int xl, xh; float hx, lx;
              if (!axis_taps(x, W, &xl, &xh, &hx, &lx)) continue;
              const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
              acc += w1 * f[yl * W + xl] + w2 * f[yl * W + xh] + w3 * f[yh * W + xl] + w4 * f[yh * W + xh];
            }
          }
          output[(((size_t)r * C + c) * PH + ph) * PW + pw] = acc / count;
        }
    }
  }
}

/* mmcv roi_align_backward (avg): grad_input must be zero-initialised. */
API void oracle_roi_align_backward(const float* grad_output, int N, int C, int H, int W,
                                   const float* rois, int R, int PH, int PW,
                                   float spatial_scale, int sampling_ratio, int aligned,
                                   float* grad_input) {
  (void)N;
  for (int r = 0; r < R; ++r) {
    const float* q = rois + (size_t)r * 5;
    const int b = (int)q[0];
    const float off = aligned ? 0.5f : 0.f;
    const float sw = q[1] * spatial_scale - off, sh = q[2] * spatial_scale - off;
    const float ew = q[3] * spatial_scale - off, eh = q[4] * spatial_scale - off;
    float rw = ew - sw, rh = eh - sh;
    if (!aligned) { rw = fmaxf(rw, 1.f); rh = fmaxf(rh, 1.f); }
    const float bin_h = rh / (float)PH, bin_w = rw / (float)PW;
    const int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rh / (float)PH);
    const int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rw / (float)PW);
    const float count = fmaxf((float)(gh * gw), 1.0f);
    for (int c = 0; c < C; ++c) {
      float* gi = grad_input + ((size_t)b * C + c) * H * W;
      for (int ph = 0; ph < PH; ++ph)
        for (int pw = 0; pw < PW; ++pw) {
          const float g = grad_output[(((size_t)r * C + c) * PH + ph) * PW + pw];
          for (int iy = 0; iy < gh; ++iy) {
            const float y = sh + (float)ph * bin_h + ((float)iy + .5f) * bin_h / (float)gh;
            int yl, yh; float hy, ly;
            if (!axis_taps(y, H, &yl, &yh, &hy, &ly)) continue;
            for (int ix = 0; ix < gw; ++ix) {
              const float x = sw + (float)pw * bin_w + ((float)ix + .5f) * bin_w / (float)gw;
              int xl, xh; float hx, lx;
              if (!axis_taps(x, W, &xl, &xh, &hx, &lx)) continue;
              gi[yl * W + xl] += g * (hy * hx) / count;
              gi[yl * W + xh] += g * (hy * lx) / count;
              gi[yh * W + xl] += g * (ly * hx) / count;
              gi[yh * W + xh] += g * (ly * lx) / count;
            }
          }
        }
    }
  }
}

/* SingleRoIExtractor.forward (single_level_roi_extractor.py:57-115): level
 * map, per-level RoIAlign, scatter back.  feats[l]: (B,C,H_l,W_l) NCHW. */
API void oracle_roi_extract_forward(int L, int B, int C, const int32_t* feat_h,
                                    const int32_t* feat_w, const float* spatial_scale,
                                    const float* const* feats, const float* rois, int R,
                                    int PH, int PW, int sampling_ratio, int aligned,
                                    float finest_scale, float* out, int64_t* lvls_out) {
  int64_t* lv = (int64_t*)malloc(sizeof(int64_t) * (size_t)(R > 0 ? R : 1));
  oracle_map_roi_levels(rois, R, finest_scale, L, lv);
  const size_t per = (size_t)C * PH * PW;
  for (int r = 0; r < R; ++r) {
    const int l = (int)lv[r];
    if (rois[(size_t)r * 5] < 0.f) { memset(out + r * per, 0, per * 4); if (lvls_out) lvls_out[r] = -1; continue; }
    oracle_roi_align_forward(feats[l], B, C, feat_h[l], feat_w[l], rois + (size_t)r * 5, 1, PH, PW,
                             spatial_scale[l], sampling_ratio, aligned, out + r * per);
    if (lvls_out) lvls_out[r] = l;
  }
  free(lv);
}

/* autograd of the above: grad_feats[l] (B,C,H_l,W_l) zero-initialised here */
API void oracle_roi_extract_backward(int L, int B, int C, const int32_t* feat_h,
                                     const int32_t* feat_w, const float* spatial_scale,
                                     const float* grad_out, const float* rois, int R, int PH,
                                     int PW, int sampling_ratio, int aligned,
                                     float finest_scale, float* const* grad_feats) {
  int64_t* lv = (int64_t*)malloc(sizeof(int64_t) * (size_t)(R > 0 ? R : 1));
  oracle_map_roi_levels(rois, R, finest_scale, L, lv);
  for (int l = 0; l < L; ++l)
    memset(grad_feats[l], 0, sizeof(float) * (size_t)B * C * feat_h[l] * feat_w[l]);
  const size_t per = (size_t)C * PH * PW;
  for (int r = 0; r < R; ++r) {
    if (rois[(size_t)r * 5] < 0.f) continue;
    const int l = (int)lv[r];
    oracle_roi_align_backward(grad_out + r * per, B, C, feat_h[l], feat_w[l], rois + (size_t)r * 5, 1,
                              PH, PW, spatial_scale[l], sampling_ratio, aligned, grad_feats[l]);
  }
  free(lv);
}

/* ------------------------------------------------------ RCNN test path --- */
/* prob_roi_head.py:232-240: softmax -> * prior -> **0.5.  Class sum in index
 * order (pinned).  cls_score (R,C1) -> scores (R,C1). */
API void oracle_fuse_scores(const float* cls_score, const float* prior, int R, int C1,
                            int prob, float* scores) {
  for (int r = 0; r < R; ++r) {
    const float* x = cls_score + (size_t)r * C1;
    float* o = scores + (size_t)r * C1;
    if (!prob) { memcpy(o, x, sizeof(float) * (size_t)C1); continue; }
    float m = x[0];
    for (int c = 1; c < C1; ++c) if (x[c] > m) m = x[c];
    float sum = 0.f;
    for (int c = 0; c < C1; ++c) { o[c] = oracle_expf(x[c] - m); sum += o[c]; }
    for (int c = 0; c < C1; ++c) o[c] = sqrtf((o[c] / sum) * prior[r]);
  }
}

/* ProbConvFCBBoxHead.get_bboxes (convfc_bbox_head.py:294-330) +
 * multiclass_nms (bbox_nms.py:8-95) for ONE image.  rois (R,5), scores (R,C+1)
 * already fused, bbox_pred (R,4C) (or (R,4) if agnostic).
 * Outputs: det_bboxes (max_per_img,5), det_labels, return #dets;
 * optional decoded (R, nbox*4) boxes and keep_flat (flat index roi*C+c). */
API int oracle_rcnn_get_bboxes_single(const float* rois, const float* scores,
                                      const float* bbox_pred, int R, int C, int agnostic,
                                      const float* means, const float* stds, float max_ratio,
                                      float img_h, float img_w, const float* scale_factor,
                                      int rescale, float score_thr, float iou_thr,
                                      int max_per_img, int split_thr, float* det_bboxes,
                                      int64_t* det_labels, float* decoded_out,
                                      int64_t* keep_flat) {
  const int nbox = agnostic ? 1 : C;
  float* r4 = (float*)malloc(sizeof(float) * 4 * (size_t)(R > 0 ? R : 1));
  for (int r = 0; r < R; ++r) memcpy(r4 + 4 * r, rois + 5 * (size_t)r + 1, 16);
  float* dec = (float*)malloc(sizeof(float) * 4 * (size_t)(R > 0 ? R : 1) * nbox);
  oracle_delta2bbox(r4, bbox_pred, R, nbox, means, stds, max_ratio, img_h, img_w, dec);
  if (rescale) /* :318-321 bboxes.view(n,-1,4) / scale_factor */
    for (size_t i = 0; i < (size_t)R * nbox; ++i)
      for (int j = 0; j < 4; ++j) dec[i * 4 + j] = dec[i * 4 + j] / scale_factor[j];
  if (decoded_out) memcpy(decoded_out, dec, sizeof(float) * 4 * (size_t)R * nbox);
  /* bbox_nms.py:43-68 */
  const int C1 = C + 1;
  int nc = 0;
  for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) if (scores[(size_t)r * C1 + c] > score_thr) ++nc;
  int nout = 0;
  if (nc > 0) {
    float* cb = (float*)malloc(sizeof(float) * 4 * (size_t)nc);
    float* cs = (float*)malloc(sizeof(float) * (size_t)nc);
    int64_t* cl = (int64_t*)malloc(sizeof(int64_t) * (size_t)nc);
    int64_t* cf = (int64_t*)malloc(sizeof(int64_t) * (size_t)nc);
    int m = 0;
    for (int r = 0; r < R; ++r)
      for (int c = 0; c < C; ++c)
        if (scores[(size_t)r * C1 + c] > score_thr) {
          memcpy(cb + 4 * m, dec + ((size_t)r * nbox + (agnostic ? 0 : c)) * 4, 16);
          cs[m] = scores[(size_t)r * C1 + c]; cl[m] = c; cf[m] = (int64_t)r * C + c; ++m;
        }
    int64_t* keep = (int64_t*)malloc(sizeof(int64_t) * (size_t)nc);
    float* dets = (float*)malloc(sizeof(float) * 5 * (size_t)nc);
    const int nk = oracle_batched_nms(cb, cs, cl, nc, iou_thr, split_thr, keep, dets);
    nout = (max_per_img > 0 && nk > max_per_img) ? max_per_img : nk;
    for (int t = 0; t < nout; ++t) {
      memcpy(det_bboxes + 5 * t, dets + 5 * t, 20);
      det_labels[t] = cl[keep[t]];
      if (keep_flat) keep_flat[t] = cf[keep[t]];
    }
    free(cb); free(cs); free(cl); free(cf); free(keep); free(dets);
  }
  free(r4); free(dec);
  return nout;
}

/* -------------------------------------------------------- boost loss --- */
/* prob_roi_head.py:107-154 + convfc_bbox_head.py:332-418 (App. A8).
 * out[8] like brcnn_boost_loss; gradients for upstream grad 1. */
API void oracle_boost_loss(const float* cls_score, const int64_t* labels,
                           const float* label_weights, const float* prior,
                           const float* bbox_pred, const float* bbox_targets,
                           const float* bbox_weights, int N, int C, int agnostic,
                           float gamma, float alpha, float wcls, float wbbox,
                           int reg_norm_mean, float* out, float* grad_cls, float* grad_bbox) {
  const int C1 = C + 1;
  float* l = (float*)malloc(sizeof(float) * (size_t)(N > 0 ? N : 1));
  float* w = (float*)malloc(sizeof(float) * (size_t)(N > 0 ? N : 1));
  double sl = 0.0, swl = 0.0, slb = 0.0; /* accumulate in double: value oracle */
  int correct = 0, npos = 0;
  for (int i = 0; i < N; ++i) {
    const float* x = cls_score + (size_t)i * C1;
    float m = x[0]; int am = 0;
    for (int c = 1; c < C1; ++c) if (x[c] > m) { m = x[c]; am = c; }
    double sum = 0.0;
    for (int c = 0; c < C1; ++c) sum += exp((double)x[c] - (double)m);
    const int64_t lab = labels[i];
    const float lw = label_weights ? label_weights[i] : 1.0f;
    const float ce = (float)(((double)m + log(sum)) - (double)x[lab]);
    l[i] = wcls * (ce * lw);                 /* loss_weight * (CE * weight) */
    const float base = 1.0f - prior[i];
    w[i] = powf(base, gamma);                /* :125-126 */
    if (alpha != 0.f) w[i] = w[i] * alpha;
    sl += l[i]; swl += (double)w[i] * l[i];
    correct += (am == lab);
    for (int c = 0; c < C1; ++c) grad_cls[(size_t)i * C1 + c] = (float)(exp((double)x[c] - (double)m) / sum);
  }
  const double s = sl / swl; /* norm_loss :151-154 */
  for (int i = 0; i < N; ++i) {
    const float lw = label_weights ? label_weights[i] : 1.0f;
    const double coef = (double)w[i] * s / (double)N * wcls * lw;
    for (int c = 0; c < C1; ++c) {
      const double p = grad_cls[(size_t)i * C1 + c];
      grad_cls[(size_t)i * C1 + c] = (float)(coef * (p - (c == labels[i] ? 1.0 : 0.0)));
    }
  }
  const int nb = agnostic ? 4 : 4 * C;
  if (grad_bbox) memset(grad_bbox, 0, sizeof(float) * (size_t)N * nb);
  for (int i = 0; i < N; ++i) if (labels[i] >= 0 && labels[i] < C) ++npos;
  const double bden = reg_norm_mean ? (npos > 0 ? npos * 4.0 : 1.0) : (double)N;
  for (int i = 0; i < N; ++i) {
    const int64_t lab = labels[i];
    if (!(lab >= 0 && lab < C)) continue;
    const size_t o = agnostic ? (size_t)i * 4 : ((size_t)i * C + lab) * 4;
    for (int j = 0; j < 4; ++j) {
      const float d = bbox_pred[o + j] - bbox_targets[(size_t)i * 4 + j];
      slb += wbbox * (fabsf(d) * bbox_weights[(size_t)i * 4 + j]);
      if (grad_bbox) {
        const float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
        grad_bbox[o + j] = (float)(wbbox * bbox_weights[(size_t)i * 4 + j] * sg / bden);
      }
    }
  }
  out[0] = N > 0 ? (float)(s * swl / (double)N) : 0.f;
  out[1] = npos > 0 ? (float)(slb / bden) : 0.f;
  out[2] = N > 0 ? (float)(100.0 * correct / (double)N) : 0.f;
  out[3] = (float)sl; out[4] = (float)swl; out[5] = (float)npos; out[6] = (float)s; out[7] = 0.f;
  free(l); free(w);
}
