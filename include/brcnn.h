/*
 * libbrcnn — C ABI of the B200-native (sm_100a) Boosting R-CNN proposal-to-RoI
 * hot path.  This is the drop-in boundary: plain pointers and sizes, no torch
 * types.  Every entry point replaces a piece of the reference
 * (mousecpn/Boosting-R-CNN = mmdet 2.17 + mmcv-full 1.4.0); the replaced
 * interface is cited as file:line relative to the reference tree, or as the
 * mmcv `_ext` symbol it stands in for (mmcv source is not vendored in the
 * reference).
 *
 * Conventions
 *   - all tensor pointers are DEVICE pointers unless the name ends in _host or
 *     the comment says "host array"; the caller owns every buffer;
 *   - nothing is allocated, nothing synchronises, every launch goes to
 *     `stream` (graph-capturable); scratch lives in the caller's `workspace`
 *     (size from the matching *_workspace_bytes query, 256-byte aligned);
 *   - return value: 0 = ok, <0 = argument/size error (BRCNN_ERR_*), >0 = the
 *     cudaError_t of a failed launch;
 *   - fp32 data, int32 counts, int64 where the reference returns torch.long;
 *   - there is NO CPU fallback: without a CUDA device every compute entry
 *     point fails.
 */
#ifndef BRCNN_H_
#define BRCNN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BRCNN_OK 0
#define BRCNN_ERR_ARG (-1)
#define BRCNN_ERR_WORKSPACE (-2)
#define BRCNN_ERR_UNSUPPORTED (-3)

#define BRCNN_MAX_LEVELS 8
#define BRCNN_MAX_ANCHORS 32

typedef void* brcnn_stream_t; /* a cudaStream_t */

/* library / build identification: "libbrcnn <ver> sm_100a" */
const char* brcnn_version(void);
/* number of kernels launched by this library since load (all streams) */
int64_t brcnn_launch_count(void);

/* ------------------------------------------------------------------------
 * (1) RPN proposal generation
 * replaces ATSSRPNHead.get_bboxes / _get_bboxes_single
 *   (mmdet/models/dense_heads/atss_rpn_head.py:466-503, 688-760),
 *   AnchorGenerator.single_level_grid_anchors (anchor_generator.py:338-381),
 *   delta2bbox (delta_xywh_bbox_coder.py:144-272) and
 *   mmcv.ops.batched_nms with ids = pyramid level (atss_rpn_head.py:756).
 * ---------------------------------------------------------------------- */
typedef struct brcnn_rpn_params {
  int32_t batch;                       /* B images                           */
  int32_t num_levels;                  /* L pyramid levels                   */
  int32_t num_anchors;                 /* A base anchors per location        */
  int32_t feat_h[BRCNN_MAX_LEVELS];    /* H_l                                */
  int32_t feat_w[BRCNN_MAX_LEVELS];    /* W_l                                */
  int32_t stride_w[BRCNN_MAX_LEVELS];  /* anchor stride (w,h)                */
  int32_t stride_h[BRCNN_MAX_LEVELS];
  int32_t nms_pre;                     /* cfg.nms_pre (<=0: keep all)        */
  int32_t max_per_img;                 /* cfg.max_per_img                    */
  float iou_threshold;                 /* cfg.nms.iou_threshold              */
  float min_bbox_size;                 /* cfg.min_bbox_size (<0: no filter)  */
  float means[4];                      /* bbox_coder.target_means            */
  float stds[4];                       /* bbox_coder.target_stds             */
  float max_ratio;                     /* |ln(wh_ratio_clip)| as fp32        */
} brcnn_rpn_params;

/* byte offsets (into workspace) of the intermediate arrays, for parity tests.
 * Kc = per-level candidate capacity = min(nms_pre, max_l H_l*W_l*A).       */
typedef struct brcnn_rpn_ws_layout {
  int64_t cand_cap;     /* Kc                                                */
  int64_t keep_cap;     /* per-segment kept-list capacity                    */
  int64_t cand_boxes;   /* float  [B][L][Kc][4]  decoded + clipped           */
  int64_t cand_key;     /* uint64 [B][L][Kc]  score_bits<<32 | ~anchor_idx   */
  int64_t cand_valid;   /* uint8  [B][L][Kc]  passes the min-size filter     */
  int64_t cand_count;   /* int32  [B][L]                                     */
  int64_t img_maxc;     /* float  [B]  boxes.max() over valid candidates     */
  int64_t kept_pos;     /* int32  [B][L][keep_cap] candidate rank of keeps   */
  int64_t kept_count;   /* int32  [B][L]                                     */
  int64_t total_bytes;
} brcnn_rpn_ws_layout;

int brcnn_rpn_workspace_layout(const brcnn_rpn_params* p, brcnn_rpn_ws_layout* out);
size_t brcnn_rpn_workspace_bytes(const brcnn_rpn_params* p);

int brcnn_rpn_get_bboxes(
    const brcnn_rpn_params* p,
    const float* const* cls_scores_host, /* host array[L] of (B,A,H_l,W_l)   */
    const float* const* bbox_preds_host, /* host array[L] of (B,4A,H_l,W_l)  */
    const float* const* iou_preds_host,  /* host array[L] of (B,A,H_l,W_l)   */
    const float* base_anchors,           /* (L,A,4)                          */
    const float* img_hw,                 /* (B,2) img_shape[:2] = (h,w)      */
    float* proposals,                    /* (B,max_per_img,5) x1,y1,x2,y2,s  */
    int32_t* num_proposals,              /* (B)                              */
    void* workspace, size_t workspace_bytes, brcnn_stream_t stream);

/* delta2bbox as a standalone operator: DeltaXYWHBBoxCoder.decode
 * (mmdet/core/bbox/coder/delta_xywh_bbox_coder.py:63-95,144-272).
 * rois (n,4), deltas (n,4*ncls) -> out (n,4*ncls); max_h < 0: no clipping
 * (max_shape=None).  means/stds are HOST arrays of 4 floats.               */
int brcnn_delta2bbox(const float* rois, const float* deltas, int32_t n,
                     int32_t ncls, const float* means_host,
                     const float* stds_host, float max_ratio, float max_h,
                     float max_w, float* out, brcnn_stream_t stream);

/* ------------------------------------------------------------------------
 * (1b) RPN loss path: anchor targets + losses + gradients
 * replaces ATSSRPNHead.loss / loss_single / get_targets with atss=False
 *   (mmdet/models/dense_heads/atss_rpn_head.py:299-464, 505-603),
 *   AnchorHead.get_anchors / _get_targets_single (anchor_head.py:126-265),
 *   AnchorGenerator.grid_anchors / valid_flags (anchor_generator.py:338-434),
 *   MaxIoUAssigner with match_low_quality=True + PseudoSampler (the RPN train_cfg of
 *   configs/boosting_rcnn), FocalLoss (mmcv sigmoid_focal_loss, losses/focal_loss.py:86),
 *   IoULoss(mode='log'), MSELoss (aug_reg_loss), CrossEntropyLoss(use_sigmoid=True) on the IoU
 *   logit, weight_reduce_loss, for reg_decoded_bbox=True, num_classes=1, pos_weight=-1,
 *   allowed_border=-1, target means 0 / stds 1.
 * ---------------------------------------------------------------------- */
typedef struct brcnn_rpn_loss_params {
  int32_t batch, num_levels, num_anchors;
  int32_t feat_h[BRCNN_MAX_LEVELS], feat_w[BRCNN_MAX_LEVELS];
  int32_t stride_w[BRCNN_MAX_LEVELS], stride_h[BRCNN_MAX_LEVELS];
  int32_t max_gts;                 /* row capacity of gt_boxes per image (<= 1024)    */
  float pos_iou_thr, neg_iou_thr, min_pos_iou;   /* train_cfg.rpn.assigner              */
  float gamma;                     /* ATSSRPNHead.gamma: bbox weight = iou_target**gamma */
  float focal_gamma, focal_alpha;  /* loss_cls gamma / alpha                          */
  int32_t cls_loss_type;           /* 0 FocalLoss (sigmoid), 1 VarifocalLoss (iou_weighted,
                                      losses/varifocal_loss.py:10-57; every anchor counts)  */
  float loss_cls_weight, loss_bbox_weight, loss_iou_weight, loss_aug_weight;
  float max_ratio;                 /* |ln(wh_ratio_clip)| as fp32                     */
} brcnn_rpn_loss_params;

size_t brcnn_rpn_loss_workspace_bytes(const brcnn_rpn_loss_params* p);

/* sums: float[3L + 2] = L x sum of weighted focal terms | L x 0.5*(IoU-log + MSE) sums |
 *   L x BCE sums | num_total_pos | sum of iou_target -- all UN-normalised: the caller divides
 *   by max(reduce_mean(num_total_pos), 1) resp. max(reduce_mean(sum iou_target), 1)
 *   (atss_rpn_head.py:441-444, 458-460) on the device, after one fused all-reduce.
 * grad_*: un-normalised gradients, same shapes as the inputs, every element written.
 * pad_hw: (B,2) img_meta['pad_shape'][:2] (valid flags); gt_boxes (B,max_gts,4) + num_gt (B). */
int brcnn_rpn_loss_forward(const brcnn_rpn_loss_params* p,
                           const float* const* cls_scores_host,  /* host array[L] (B,A,H,W)   */
                           const float* const* bbox_preds_host,  /* host array[L] (B,4A,H,W)  */
                           const float* const* iou_preds_host,   /* host array[L] (B,A,H,W)   */
                           const float* base_anchors,            /* (L,A,4)                   */
                           const float* gt_boxes, const int32_t* num_gt, const float* pad_hw,
                           float* sums, float* const* grad_cls_host, float* const* grad_bbox_host,
                           float* const* grad_iou_host, void* workspace, size_t workspace_bytes,
                           brcnn_stream_t stream);

/* backward: out = raw * scale[kind * L + level], kind 0 = cls, 1 = bbox, 2 = iou; `scale` is a
 * DEVICE array of 3L floats (upstream gradient / normaliser).  out may alias raw. */
int brcnn_rpn_loss_scale(const brcnn_rpn_loss_params* p, const float* const* raw_cls_host,
                         const float* const* raw_bbox_host, const float* const* raw_iou_host,
                         const float* scale, float* const* out_cls_host,
                         float* const* out_bbox_host, float* const* out_iou_host,
                         brcnn_stream_t stream);

/* ------------------------------------------------------------------------
 * (2) NMS operators — stand-ins for mmcv `_ext.nms` and the Python wrappers
 * mmcv.ops.nms / mmcv.ops.batched_nms (call sites atss_rpn_head.py:756,
 * mmdet/core/post_processing/bbox_nms.py:86).  Order: score descending,
 * ties by lower input index; suppression test inter/(a+b-inter) > thr.
 * ---------------------------------------------------------------------- */
size_t brcnn_nms_workspace_bytes(int32_t num_boxes);

/* idxs may be NULL (plain / class-agnostic nms).  num_ids: caller's bound on the
 * id range (idxs in [0, num_ids)); <= 0 = unknown.  Three paths, same results:
 *   num_ids <= 8 (pyramid levels; or idxs == NULL): the ids are walked as sorted
 *     lists in global score order by a cluster of 8 CTAs (rpn_nms.cuh);
 *   8 < num_ids <= 1024 and num_boxes <= 8192 (class-wise NMS): one fused-NMS CTA
 *     per id on the id-sorted boxes, kept lists merged by an in-smem sort;
 *   otherwise: one segment of offset boxes (mmcv's formulation).
 * max_num: mmcv's `max_num` (nms(...) argument / nms_cfg key): only the first max_num keeps
 *   are wanted (<= 0: all); the sweep stops there instead of slicing afterwards.
 * keep: int64[num_boxes] (first *num_keep entries valid, score-descending),
 * dets: optional float[num_boxes][5] rows cat(boxes[keep], scores[keep]).     */
int brcnn_batched_nms(const float* boxes, const float* scores,
                      const int64_t* idxs, int32_t num_boxes, int32_t num_ids,
                      float iou_threshold, int32_t offset, int32_t max_num,
                      int64_t* keep, float* dets, int32_t* num_keep, void* workspace,
                      size_t workspace_bytes, brcnn_stream_t stream);

/* ------------------------------------------------------------------------
 * (3) RoI feature extraction
 * replaces SingleRoIExtractor.map_roi_levels / forward
 *   (mmdet/models/roi_heads/roi_extractors/single_level_roi_extractor.py:36-115)
 * and mmcv `_ext.roi_align_forward` / `roi_align_backward`
 *   (built at base_roi_extractor.py:54-59), pool_mode = avg.
 * Feature maps are NHWC (channels_last) fp32: (B, H_l, W_l, C).
 * ---------------------------------------------------------------------- */
typedef struct brcnn_roi_params {
  int32_t batch;
  int32_t channels;                      /* C, multiple of 4                 */
  int32_t num_levels;
  int32_t feat_h[BRCNN_MAX_LEVELS];
  int32_t feat_w[BRCNN_MAX_LEVELS];
  float spatial_scale[BRCNN_MAX_LEVELS]; /* 1/stride                         */
  int32_t pooled_h, pooled_w;            /* roi_layer.output_size            */
  int32_t sampling_ratio;                /* 0 = adaptive ceil(roi/pooled)    */
  int32_t aligned;                       /* mmcv default 1                   */
  float finest_scale;                    /* 56                               */
  int32_t out_layout;                    /* memory order of `out` / `grad_out`:
                                            0 = (R,C,ph,pw) as mmcv RoIAlign returns it,
                                            1 = (R,ph,pw,C): the channels-last RoI feature
                                            hand-off to the first FC (permuted weight columns,
                                            convfc_bbox_head.py:164); forward epilogue and
                                            backward then need no transpose               */
} brcnn_roi_params;

/* bbox2roi (mmdet/core/bbox/transforms.py:59-78) on the padded proposal
 * layout of brcnn_rpn_get_bboxes: proposals (B,cap,5) + num (B) ->
 * rois (B*cap,5) [b,x1,y1,x2,y2] with b = -1 on padding rows, and (optional)
 * prior (B*cap) = proposals[..., 4] (prob_roi_head.py:214).                  */
int brcnn_bbox2roi_padded(const float* proposals, const int32_t* num_proposals,
                          int32_t batch, int32_t cap, float* rois, float* prior,
                          brcnn_stream_t stream);

/* target_lvls of map_roi_levels: int64[R] */
int brcnn_map_roi_levels(const float* rois, int32_t num_rois,
                         float finest_scale, int32_t num_levels,
                         int64_t* target_lvls, brcnn_stream_t stream);

/* rois: (R,5) [batch_idx, x1,y1,x2,y2]; rows with batch_idx < 0 are padding
 * and produce zeros.  out: (R, C, pooled_h, pooled_w) NCHW-contiguous, the
 * layout ConvFCBBoxHead flattens (convfc_bbox_head.py:164), or
 * (R, pooled_h, pooled_w, C) when p->out_layout == 1.
 * roi_levels: optional int32[R] (level used for each RoI).
 * workspace: optional scheduling scratch of brcnn_roi_extract_forward_workspace_bytes
 * bytes, ZERO-FILLED by the caller once; every call leaves it zero-filled, so one
 * buffer serves all calls issued on the same stream (calls that may run concurrently
 * need a buffer each).  With it the persistent CTAs pull RoIs from an atomic counter
 * (RoI cost varies ~4x with the footprint); NULL selects a static round-robin
 * schedule - same results bit for bit, longer tail.                          */
size_t brcnn_roi_extract_forward_workspace_bytes(const brcnn_roi_params* p);
int brcnn_roi_extract_forward(const brcnn_roi_params* p,
                              const float* const* feats_nhwc_host,
                              const float* rois, int32_t num_rois, float* out,
                              int32_t* roi_levels, void* workspace,
                              size_t workspace_bytes, brcnn_stream_t stream);

size_t brcnn_roi_extract_backward_workspace_bytes(const brcnn_roi_params* p,
                                                  int32_t num_rois);
/* grad_feats_nhwc: host array[L] of (B,H_l,W_l,C) buffers; every element is
 * written exactly once (levels without RoIs get zeros), no float atomics,
 * bit-reproducible run to run.  grad_out is in the layout p->out_layout names. */
int brcnn_roi_extract_backward(const brcnn_roi_params* p,
                               const float* grad_out, const float* rois,
                               int32_t num_rois,
                               float* const* grad_feats_nhwc_host,
                               void* workspace, size_t workspace_bytes,
                               brcnn_stream_t stream);

/* layout helpers for callers that hold NCHW maps (the reference neck's
 * default): (B,C,H,W) <-> (B,H,W,C), fp32                                   */
int brcnn_nchw_to_nhwc(const float* in, float* out, int32_t batch,
                       int32_t channels, int32_t hw, brcnn_stream_t stream);
int brcnn_nhwc_to_nchw(const float* in, float* out, int32_t batch,
                       int32_t channels, int32_t hw, brcnn_stream_t stream);
/* the same for up to BRCNN_MAX_LEVELS maps in ONE launch (the whole pyramid):
 * host arrays of device pointers; map i is (batch, channels, hw_host[i])    */
int brcnn_nchw_to_nhwc_multi(const float* const* in_host, float* const* out_host,
                             int32_t num_maps, int32_t batch, int32_t channels,
                             const int32_t* hw_host, brcnn_stream_t stream);
int brcnn_nhwc_to_nchw_multi(const float* const* in_host, float* const* out_host,
                             int32_t num_maps, int32_t batch, int32_t channels,
                             const int32_t* hw_host, brcnn_stream_t stream);

/* ------------------------------------------------------------------------
 * (4) Boosting reweighted R-CNN loss, forward + gradient in one pass
 * replaces ProbRoIHead._bbox_forward_train_boost / norm_loss
 *   (mmdet/models/roi_heads/prob_roi_head.py:107-154),
 *   ProbConvFCBBoxHead.loss (bbox_heads/convfc_bbox_head.py:332-418),
 *   cross_entropy (losses/cross_entropy_loss.py:10-50), weight_reduce_loss
 *   (losses/utils.py:28-55), L1Loss (losses/smooth_l1_loss.py:35-52) and
 *   accuracy (losses/accuracy.py:6-51).
 * ---------------------------------------------------------------------- */
typedef struct brcnn_loss_params {
  int32_t num_rois;            /* N sampled RoIs                             */
  int32_t num_classes;         /* C foreground classes; bg label == C        */
  int32_t reg_class_agnostic;  /* bbox_pred (N,4) instead of (N,4C)          */
  float gamma;                 /* ProbRoIHead.gamma: w = (1-prior)^gamma     */
  float alpha;                 /* ProbRoIHead.alpha (0 = off)                */
  float loss_cls_weight;       /* CrossEntropyLoss.loss_weight               */
  float loss_bbox_weight;      /* L1Loss.loss_weight                         */
  int32_t reg_norm_mean;       /* reg_norm == 'mean'                         */
} brcnn_loss_params;

/* out_scalars: float[8] = {loss_cls, loss_bbox, acc, sum_l, sum_wl, n_pos,
 *                          weight_scale s, 0}
 * grad_cls_score (N,C+1) and grad_bbox_pred (N,4C | N,4) are d(loss_cls)/d.
 * and d(loss_bbox)/d. for an upstream gradient of 1.                       */
size_t brcnn_boost_loss_workspace_bytes(const brcnn_loss_params* p);
int brcnn_boost_loss(const brcnn_loss_params* p, const float* cls_score,
                     const int64_t* labels, const float* label_weights,
                     const float* prior, const float* bbox_pred,
                     const float* bbox_targets, const float* bbox_weights,
                     float* out_scalars, float* grad_cls_score,
                     float* grad_bbox_pred, void* workspace, size_t workspace_bytes,
                     brcnn_stream_t stream);

/* ------------------------------------------------------------------------
 * (4b) R-CNN training front-end: assign + sample + targets + prior
 * (the caller side of the RoI kernels)
 * replaces MaxIoUAssigner.assign (core/bbox/assigners/max_iou_assigner.py:61-212,
 *   match_low_quality=False as in every boosting_rcnn R-CNN train_cfg),
 *   AssignResult.add_gt_ (assign_result.py:191-205), the index bookkeeping of
 *   RandomSampler.sample / SamplingResult (samplers/random_sampler.py:32-82,
 *   base_sampler.py:35-102, sampling_result.py:26-55), BBoxHead.get_targets
 *   (roi_heads/bbox_heads/bbox_head.py:122-253), bbox2delta
 *   (coder/delta_xywh_bbox_coder.py:98-141) and the prior vector of
 *   ProbRoIHead.forward_train (roi_heads/prob_roi_head.py:51-64).
 * The random permutations stay on the host: the reference draws them with
 * torch.randperm on the CPU generator (random_sampler.py:58).
 * Proposals are in the padded layout of brcnn_rpn_get_bboxes.
 * ---------------------------------------------------------------------- */
typedef struct brcnn_assign_params {
  int32_t batch;
  int32_t max_props;           /* M: row capacity of proposals per image       */
  int32_t max_gts;             /* Gmax: row capacity of gt_boxes per image     */
  float pos_iou_thr, neg_iou_thr, min_pos_iou;
  int32_t match_low_quality;   /* must be 0 (else BRCNN_ERR_UNSUPPORTED)       */
} brcnn_assign_params;

/* gt_inds: int32 (B, Gmax + M) in the sampler's index space after add_gt_
 *   (slot g < num_gt[b]: GT g, value g+1; slot num_gt[b] + j: proposal j;
 *   values: >0 assigned GT index + 1, 0 background, -1 ignore, -2 unused slot)
 * counts:  int32 (B, 2) = number of positive / negative candidates (zeroed here) */
int brcnn_rcnn_assign(const brcnn_assign_params* p, const float* proposals,
                      const int32_t* num_props, const float* gt_boxes,
                      const int32_t* num_gt, int32_t* gt_inds, int32_t* counts,
                      brcnn_stream_t stream);

typedef struct brcnn_sample_params {
  int32_t batch, max_props, max_gts;
  int32_t num_classes;         /* background label                             */
  int32_t perm_cap;            /* columns of perm_pos / perm_neg               */
  int32_t max_sel;             /* max rows selected per image per list         */
  float means[4], stds[4];     /* bbox_coder target_means / target_stds        */
  float pos_weight;            /* label weight of positives (1 if cfg <= 0)    */
} brcnn_sample_params;

/* plan: int32 (B,5) written by the host from `counts`:
 *   [n_pos_selected, n_neg_selected, first output row, use perm_pos, use perm_neg]
 * perm_pos / perm_neg: int32 (B, perm_cap), torch.randperm(n_candidates)[:n_selected]
 *   when the candidate list is longer than the quota (use flag = 1).
 * Output rows per image: positives then negatives, each index-ascending
 * (SamplingResult.bboxes): rois (N,5), labels int64 (N), label_weights (N),
 * bbox_targets (N,4), bbox_weights (N,4), prior (N).                          */
int brcnn_rcnn_sample_targets(const brcnn_sample_params* p, const float* proposals,
                              const int32_t* num_props, const float* gt_boxes,
                              const int64_t* gt_labels, const int32_t* num_gt,
                              const int32_t* gt_inds, const int32_t* plan,
                              const int32_t* perm_pos, const int32_t* perm_neg,
                              float* rois, int64_t* labels, float* label_weights,
                              float* bbox_targets, float* bbox_weights, float* prior,
                              brcnn_stream_t stream);

/* ------------------------------------------------------------------------
 * (5) Probabilistic score fusion + per-class decode + class-wise NMS
 * replaces ProbRoIHead.simple_test_bboxes fusion (prob_roi_head.py:232-240),
 *   ProbConvFCBBoxHead.get_bboxes (convfc_bbox_head.py:294-330) and
 *   multiclass_nms (mmdet/core/post_processing/bbox_nms.py:8-95).
 * RoIs are in the padded layout of brcnn_rpn_get_bboxes: image b owns rows
 * [b*rois_per_img, b*rois_per_img + num_rois[b]).
 * ---------------------------------------------------------------------- */
typedef struct brcnn_rcnn_params {
  int32_t batch;
  int32_t rois_per_img;        /* row capacity per image                     */
  int32_t num_classes;         /* C                                          */
  int32_t reg_class_agnostic;
  int32_t prob;                /* ProbRoIHead.prob: sqrt(softmax*prior)      */
  int32_t rescale;             /* divide boxes by scale_factor               */
  float means[4];
  float stds[4];
  float max_ratio;
  float score_thr;             /* test_cfg.rcnn.score_thr                    */
  float iou_threshold;         /* test_cfg.rcnn.nms.iou_threshold            */
  int32_t max_per_img;         /* test_cfg.rcnn.max_per_img                  */
} brcnn_rcnn_params;

typedef struct brcnn_rcnn_ws_layout {
  int64_t scores;      /* float  [B*Rc][C+1]  fused scores                   */
  int64_t bboxes;      /* float  [B*Rc][C][4] decoded (+rescaled) boxes      */
  int64_t img_maxc;    /* float  [B]                                         */
  int64_t seg_count;   /* int32  [B][C]  candidates above score_thr          */
  int64_t seg_key;     /* uint64 [B][C][Rc] sorted score_bits<<32|~flat_idx  */
  int64_t kept_pos;    /* int32  [B][C][Rc]                                  */
  int64_t kept_count;  /* int32  [B][C]                                      */
  int64_t total_bytes;
} brcnn_rcnn_ws_layout;

int brcnn_rcnn_workspace_layout(const brcnn_rcnn_params* p, brcnn_rcnn_ws_layout* out);
size_t brcnn_rcnn_workspace_bytes(const brcnn_rcnn_params* p);

int brcnn_rcnn_get_bboxes(
    const brcnn_rcnn_params* p,
    const float* rois,          /* (B*Rc,5) [b,x1,y1,x2,y2]                   */
    const float* prior,         /* (B*Rc) proposal score (proposals[:, -1])   */
    const int32_t* num_rois,    /* (B)                                        */
    const float* cls_score,     /* (B*Rc, C+1) logits                         */
    const float* bbox_pred,     /* (B*Rc, 4C | 4)                             */
    const float* img_hw,        /* (B,2)                                      */
    const float* scale_factor,  /* (B,4)                                      */
    float* det_bboxes,          /* (B,max_per_img,5)                          */
    int64_t* det_labels,        /* (B,max_per_img)                            */
    int32_t* num_dets,          /* (B)                                        */
    void* workspace, size_t workspace_bytes, brcnn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* BRCNN_H_ */
