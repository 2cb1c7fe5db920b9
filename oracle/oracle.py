"""numpy-facing wrapper of oracle/brcnn_oracle.c.

TEST INFRASTRUCTURE ONLY (see the C file's header): imported by tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs, never by
boosting_rcnn_b200/.
"""
import ctypes
import os
import subprocess
from ctypes import POINTER, c_float, c_int32, c_int64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, 'brcnn_oracle.c')
SO = os.path.join(_HERE, '_oracle.so')
_lib = None


def build():
    cmd = ['gcc', '-O2', '-ffp-contract=off', '-fno-fast-math', '-fPIC', '-shared',
           '-o', SO, SRC, '-lm']
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('gcc failed:\n' + res.stdout + res.stderr)
    return SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
            build()
        _lib = ctypes.CDLL(SO)
        _lib.oracle_expf.restype = c_float
        _lib.oracle_expf.argtypes = [c_float]
        _lib.oracle_sigmoid.restype = c_float
        _lib.oracle_sigmoid.argtypes = [c_float]
        for n in ('oracle_nms_cpu', 'oracle_batched_nms', 'oracle_rpn_get_bboxes_single',
                  'oracle_rcnn_get_bboxes_single'):
            getattr(_lib, n).restype = c_int32
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(c_void_p) if a is not None else None


def _ptrs(arrs):
    out = (c_void_p * len(arrs))()
    for i, a in enumerate(arrs):
        out[i] = a.ctypes.data
    return out


def _i32(seq):
    return np.ascontiguousarray(np.asarray(seq, dtype=np.int32))


def max_ratio_f32(wh_ratio_clip=16 / 1000):
    return float(np.float32(np.abs(np.log(wh_ratio_clip))))


def expf(x):
    x = _f(x)
    out = np.empty_like(x)
    lib().oracle_expf_array(_p(x), c_int64(x.size), _p(out))
    return out


def sigmoid(x):
    x = _f(x)
    return (np.float32(1.0) / (np.float32(1.0) + expf(-x))).astype(np.float32)


def delta2bbox(rois, deltas, means=(0., 0., 0., 0.), stds=(1., 1., 1., 1.),
               max_shape=None, wh_ratio_clip=16 / 1000):
    rois, deltas = _f(rois), _f(deltas)
    n = rois.shape[0]
    ncls = deltas.shape[1] // 4 if n else 1
    out = np.empty_like(deltas)
    mh, mw = (-1.0, -1.0) if max_shape is None else (float(max_shape[0]), float(max_shape[1]))
    lib().oracle_delta2bbox(_p(rois), _p(deltas), c_int32(n), c_int32(ncls),
                            _p(_f(means)), _p(_f(stds)), c_float(max_ratio_f32(wh_ratio_clip)),
                            c_float(mh), c_float(mw), _p(out))
    return out


def nms_cpu(boxes, scores, iou_threshold, offset=0):
    boxes, scores = _f(boxes), _f(scores)
    n = boxes.shape[0]
    keep = np.empty((max(n, 1),), dtype=np.int64)
    k = lib().oracle_nms_cpu(_p(boxes), _p(scores), c_int32(n), c_float(iou_threshold),
                             c_int32(offset), _p(keep))
    return keep[:k].copy()


def batched_nms(boxes, scores, idxs, iou_threshold, split_thr=10000):
    boxes, scores = _f(boxes), _f(scores)
    n = boxes.shape[0]
    if idxs is not None:
        idxs = np.ascontiguousarray(idxs, dtype=np.int64)
    keep = np.empty((max(n, 1),), dtype=np.int64)
    dets = np.empty((max(n, 1), 5), dtype=np.float32)
    k = lib().oracle_batched_nms(_p(boxes), _p(scores), _p(idxs), c_int32(n),
                                 c_float(iou_threshold), c_int32(split_thr), _p(keep),
                                 _p(dets))
    return dets[:k].copy(), keep[:k].copy()


def rpn_get_bboxes_single(cls, bbox, iou, base_anchors, strides, img_shape, nms_pre,
                          max_per_img, iou_threshold, min_bbox_size=0,
                          means=(0., 0., 0., 0.), stds=(1., 1., 1., 1.),
                          split_thr=10000, debug=False):
    """cls[l]: (A,H,W), bbox[l]: (4A,H,W), iou[l]: (A,H,W) for ONE image."""
    L = len(cls)
    cls = [_f(c) for c in cls]
    bbox = [_f(c) for c in bbox]
    iou = [_f(c) for c in iou]
    A = cls[0].shape[0]
    fh = _i32([c.shape[1] for c in cls])
    fw = _i32([c.shape[2] for c in cls])
    sw = _i32([s[0] if isinstance(s, (tuple, list)) else s for s in strides])
    sh = _i32([s[1] if isinstance(s, (tuple, list)) else s for s in strides])
    base_anchors = _f(base_anchors)
    props = np.zeros((max_per_img, 5), dtype=np.float32)
    ns = [int(fh[l]) * int(fw[l]) * A for l in range(L)]
    K = sum(min(n, nms_pre) if nms_pre > 0 else n for n in ns)
    topk = np.full((K,), -1, dtype=np.int32)
    cboxes = np.zeros((K, 4), dtype=np.float32)
    cscores = np.zeros((K,), dtype=np.float32)
    cn = np.zeros((L,), dtype=np.int32)
    n = lib().oracle_rpn_get_bboxes_single(
        c_int32(L), c_int32(A), _p(fh), _p(fw), _p(sw), _p(sh), _ptrs(cls), _ptrs(bbox),
        _ptrs(iou), _p(base_anchors), c_float(img_shape[0]), c_float(img_shape[1]),
        c_int32(nms_pre), c_int32(max_per_img), c_float(iou_threshold),
        c_float(min_bbox_size), _p(_f(means)), _p(_f(stds)), c_float(max_ratio_f32()),
        c_int32(split_thr), _p(props), _p(topk), _p(cboxes), _p(cscores), _p(cn))
    if debug:
        return props[:n].copy(), dict(topk_idx=topk, cand_boxes=cboxes,
                                      cand_scores=cscores, cand_n=cn)
    return props[:n].copy()


def map_roi_levels(rois, num_levels, finest_scale=56):
    rois = _f(rois)
    out = np.empty((rois.shape[0],), dtype=np.int64)
    lib().oracle_map_roi_levels(_p(rois), c_int32(rois.shape[0]), c_float(finest_scale),
                                c_int32(num_levels), _p(out))
    return out


def roi_align_forward(inp, rois, output_size, spatial_scale, sampling_ratio=0,
                      aligned=True):
    inp, rois = _f(inp), _f(rois)
    N, C, H, W = inp.shape
    ph, pw = (output_size, output_size) if isinstance(output_size, int) else output_size
    R = rois.shape[0]
    out = np.empty((R, C, ph, pw), dtype=np.float32)
    lib().oracle_roi_align_forward(_p(inp), c_int32(N), c_int32(C), c_int32(H), c_int32(W),
                                   _p(rois), c_int32(R), c_int32(ph), c_int32(pw),
                                   c_float(spatial_scale), c_int32(sampling_ratio),
                                   c_int32(int(aligned)), _p(out))
    return out


def roi_align_backward(grad_out, rois, input_shape, spatial_scale, sampling_ratio=0,
                       aligned=True):
    grad_out, rois = _f(grad_out), _f(rois)
    N, C, H, W = input_shape
    R, _, ph, pw = grad_out.shape
    gi = np.zeros((N, C, H, W), dtype=np.float32)
    lib().oracle_roi_align_backward(_p(grad_out), c_int32(N), c_int32(C), c_int32(H),
                                    c_int32(W), _p(rois), c_int32(R), c_int32(ph),
                                    c_int32(pw), c_float(spatial_scale),
                                    c_int32(sampling_ratio), c_int32(int(aligned)), _p(gi))
    return gi


def roi_extract_forward(feats, rois, spatial_scales, output_size=7, sampling_ratio=0,
                        aligned=True, finest_scale=56):
    """feats: list of (B,C,H,W) NCHW arrays.  Returns (out (R,C,ph,pw), lvls)."""
    feats = [_f(f) for f in feats]
    rois = _f(rois)
    L = len(feats)
    B, C = feats[0].shape[:2]
    ph, pw = (output_size, output_size) if isinstance(output_size, int) else output_size
    R = rois.shape[0]
    out = np.empty((R, C, ph, pw), dtype=np.float32)
    lv = np.empty((R,), dtype=np.int64)
    lib().oracle_roi_extract_forward(
        c_int32(L), c_int32(B), c_int32(C), _p(_i32([f.shape[2] for f in feats])),
        _p(_i32([f.shape[3] for f in feats])), _p(_f(spatial_scales)), _ptrs(feats),
        _p(rois), c_int32(R), c_int32(ph), c_int32(pw), c_int32(sampling_ratio),
        c_int32(int(aligned)), c_float(finest_scale), _p(out), _p(lv))
    return out, lv


def roi_extract_backward(grad_out, rois, feat_shapes, spatial_scales, sampling_ratio=0,
                         aligned=True, finest_scale=56):
    grad_out, rois = _f(grad_out), _f(rois)
    L = len(feat_shapes)
    B, C = feat_shapes[0][:2]
    R, _, ph, pw = grad_out.shape
    grads = [np.empty(s, dtype=np.float32) for s in feat_shapes]
    lib().oracle_roi_extract_backward(
        c_int32(L), c_int32(B), c_int32(C), _p(_i32([s[2] for s in feat_shapes])),
        _p(_i32([s[3] for s in feat_shapes])), _p(_f(spatial_scales)), _p(grad_out),
        _p(rois), c_int32(R), c_int32(ph), c_int32(pw), c_int32(sampling_ratio),
        c_int32(int(aligned)), c_float(finest_scale), _ptrs(grads))
    return grads


def fuse_scores(cls_score, prior, prob=True):
    cls_score = _f(cls_score)
    R, C1 = cls_score.shape
    prior = _f(prior) if prior is not None else np.ones((R,), np.float32)
    out = np.empty_like(cls_score)
    lib().oracle_fuse_scores(_p(cls_score), _p(prior), c_int32(R), c_int32(C1),
                             c_int32(int(prob)), _p(out))
    return out


def rcnn_get_bboxes_single(rois, scores, bbox_pred, img_shape, scale_factor, num_classes,
                           score_thr, iou_threshold, max_per_img, rescale=False,
                           means=(0., 0., 0., 0.), stds=(.1, .1, .2, .2),
                           reg_class_agnostic=False, split_thr=10000, debug=False):
    rois, scores, bbox_pred = _f(rois), _f(scores), _f(bbox_pred)
    R = rois.shape[0]
    C = num_classes
    nbox = 1 if reg_class_agnostic else C
    cap = max_per_img if max_per_img > 0 else max(R * C, 1)
    det = np.zeros((cap, 5), dtype=np.float32)
    lab = np.zeros((cap,), dtype=np.int64)
    dec = np.zeros((max(R, 1), nbox * 4), dtype=np.float32)
    flat = np.zeros((cap,), dtype=np.int64)
    sf = _f(scale_factor if scale_factor is not None else [1, 1, 1, 1])
    n = lib().oracle_rcnn_get_bboxes_single(
        _p(rois), _p(scores), _p(bbox_pred), c_int32(R), c_int32(C),
        c_int32(int(reg_class_agnostic)), _p(_f(means)), _p(_f(stds)),
        c_float(max_ratio_f32()), c_float(img_shape[0]), c_float(img_shape[1]), _p(sf),
        c_int32(int(rescale)), c_float(score_thr), c_float(iou_threshold),
        c_int32(max_per_img), c_int32(split_thr), _p(det), _p(lab), _p(dec), _p(flat))
    if debug:
        return det[:n].copy(), lab[:n].copy(), dict(decoded=dec[:R], keep_flat=flat[:n].copy())
    return det[:n].copy(), lab[:n].copy()


def boost_loss(cls_score, labels, prior, bbox_pred, bbox_targets, bbox_weights,
               num_classes, label_weights=None, reg_class_agnostic=False, gamma=0.5,
               alpha=0.0, loss_cls_weight=1.0, loss_bbox_weight=1.0, reg_norm_mean=False):
    cls_score, prior, bbox_pred = _f(cls_score), _f(prior), _f(bbox_pred)
    bbox_targets, bbox_weights = _f(bbox_targets), _f(bbox_weights)
    labels = np.ascontiguousarray(labels, dtype=np.int64)
    lw = _f(label_weights) if label_weights is not None else None
    N = cls_score.shape[0]
    out = np.zeros((8,), dtype=np.float32)
    gc = np.zeros_like(cls_score)
    gb = np.zeros_like(bbox_pred)
    lib().oracle_boost_loss(_p(cls_score), _p(labels), _p(lw), _p(prior), _p(bbox_pred),
                            _p(bbox_targets), _p(bbox_weights), c_int32(N),
                            c_int32(num_classes), c_int32(int(reg_class_agnostic)),
                            c_float(gamma), c_float(alpha), c_float(loss_cls_weight),
                            c_float(loss_bbox_weight), c_int32(int(reg_norm_mean)),
                            _p(out), _p(gc), _p(gb))
    return dict(loss_cls=out[0], loss_bbox=out[1], acc=out[2], scalars=out,
                grad_cls=gc, grad_bbox=gb)


# ---------------------------------------------------------------------------
# R-CNN training front-end: assign + sample + targets + prior (numpy restatement;
# SURVEY.md §8 a11 / f1).  The random permutations are injected (`randperm(n)`), because the
# reference draws them with torch.randperm on the CPU generator (random_sampler.py:58).
# ---------------------------------------------------------------------------
def bbox_overlaps(b1, b2, eps=1e-6):
    """mmdet/core/bbox/iou_calculators/iou2d_calculator.py:75-260, mode='iou',
    is_aligned=False: (len(b1), len(b2)) fp32 IoU matrix, op for op."""
    b1, b2 = _f(b1).reshape(-1, 4), _f(b2).reshape(-1, 4)
    area1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    area2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    lt = np.maximum(b1[:, None, :2], b2[None, :, :2])
    rb = np.minimum(b1[:, None, 2:], b2[None, :, 2:])
    wh = np.maximum(rb - lt, np.float32(0))
    overlap = wh[..., 0] * wh[..., 1]
    union = area1[:, None] + area2[None, :] - overlap
    union = np.maximum(union, np.float32(eps))
    return (overlap / union).astype(np.float32)


def bbox_overlaps_aligned(b1, b2, eps=1e-6):
    """iou2d_calculator.py:214-226,250-253: is_aligned=True."""
    b1, b2 = _f(b1).reshape(-1, 4), _f(b2).reshape(-1, 4)
    area1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    area2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    lt = np.maximum(b1[:, :2], b2[:, :2])
    rb = np.minimum(b1[:, 2:], b2[:, 2:])
    wh = np.maximum(rb - lt, np.float32(0))
    overlap = wh[:, 0] * wh[:, 1]
    union = np.maximum(area1 + area2 - overlap, np.float32(eps))
    return (overlap / union).astype(np.float32)


def max_iou_assign(bboxes, gt_bboxes, pos_iou_thr, neg_iou_thr, min_pos_iou=0.0,
                   match_low_quality=True, gt_max_assign_all=True):
    """MaxIoUAssigner.assign / assign_wrt_overlaps
    (mmdet/core/bbox/assigners/max_iou_assigner.py:61-212), no ignore regions.
    Returns (gt_inds (n,) int64: 0 bg, -1 ignore, k+1 assigned GT; max_overlaps (n,))."""
    bboxes, gt_bboxes = _f(bboxes)[:, :4], _f(gt_bboxes).reshape(-1, 4)
    n, k = bboxes.shape[0], gt_bboxes.shape[0]
    gt_inds = np.full((n,), -1, dtype=np.int64)
    if k == 0 or n == 0:                                   # :148-160
        if k == 0:
            gt_inds[:] = 0
        return gt_inds, np.zeros((n,), np.float32)
    ov = bbox_overlaps(gt_bboxes, bboxes)                  # (k, n)
    max_ov, argmax_ov = ov.max(0), ov.argmax(0)            # first maximum
    gt_max = ov.max(1)
    if isinstance(neg_iou_thr, (tuple, list)):             # :173-181
        gt_inds[(max_ov >= np.float32(neg_iou_thr[0])) & (max_ov < np.float32(neg_iou_thr[1]))] = 0
    else:
        gt_inds[(max_ov >= 0) & (max_ov < np.float32(neg_iou_thr))] = 0
    pos = max_ov >= np.float32(pos_iou_thr)                # :184-185
    gt_inds[pos] = argmax_ov[pos] + 1
    if match_low_quality:                                  # :187-202, GT order matters
        for i in range(k):
            if gt_max[i] >= np.float32(min_pos_iou):
                if gt_max_assign_all:
                    gt_inds[ov[i] == gt_max[i]] = i + 1
                else:
                    gt_inds[ov[i].argmax()] = i + 1
    return gt_inds, max_ov


def bbox2delta(proposals, gt, means=(0., 0., 0., 0.), stds=(1., 1., 1., 1.)):
    """mmdet/core/bbox/coder/delta_xywh_bbox_coder.py:98-141, fp32."""
    p, g = _f(proposals).reshape(-1, 4), _f(gt).reshape(-1, 4)
    h = np.float32(0.5)
    px, py = (p[:, 0] + p[:, 2]) * h, (p[:, 1] + p[:, 3]) * h
    pw, ph = p[:, 2] - p[:, 0], p[:, 3] - p[:, 1]
    gx, gy = (g[:, 0] + g[:, 2]) * h, (g[:, 1] + g[:, 3]) * h
    gw, gh = g[:, 2] - g[:, 0], g[:, 3] - g[:, 1]
    with np.errstate(divide='ignore', invalid='ignore'):
        d = np.stack([(gx - px) / pw, (gy - py) / ph, np.log(gw / pw), np.log(gh / ph)], -1)
    d = (d - np.asarray(means, np.float32)[None]) / np.asarray(stds, np.float32)[None]
    return d.astype(np.float32)


def random_sample(gt_inds, num, pos_fraction, neg_pos_ub, randperm):
    """RandomSampler._sample_pos/_sample_neg + BaseSampler.sample index bookkeeping
    (samplers/random_sampler.py:32-82, base_sampler.py:78-101): gallery[randperm(n)[:k]]
    when a list is longer than its quota, then .unique() (= ascending sort)."""
    def pick(mask, k):
        inds = np.nonzero(mask)[0]
        if inds.size <= k:
            return inds
        perm = np.asarray(randperm(inds.size))[:k]
        return np.unique(inds[perm])
    pos = pick(gt_inds > 0, int(num * pos_fraction))
    n_neg = num - pos.size
    if neg_pos_ub >= 0:
        n_neg = min(n_neg, int(neg_pos_ub * max(1, pos.size)))
    neg = pick(gt_inds == 0, n_neg)
    return pos, neg


def rcnn_train_prep(proposals, gt_bboxes, gt_labels, num_classes, pos_iou_thr, neg_iou_thr,
                    min_pos_iou, num, pos_fraction, neg_pos_ub, means, stds, randperm,
                    pos_weight=-1, match_low_quality=False):
    """Body of ProbRoIHead.forward_train up to the head forward
    (mmdet/models/roi_heads/prob_roi_head.py:33-64) + bbox2roi (transforms.py:59-78) +
    BBoxHead.get_targets (bbox_head.py:122-253) for a batch given as per-image lists.
    Returns dict(rois, labels, label_weights, bbox_targets, bbox_weights, prior, rows,
    gt_inds=[per image, after add_gt_])."""
    out = dict(rois=[], labels=[], label_weights=[], bbox_targets=[], bbox_weights=[], prior=[],
               rows=[], gt_inds=[], pos_inds=[], neg_inds=[])
    for b, (pr, gb, gl) in enumerate(zip(proposals, gt_bboxes, gt_labels)):
        pr, gb = _f(pr), _f(gb).reshape(-1, 4)
        G = gb.shape[0]
        gi, _ = max_iou_assign(pr[:, :4], gb, pos_iou_thr, neg_iou_thr, min_pos_iou,
                               match_low_quality)
        boxes = pr[:, :4]
        if G > 0:                                          # add_gt_as_proposals (:78-88)
            boxes = np.concatenate([gb, boxes], 0)
            gi = np.concatenate([np.arange(1, G + 1, dtype=np.int64), gi])  # add_gt_
        pos, neg = random_sample(gi, num, pos_fraction, neg_pos_ub, randperm)
        pos_boxes, neg_boxes = boxes[pos], boxes[neg]
        n_pos, n_neg = pos.size, neg.size
        # prior extraction, prob_roi_head.py:51-64
        pos_prior = pr[pos[G:] - G, -1]
        neg_prior = np.float32(1) - pr[neg - G, -1]
        prior = np.concatenate([np.zeros((G,), np.float32), pos_prior, neg_prior]).astype(np.float32)
        # get_targets, bbox_head.py:122-186
        labels = np.full((n_pos + n_neg,), num_classes, dtype=np.int64)
        lw = np.zeros((n_pos + n_neg,), np.float32)
        bt = np.zeros((n_pos + n_neg, 4), np.float32)
        bw = np.zeros((n_pos + n_neg, 4), np.float32)
        if n_pos > 0:
            assigned = gi[pos] - 1
            labels[:n_pos] = np.asarray(gl, np.int64)[assigned]
            lw[:n_pos] = 1.0 if pos_weight <= 0 else pos_weight
            bt[:n_pos] = bbox2delta(pos_boxes, gb[assigned], means, stds)
            bw[:n_pos] = 1
        if n_neg > 0:
            lw[-n_neg:] = 1.0
        sb = np.concatenate([pos_boxes, neg_boxes], 0)
        out['rois'].append(np.concatenate([np.full((sb.shape[0], 1), b, np.float32), sb], 1))
        for k, v in (('labels', labels), ('label_weights', lw), ('bbox_targets', bt),
                     ('bbox_weights', bw), ('prior', prior)):
            out[k].append(v)
        out['rows'].append(n_pos + n_neg)
        out['gt_inds'].append(gi)
        out['pos_inds'].append(pos)
        out['neg_inds'].append(neg)
    for k in ('rois', 'labels', 'label_weights', 'bbox_targets', 'bbox_weights', 'prior'):
        out[k] = np.concatenate(out[k], 0)
    return out


# ---------------------------------------------------------------------------
# RPN loss path: anchor targets + focal / IoU / MSE / BCE losses and
# their gradients, numpy restatement of ATSSRPNHead.loss with atss=False
# (mmdet/models/dense_heads/atss_rpn_head.py:299-464,505-603; anchor_head.py:126-265).
# Assignment runs in fp32 exactly like the reference (IoU thresholds and the
# `overlaps == gt_max_overlaps` test of match_low_quality are bit-sensitive); the loss
# arithmetic runs in float64 (the parity bar on values / gradients is 1e-5 relative).
# ---------------------------------------------------------------------------
def grid_anchors_level(base, stride, H, W):
    """AnchorGenerator.single_level_grid_anchors (anchor_generator.py:338-381):
    anchor[(y*W+x)*A + a] = base[a] + [x*sw, y*sh, x*sw, y*sh], fp32."""
    sw, sh = (stride, stride) if np.isscalar(stride) else stride
    xs = (np.arange(W, dtype=np.float32) * np.float32(sw))
    ys = (np.arange(H, dtype=np.float32) * np.float32(sh))
    sx, sy = np.meshgrid(xs, ys)
    shifts = np.stack([sx.ravel(), sy.ravel(), sx.ravel(), sy.ravel()], 1).astype(np.float32)
    return (shifts[:, None, :] + _f(base)[None, :, :]).reshape(-1, 4).astype(np.float32)


def valid_flags_level(H, W, A, stride, pad_hw):
    """AnchorGenerator.valid_flags / single_level_valid_flags (anchor_generator.py:383-434)."""
    sw, sh = (stride, stride) if np.isscalar(stride) else stride
    vh = min(int(np.ceil(pad_hw[0] / sh)), H)
    vw = min(int(np.ceil(pad_hw[1] / sw)), W)
    vy = np.arange(H) < vh
    vx = np.arange(W) < vw
    return np.repeat((vy[:, None] & vx[None, :]).ravel(), A)


def rpn_anchor_targets(gt_bboxes, img_metas, base_anchors, strides, sizes, pos_iou_thr=0.5,
                       neg_iou_thr=0.5, min_pos_iou=0.0, match_low_quality=True):
    """get_anchors + get_targets/_get_targets_single with PseudoSampler, reg_decoded_bbox=True,
    num_classes=1, allowed_border=-1.  Returns per image: anchors (n,4), labels (n,) in
    {0 fg, 1 bg}, label_weights (n,), assigned gt index (n,) (-1 where not positive)."""
    A = base_anchors.shape[1]
    out = []
    for gts, meta in zip(gt_bboxes, img_metas):
        anchors = np.concatenate([grid_anchors_level(base_anchors[l], strides[l], H, W)
                                  for l, (H, W) in enumerate(sizes)], 0)
        valid = np.concatenate([valid_flags_level(H, W, A, strides[l], meta['pad_shape'][:2])
                                for l, (H, W) in enumerate(sizes)], 0)
        n = anchors.shape[0]
        labels = np.full((n,), 1, dtype=np.int64)          # unmap fill = num_classes (bg)
        lw = np.zeros((n,), np.float32)
        assigned = np.full((n,), -1, dtype=np.int64)
        if valid.any():
            gi, _ = max_iou_assign(anchors[valid], gts, pos_iou_thr, neg_iou_thr, min_pos_iou,
                                   match_low_quality)
            vi = np.nonzero(valid)[0]
            pos, neg = vi[gi > 0], vi[gi == 0]
            labels[pos] = 0
            lw[pos] = 1.0
            lw[neg] = 1.0
            assigned[pos] = gi[gi > 0] - 1
        out.append(dict(anchors=anchors, labels=labels, label_weights=lw, assigned=assigned))
    return out


def rpn_loss(cls, box, iou, gt_bboxes, img_metas, base_anchors, strides, gamma=0.5,
             focal_gamma=2.0, focal_alpha=0.25, w_cls=1.0, w_bbox=1.0, w_iou=1.0, w_aug=1.0,
             pos_iou_thr=0.5, neg_iou_thr=0.5, min_pos_iou=0.0, wh_ratio_clip=16 / 1000,
             world_size=1, other_ranks_num_pos=0.0, other_ranks_iou_sum=0.0, cls_loss='focal'):
    """ATSSRPNHead.loss / loss_single.  cls[l] (B,A,H,W), box[l] (B,4A,H,W), iou[l] (B,A,H,W).
    Returns dict: loss_rpn_cls / loss_rpn_bbox / loss_rpn_iou (L,) float32, grad_cls / grad_box /
    grad_iou (lists, gradients of the sum of all 3L losses), num_pos, iou_sum.
    `other_ranks_*` + `world_size` model reduce_mean (atss_rpn_head.py:441,459)."""
    L = len(cls)
    B, A = cls[0].shape[0], cls[0].shape[1]
    sizes = [tuple(c.shape[-2:]) for c in cls]
    tg = rpn_anchor_targets(gt_bboxes, img_metas, base_anchors, strides, sizes, pos_iou_thr,
                            neg_iou_thr, min_pos_iou, True)
    num_pos = float(sum(int((t['labels'] == 0).sum()) for t in tg))
    nts = max((num_pos + other_ranks_num_pos) / world_size, 1.0)
    M = float(np.float32(np.abs(np.log(wh_ratio_clip))))
    EPS, eps = 1e-12, 1e-6
    lvl_off = np.cumsum([0] + [h * w * A for h, w in sizes])
    res = dict(loss_rpn_cls=np.zeros(L), loss_rpn_bbox=np.zeros(L), loss_rpn_iou=np.zeros(L),
               grad_cls=[], grad_box=[], grad_iou=[])
    raw_bbox, iou_sum_l = [], []
    for l, (H, W) in enumerate(sizes):
        sl = slice(lvl_off[l], lvl_off[l + 1])
        x = np.stack([cls[l][b].transpose(1, 2, 0).reshape(-1) for b in range(B)]).astype(np.float64)
        d = np.stack([box[l][b].transpose(1, 2, 0).reshape(-1, 4) for b in range(B)]).astype(np.float64)
        u = np.stack([iou[l][b].transpose(1, 2, 0).reshape(-1) for b in range(B)]).astype(np.float64)
        lab = np.stack([t['labels'][sl] for t in tg])
        lw = np.stack([t['label_weights'][sl] for t in tg]).astype(np.float64)
        anc = np.stack([t['anchors'][sl] for t in tg]).astype(np.float64)
        # ---- classification: py_sigmoid_focal_loss (focal_loss.py:13-58); the varifocal
        # variant needs iou_target and is evaluated after the positives block ----
        t = (lab == 0).astype(np.float64)
        p = 1.0 / (1.0 + np.exp(-x))
        if cls_loss == 'focal':
            bce = np.maximum(x, 0) - x * t + np.log1p(np.exp(-np.abs(x)))
            pt = (1 - p) * t + p * (1 - t)
            aw = focal_alpha * t + (1 - focal_alpha) * (1 - t)
            fw = aw * pt ** focal_gamma
            res['loss_rpn_cls'][l] = w_cls * (bce * fw * lw).sum() / nts
            dfw = aw * focal_gamma * pt ** (focal_gamma - 1) * (1 - 2 * t) * p * (1 - p)
            g_cls = w_cls * lw * (fw * (p - t) + bce * dfw) / nts
        # ---- positives ----
        g_box = np.zeros_like(d)
        g_iou = np.zeros_like(u)
        pm = lab == 0
        lb_sum = 0.0
        isum = 0.0
        if pm.any():
            gt = np.stack([np.where(t_['assigned'][sl, None] >= 0,
                                    _f(g_).reshape(-1, 4)[np.maximum(t_['assigned'][sl], 0)]
                                    if len(g_) else np.zeros((sl.stop - sl.start, 4), np.float32),
                                    0.0) for t_, g_ in zip(tg, gt_bboxes)]).astype(np.float64)
            a_, d_, g_ = anc[pm], d[pm], gt[pm]
            px, py = (a_[:, 0] + a_[:, 2]) * 0.5, (a_[:, 1] + a_[:, 3]) * 0.5
            pw, ph = a_[:, 2] - a_[:, 0], a_[:, 3] - a_[:, 1]
            dwc, dhc = np.clip(d_[:, 2], -M, M), np.clip(d_[:, 3], -M, M)
            gw, gh = pw * np.exp(dwc), ph * np.exp(dhc)
            gx, gy = px + pw * d_[:, 0], py + ph * d_[:, 1]
            x1, y1, x2, y2 = gx - gw * 0.5, gy - gh * 0.5, gx + gw * 0.5, gy + gh * 0.5
            # aligned IoU (iou2d_calculator.py:214-253) and its derivative
            ltx, lty = np.maximum(x1, g_[:, 0]), np.maximum(y1, g_[:, 1])
            rbx, rby = np.minimum(x2, g_[:, 2]), np.minimum(y2, g_[:, 3])
            iwr, ihr = rbx - ltx, rby - lty
            iw, ih = np.maximum(iwr, 0), np.maximum(ihr, 0)
            inter = iw * ih
            ap = (x2 - x1) * (y2 - y1)
            ag = (g_[:, 2] - g_[:, 0]) * (g_[:, 3] - g_[:, 1])
            union = ap + ag - inter
            uc = np.maximum(union, eps)
            iou_t = inter / uc                                  # iou_target (detached) == loss IoU
            wgt = np.maximum(iou_t ** gamma, EPS)
            # encoded targets (bbox2delta) and the MSE "aug" loss
            tx, ty = (g_[:, 0] + g_[:, 2]) * 0.5, (g_[:, 1] + g_[:, 3]) * 0.5
            tw, th = g_[:, 2] - g_[:, 0], g_[:, 3] - g_[:, 1]
            with np.errstate(divide='ignore', invalid='ignore'):
                enc = np.stack([(tx - px) / pw, (ty - py) / ph, np.log(tw / pw), np.log(th / ph)], 1)
            diff = d_ - enc
            l_aug = w_aug * (wgt[:, None] * diff ** 2).sum()
            iou_c = np.maximum(iou_t, eps)
            l_iou = w_bbox * (wgt * -np.log(iou_c)).sum()
            lb_sum = 0.5 * (l_iou + l_aug)
            isum = iou_t.sum()
            # d(-log iou)/d box
            sel = lambda a, b: np.where(a > b, 1.0, np.where(a == b, 0.5, 0.0))
            dl_x1, dl_y1 = sel(x1, g_[:, 0]), sel(y1, g_[:, 1])           # d ltx/d x1 ...
            dr_x2, dr_y2 = sel(g_[:, 2], x2), sel(g_[:, 3], y2)           # d rbx/d x2 ...
            mw, mh = (iwr >= 0).astype(np.float64), (ihr >= 0).astype(np.float64)
            di = [-ih * mw * dl_x1, -iw * mh * dl_y1, ih * mw * dr_x2, iw * mh * dr_y2]
            da = [-(y2 - y1), -(x2 - x1), (y2 - y1), (x2 - x1)]
            um = sel(union, eps)
            coef = -(iou_t >= eps).astype(np.float64) / iou_c * wgt * w_bbox * 0.5
            gb = [coef * (di[k] / uc - inter * um * (da[k] - di[k]) / uc ** 2) for k in range(4)]
            cw_, ch_ = (np.abs(d_[:, 2]) <= M), (np.abs(d_[:, 3]) <= M)
            gd = np.stack([(gb[0] + gb[2]) * pw, (gb[1] + gb[3]) * ph,
                           (gb[2] - gb[0]) * 0.5 * gw * cw_, (gb[3] - gb[1]) * 0.5 * gh * ch_], 1)
            gd += 0.5 * w_aug * 2 * wgt[:, None] * diff
            g_box[pm] = gd
            # centerness/IoU branch: BCE with logits against iou_target
            xu = u[pm]
            res['loss_rpn_iou'][l] = w_iou * (np.maximum(xu, 0) - xu * iou_t +
                                              np.log1p(np.exp(-np.abs(xu)))).sum() / nts
            g_iou[pm] = w_iou * (1.0 / (1.0 + np.exp(-xu)) - iou_t) / nts
        if cls_loss == 'varifocal':
            # VarifocalLoss(iou_weighted=True) on EVERY anchor (no label weights are passed,
            # atss_rpn_head.py:393-397; varifocal_loss.py:45-57): target q = iou_target on
            # positives, 0 elsewhere
            q = np.zeros_like(x)
            if pm.any():
                q[pm] = iou_t
            posq = q > 0
            bce = np.maximum(x, 0) - x * q + np.log1p(np.exp(-np.abs(x)))
            dq = p - q
            fw = np.where(posq, q, focal_alpha * np.abs(dq) ** focal_gamma)
            res['loss_rpn_cls'][l] = w_cls * (bce * fw).sum() / nts
            dfw = np.where(posq, 0.0, focal_alpha * focal_gamma * np.abs(dq) ** (focal_gamma - 1)
                           * np.sign(dq) * p * (1 - p))
            g_cls = w_cls * (fw * (p - q) + bce * dfw) / nts
        raw_bbox.append((lb_sum, g_box))
        iou_sum_l.append(isum)
        res['grad_cls'].append(g_cls)
        res['grad_iou'].append(g_iou)
    iou_sum = float(sum(iou_sum_l))
    baf = max((iou_sum + other_ranks_iou_sum) / world_size, 1.0)
    for l, (H, W) in enumerate(sizes):
        lb, gbx = raw_bbox[l]
        res['loss_rpn_bbox'][l] = lb / baf
        gbx = gbx / baf
        to_nchw = lambda g, c: np.stack([g[b].reshape(H, W, A * c).transpose(2, 0, 1)
                                         for b in range(B)]).astype(np.float32)
        res['grad_box'].append(to_nchw(gbx, 4))
        res['grad_cls'][l] = to_nchw(res['grad_cls'][l], 1)
        res['grad_iou'][l] = to_nchw(res['grad_iou'][l], 1)
    for k in ('loss_rpn_cls', 'loss_rpn_bbox', 'loss_rpn_iou'):
        res[k] = res[k].astype(np.float32)
    res['num_pos'], res['iou_sum'], res['targets'] = num_pos, iou_sum, tg
    return res
