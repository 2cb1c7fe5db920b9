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
