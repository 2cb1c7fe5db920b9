"""Torch-facing operators over libbrcnn (C ABI, hand-written sm_100a kernels).

Mirrors the slice of ``mmcv.ops`` the reference path uses
(``batched_nms``, ``nms``, ``RoIAlign`` / ``roi_align`` — call sites
mmdet/models/dense_heads/atss_rpn_head.py:6,756, mmdet/core/post_processing/
bbox_nms.py:3,86, mmdet/models/roi_heads/roi_extractors/base_roi_extractor.py:
54-59) plus the fused batch-level entry points the drop-in heads call.

torch is plumbing here: it owns device memory and the current stream; every
computation is a libbrcnn kernel.  CPU tensors are rejected (no fallback).
"""
import math
from ctypes import c_int32

import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function
from torch.nn.modules.utils import _pair

from . import _lib
from ._lib import (AssignParams, LossParams, RcnnParams, RcnnWsLayout, RoiParams, RpnLossParams,
                   RpnParams, RpnWsLayout, SampleParams, check, ptr_array)


def _stream():
    return torch.cuda.current_stream().cuda_stream


#: reduce_mean of the RPN loss normalisers over the ranks (atss_rpn_head.py:441,459).  A caller
#: that evaluates the loss on ONE rank only (a profiler, a per-stage timer) must switch it off:
#: a collective the other ranks do not issue desynchronises the process group.
RPN_LOSS_REDUCE_MEAN = True

_SCHED_SLOT_INTS = 64          # brcnn_roi_extract_forward_workspace_bytes / 4
_SCHED_POOLS = {}              # device index -> [zeroed int32 pool, next free slot, {stream: slot}]


def _sched_scratch(device, nbytes):
    """Zero-filled scheduling scratch of the persistent RoIAlign forward.  The kernel leaves
    it zero-filled, so a slot is cleared exactly once, when its pool is allocated.  Eager
    calls share one slot per stream (calls on a stream are ordered); every call recorded
    into a CUDA graph takes a slot of its own, because graphs captured on one stream may
    later replay concurrently on several."""
    assert nbytes <= 4 * _SCHED_SLOT_INTS
    capturing = torch.cuda.is_current_stream_capturing()
    pool = _SCHED_POOLS.get(device.index)
    if pool is None and not capturing:
        pool = [torch.zeros((1024, _SCHED_SLOT_INTS), dtype=torch.int32, device=device), 0, {}]
        _SCHED_POOLS[device.index] = pool
    if pool is not None and pool[1] < pool[0].size(0):
        if capturing:
            slot = pool[1]
            pool[1] += 1
            return pool[0][slot]
        key = _stream()
        if key not in pool[2]:
            pool[2][key] = pool[1]
            pool[1] += 1
        return pool[0][pool[2][key]]
    if not capturing and _stream() in pool[2]:
        return pool[0][pool[2][_stream()]]
    # pool exhausted / first use inside a capture: a fresh cleared buffer (inside a capture the
    # clear is recorded into the graph: one extra fill per replay)
    return torch.zeros((_SCHED_SLOT_INTS,), dtype=torch.int32, device=device)


def _first_cuda_tensor(args):
    for a in args:
        if isinstance(a, torch.Tensor):
            if a.is_cuda:
                return a
        elif isinstance(a, (list, tuple)):
            t = _first_cuda_tensor(a)
            if t is not None:
                return t
    return None


def _check_same_device(xs, device, name):
    for a in xs:
        if isinstance(a, torch.Tensor):
            if a.is_cuda and a.device != device:
                raise RuntimeError(f'{name}: tensors on {a.device} and {device}')
        elif isinstance(a, (list, tuple)):
            _check_same_device(a, device, name)


def _device_guard(fn):
    """Run ``fn`` with the device of its first CUDA tensor argument current (like the mmcv ops'
    device guard): launches, cudaFuncSetAttribute and ``_stream()`` then refer to the GPU the
    tensors live on even when the caller never called ``torch.cuda.set_device``; all CUDA
    tensor arguments must share that device.  (No closures over the arguments: a recursive
    local function would form a reference cycle that keeps the first tensor — and the
    autograd graph behind it — alive until the cyclic GC runs.)"""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        t = _first_cuda_tensor(args)
        if t is None:
            t = _first_cuda_tensor(tuple(kwargs.values()))
        if t is None:
            return fn(*args, **kwargs)
        dev = t.device
        del t
        _check_same_device(args, dev, fn.__name__)
        _check_same_device(tuple(kwargs.values()), dev, fn.__name__)
        if dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapper


def host_to_device(data, dtype, device):
    """Small host list / CPU tensor -> device through PINNED staging memory.  A copy from
    pageable memory (``torch.tensor(x).to(dev)``, ``torch.tensor(x, device=dev)``) makes the
    host wait until the stream has drained up to the copy — a hidden synchronisation that
    stops the host from queueing work ahead of the GPU.  The caching host allocator keeps the
    staging block alive until the copy has executed."""
    if torch.device(device).type != 'cuda':
        return torch.as_tensor(data, dtype=dtype).to(device)
    if isinstance(data, torch.Tensor):
        t = data.to(dtype).pin_memory()
    else:
        t = torch.tensor(data, dtype=dtype, pin_memory=True)
    return t.to(device, non_blocking=True)


def _f32c(t, name):
    if not t.is_cuda:
        raise RuntimeError(
            f'{name} must be a CUDA tensor: boosting_rcnn_b200 has no CPU path')
    if t.dtype != torch.float32:
        t = t.float()
    t = t.contiguous()
    if t.data_ptr() % 16:
        # a "contiguous" view can still start mid-row (e.g. rois[:1, 1:]); the
        # kernels read boxes / feature quads with 128-bit loads
        t = t.clone()
    return t


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def max_ratio_f32(wh_ratio_clip=16 / 1000):
    """``np.abs(np.log(wh_ratio_clip))`` as torch.clamp applies it to fp32
    (delta_xywh_bbox_coder.py:226,234-235)."""
    return float(np.float32(np.abs(np.log(wh_ratio_clip))))


# --------------------------------------------------------------------------
# RPN proposals
# --------------------------------------------------------------------------
def make_rpn_params(batch, featmap_sizes, strides, num_anchors, nms_pre,
                    max_per_img, iou_threshold, min_bbox_size,
                    means=(0., 0., 0., 0.), stds=(1., 1., 1., 1.),
                    wh_ratio_clip=16 / 1000):
    p = RpnParams()
    p.batch, p.num_levels, p.num_anchors = batch, len(featmap_sizes), num_anchors
    for l, ((h, w), s) in enumerate(zip(featmap_sizes, strides)):
        sw, sh = _pair(s)
        p.feat_h[l], p.feat_w[l] = int(h), int(w)
        p.stride_w[l], p.stride_h[l] = int(sw), int(sh)
    p.nms_pre, p.max_per_img = int(nms_pre), int(max_per_img)
    p.iou_threshold, p.min_bbox_size = float(iou_threshold), float(min_bbox_size)
    for i in range(4):
        p.means[i], p.stds[i] = float(means[i]), float(stds[i])
    p.max_ratio = max_ratio_f32(wh_ratio_clip)
    return p


def rpn_workspace_layout(p):
    lay = RpnWsLayout()
    check(_lib.load().brcnn_rpn_workspace_layout(p, lay), 'brcnn_rpn_workspace_layout')
    return lay


@_device_guard
def rpn_get_bboxes(p, cls_scores, bbox_preds, iou_preds, base_anchors, img_hw,
                   return_workspace=False):
    """Batched ATSSRPNHead.get_bboxes.  Returns (proposals (B,M,5) zero padded,
    num_proposals (B,) int32)."""
    lib = _lib.load()
    cls_scores = [_f32c(t, 'cls_scores') for t in cls_scores]
    bbox_preds = [_f32c(t, 'bbox_preds') for t in bbox_preds]
    iou_preds = [_f32c(t, 'iou_preds') for t in iou_preds]
    base_anchors = _f32c(base_anchors, 'base_anchors')
    img_hw = _f32c(img_hw, 'img_hw')
    dev = cls_scores[0].device
    B, L, A = p.batch, p.num_levels, p.num_anchors
    for l in range(L):
        assert tuple(cls_scores[l].shape) == (B, A, p.feat_h[l], p.feat_w[l]), cls_scores[l].shape
        assert tuple(bbox_preds[l].shape) == (B, 4 * A, p.feat_h[l], p.feat_w[l])
        assert tuple(iou_preds[l].shape) == (B, A, p.feat_h[l], p.feat_w[l])
    assert tuple(base_anchors.shape) == (L, A, 4)
    assert tuple(img_hw.shape) == (B, 2)
    nbytes = lib.brcnn_rpn_workspace_bytes(p)
    if nbytes == 0:
        raise RuntimeError('brcnn_rpn_workspace_bytes: bad parameters')
    ws = _ws(nbytes, dev)
    proposals = torch.empty((B, p.max_per_img, 5), dtype=torch.float32, device=dev)
    num = torch.empty((B,), dtype=torch.int32, device=dev)
    rc = lib.brcnn_rpn_get_bboxes(
        p, ptr_array([t.data_ptr() for t in cls_scores]),
        ptr_array([t.data_ptr() for t in bbox_preds]),
        ptr_array([t.data_ptr() for t in iou_preds]),
        base_anchors.data_ptr(), img_hw.data_ptr(), proposals.data_ptr(),
        num.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
    check(rc, 'brcnn_rpn_get_bboxes')
    if return_workspace:
        return proposals, num, ws
    return proposals, num


@_device_guard
def delta2bbox(rois, deltas, means=(0., 0., 0., 0.), stds=(1., 1., 1., 1.),
               max_shape=None, wh_ratio_clip=16 / 1000, clip_border=True):
    """mmdet delta2bbox for (N,4) rois and (N,4*k) deltas (one image)."""
    import ctypes
    rois, deltas = _f32c(rois, 'rois'), _f32c(deltas, 'deltas')
    n = rois.size(0)
    assert rois.dim() == 2 and rois.size(1) == 4 and deltas.size(0) == n
    assert deltas.size(1) % 4 == 0
    out = torch.empty_like(deltas)
    if n == 0:
        return out
    mh, mw = -1.0, -1.0
    if clip_border and max_shape is not None:
        mh, mw = float(max_shape[0]), float(max_shape[1])
    m = (ctypes.c_float * 4)(*[float(v) for v in means])
    s = (ctypes.c_float * 4)(*[float(v) for v in stds])
    check(_lib.load().brcnn_delta2bbox(rois.data_ptr(), deltas.data_ptr(), n,
                                       deltas.size(1) // 4, m, s,
                                       max_ratio_f32(wh_ratio_clip), mh, mw,
                                       out.data_ptr(), _stream()), 'brcnn_delta2bbox')
    return out


# --------------------------------------------------------------------------
# RPN loss path (anchor targets + focal / IoU / MSE / BCE losses + gradients)
# --------------------------------------------------------------------------
def make_rpn_loss_params(batch, featmap_sizes, strides, num_anchors, max_gts, pos_iou_thr=0.5,
                         neg_iou_thr=0.5, min_pos_iou=0.0, gamma=0.5, focal_gamma=2.0,
                         focal_alpha=0.25, loss_cls_weight=1.0, loss_bbox_weight=1.0,
                         loss_iou_weight=1.0, loss_aug_weight=1.0, wh_ratio_clip=16 / 1000,
                         cls_loss='focal'):
    p = RpnLossParams()
    p.batch, p.num_levels, p.num_anchors = int(batch), len(featmap_sizes), int(num_anchors)
    for l, ((h, w), s) in enumerate(zip(featmap_sizes, strides)):
        sw, sh = _pair(s)
        p.feat_h[l], p.feat_w[l] = int(h), int(w)
        p.stride_w[l], p.stride_h[l] = int(sw), int(sh)
    p.max_gts = int(max_gts)
    p.pos_iou_thr, p.neg_iou_thr, p.min_pos_iou = float(pos_iou_thr), float(neg_iou_thr), float(min_pos_iou)
    p.gamma, p.focal_gamma, p.focal_alpha = float(gamma), float(focal_gamma), float(focal_alpha)
    p.cls_loss_type = {'focal': 0, 'varifocal': 1}[cls_loss]
    p.loss_cls_weight, p.loss_bbox_weight = float(loss_cls_weight), float(loss_bbox_weight)
    p.loss_iou_weight, p.loss_aug_weight = float(loss_iou_weight), float(loss_aug_weight)
    p.max_ratio = max_ratio_f32(wh_ratio_clip)
    return p


class _RpnLossFunction(Function):
    """Returns 3L scalar losses (L x cls, L x bbox, L x iou) + the raw sums tensor.

    One fused all-reduce of the two normalisers replaces the reference's two
    ``reduce_mean(...).item()`` round trips (atss_rpn_head.py:441-444, 458-460); the
    division happens on the device, the step stays free of host syncs."""

    @staticmethod
    def forward(ctx, params, base_anchors, gt_boxes, num_gt, pad_hw, *heads):
        lib = _lib.load()
        L = params.num_levels
        assert len(heads) == 3 * L
        cls = [_f32c(t, 'cls_scores') for t in heads[:L]]
        box = [_f32c(t, 'bbox_preds') for t in heads[L:2 * L]]
        iou = [_f32c(t, 'iou_preds') for t in heads[2 * L:]]
        dev = cls[0].device
        B, A = params.batch, params.num_anchors
        for l in range(L):
            hw = (params.feat_h[l], params.feat_w[l])
            assert tuple(cls[l].shape) == (B, A) + hw and tuple(box[l].shape) == (B, 4 * A) + hw
            assert tuple(iou[l].shape) == (B, A) + hw
        base_anchors = _f32c(base_anchors, 'base_anchors')
        gt_boxes, pad_hw = _f32c(gt_boxes, 'gt_boxes'), _f32c(pad_hw, 'pad_hw')
        num_gt = num_gt.to(torch.int32).contiguous()
        sums = torch.empty((3 * L + 2,), dtype=torch.float32, device=dev)
        raw = [torch.empty_like(t) for t in cls + box + iou]
        ws = _ws(lib.brcnn_rpn_loss_workspace_bytes(params), dev)
        check(lib.brcnn_rpn_loss_forward(
            params, ptr_array([t.data_ptr() for t in cls]), ptr_array([t.data_ptr() for t in box]),
            ptr_array([t.data_ptr() for t in iou]), base_anchors.data_ptr(), gt_boxes.data_ptr(),
            num_gt.data_ptr(), pad_hw.data_ptr(), sums.data_ptr(),
            ptr_array([t.data_ptr() for t in raw[:L]]), ptr_array([t.data_ptr() for t in raw[L:2 * L]]),
            ptr_array([t.data_ptr() for t in raw[2 * L:]]), ws.data_ptr(), ws.numel(), _stream()),
            'brcnn_rpn_loss_forward')
        # normalisers: reduce_mean over the ranks, clamp at 1 (one fused all-reduce, on device)
        norm = sums[3 * L:].clone()
        import torch.distributed as dist
        if RPN_LOSS_REDUCE_MEAN and dist.is_available() and dist.is_initialized() \
                and dist.get_world_size() > 1:
            dist.all_reduce(norm.div_(dist.get_world_size()), op=dist.ReduceOp.SUM)
        norm = norm.clamp_(min=1.0)                  # [num_total_samples, bbox_avg_factor]
        inv = torch.cat([norm[0:1].expand(L), norm[1:2].expand(L), norm[0:1].expand(L)]).reciprocal()
        losses = sums[:3 * L] * inv
        ctx.save_for_backward(inv, *raw)
        ctx.params = params
        ctx.mark_non_differentiable(sums)
        return (*losses.unbind(0), sums)

    @staticmethod
    def backward(ctx, *grads):
        lib = _lib.load()
        inv, *raw = ctx.saved_tensors
        p = ctx.params
        L = p.num_levels
        up = torch.stack([g if g is not None else inv.new_zeros(()) for g in grads[:3 * L]])
        scale = (up.to(inv) * inv).contiguous()
        out = [torch.empty_like(t) for t in raw]
        check(lib.brcnn_rpn_loss_scale(
            p, ptr_array([t.data_ptr() for t in raw[:L]]), ptr_array([t.data_ptr() for t in raw[L:2 * L]]),
            ptr_array([t.data_ptr() for t in raw[2 * L:]]), scale.data_ptr(),
            ptr_array([t.data_ptr() for t in out[:L]]), ptr_array([t.data_ptr() for t in out[L:2 * L]]),
            ptr_array([t.data_ptr() for t in out[2 * L:]]), _stream()), 'brcnn_rpn_loss_scale')
        return (None, None, None, None, None, *out)


@_device_guard
def rpn_loss(params, cls_scores, bbox_preds, iou_preds, base_anchors, gt_boxes, num_gt, pad_hw):
    """Fused ATSSRPNHead.loss (atss=False).  gt_boxes (B,max_gts,4) zero padded + num_gt (B,)
    int32, pad_hw (B,2) = img_meta['pad_shape'][:2].  Returns (loss_cls [L], loss_bbox [L],
    loss_iou [L], sums (3L+2,)): lists of 0-dim tensors carrying gradients to the head outputs."""
    L = params.num_levels
    out = _RpnLossFunction.apply(params, base_anchors, gt_boxes, num_gt, pad_hw, *cls_scores,
                                 *bbox_preds, *iou_preds)
    return list(out[:L]), list(out[L:2 * L]), list(out[2 * L:3 * L]), out[3 * L]


# --------------------------------------------------------------------------
# mmcv.ops.nms / batched_nms mirrors
# --------------------------------------------------------------------------
@_device_guard
def _nms_raw(boxes, scores, idxs, iou_threshold, offset, num_ids=None, max_num=-1):
    """-> (dets (K,5), keep (K,), num (1,) int32): padded outputs + device count, no host sync.
    ``num_ids``: caller's bound on the id range (ids in [0, num_ids)); None = unknown (the
    library then treats the boxes as one segment of offset boxes, like mmcv).  ``max_num``: only
    the first max_num keeps are wanted (the sweeps stop there)."""
    lib = _lib.load()
    boxes = _f32c(boxes, 'boxes')
    scores = _f32c(scores, 'scores')
    K = boxes.size(0)
    dev = boxes.device
    keep = torch.empty((K,), dtype=torch.int64, device=dev)
    dets = torch.empty((K, 5), dtype=torch.float32, device=dev)
    num = torch.zeros((1,), dtype=torch.int32, device=dev)
    if K == 0:
        return dets, keep, num
    nid = 1
    if idxs is not None:
        idxs = idxs.to(torch.int64).contiguous()
        nid = int(num_ids) if num_ids is not None else 0
    ws = _ws(lib.brcnn_nms_workspace_bytes(K), dev)
    rc = lib.brcnn_batched_nms(
        boxes.data_ptr(), scores.data_ptr(),
        idxs.data_ptr() if idxs is not None else None, K, nid, float(iou_threshold),
        int(offset), int(max_num), keep.data_ptr(), dets.data_ptr(), num.data_ptr(),
        ws.data_ptr(), ws.numel(), _stream())
    check(rc, 'brcnn_batched_nms')
    return dets, keep, num


def _trim(dets, keep, num):
    """the reference API returns variable-length tensors: the one host read of the operator"""
    n = int(num.item())
    return dets[:n], keep[:n]


def nms(boxes, scores, iou_threshold, offset=0, score_threshold=0, max_num=-1):
    """mmcv.ops.nms: returns (dets (k,5), inds (k,) int64), score descending."""
    assert boxes.size(1) == 4 and boxes.size(0) == scores.size(0)
    assert offset in (0, 1)
    valid_inds = None
    if score_threshold > 0:
        valid_mask = scores > score_threshold
        valid_inds = torch.nonzero(valid_mask, as_tuple=False).squeeze(dim=1)
        boxes, scores = boxes[valid_mask], scores[valid_mask]
    dets, inds = _trim(*_nms_raw(boxes, scores, None, iou_threshold, offset, max_num=max_num))
    if max_num > 0:
        dets, inds = dets[:max_num], inds[:max_num]
    if valid_inds is not None:
        inds = valid_inds[inds]
    return dets, inds


def batched_nms(boxes, scores, idxs, nms_cfg, class_agnostic=False):
    """mmcv.ops.batched_nms.  ``split_thr`` is accepted and
    ignored: with the pinned (score desc, index asc) order the split and
    unsplit paths return identical results.  Extra key ``num_ids`` (bound on the id range,
    e.g. the number of classes / pyramid levels) selects the per-id kernels without any host
    read of ``idxs``; without it the bound is read from ``idxs.max()`` (one host sync, like
    the data-dependent output size)."""
    nms_cfg_ = dict(nms_cfg)
    class_agnostic = nms_cfg_.pop('class_agnostic', class_agnostic)
    nms_type = nms_cfg_.pop('type', 'nms')
    if nms_type != 'nms':
        raise NotImplementedError(f'nms type {nms_type!r} is outside the hot path')
    nms_cfg_.pop('split_thr', None)
    max_num = nms_cfg_.pop('max_num', -1)
    iou_threshold = nms_cfg_.pop('iou_threshold')
    num_ids = nms_cfg_.pop('num_ids', None)
    if not class_agnostic and num_ids is None and boxes.size(0) > 0:
        num_ids = int(idxs.max().item()) + 1 if int(idxs.min().item()) >= 0 else 0
    dets, keep = _trim(*_nms_raw(boxes, scores, None if class_agnostic else idxs,
                                 iou_threshold, 0, num_ids, max_num))
    if max_num > 0:
        dets, keep = dets[:max_num], keep[:max_num]
    return dets, keep


# --------------------------------------------------------------------------
# RoI extraction (level map + multi-level RoIAlign), autograd enabled
# --------------------------------------------------------------------------
def make_roi_params(batch, channels, featmap_sizes, spatial_scales, output_size,
                    sampling_ratio=0, aligned=True, finest_scale=56, out_layout=0):
    p = RoiParams()
    p.batch, p.channels, p.num_levels = int(batch), int(channels), len(featmap_sizes)
    for l, ((h, w), s) in enumerate(zip(featmap_sizes, spatial_scales)):
        p.feat_h[l], p.feat_w[l], p.spatial_scale[l] = int(h), int(w), float(s)
    oh, ow = _pair(output_size)
    p.pooled_h, p.pooled_w = int(oh), int(ow)
    p.sampling_ratio, p.aligned = int(sampling_ratio), int(bool(aligned))
    p.finest_scale = float(finest_scale)
    p.out_layout = int(out_layout)
    return p


@_device_guard
def to_nhwc(x):
    """(B,C,H,W) tensor -> NHWC-contiguous storage (returned as a (B,H,W,C)
    tensor).  channels_last inputs are a free view; NCHW-contiguous ones go
    through the library's transpose kernel."""
    assert x.dim() == 4
    if not x.is_cuda:
        raise RuntimeError('to_nhwc: CUDA tensor required (no CPU path)')
    if x.dtype != torch.float32:
        x = x.float()
    B, C, H, W = x.shape
    xp = x.permute(0, 2, 3, 1)
    if xp.is_contiguous():
        return xp
    x = x.contiguous()
    out = torch.empty((B, H, W, C), dtype=torch.float32, device=x.device)
    check(_lib.load().brcnn_nchw_to_nhwc(x.data_ptr(), out.data_ptr(), B, C, H * W,
                                         _stream()), 'brcnn_nchw_to_nhwc')
    return out


@_device_guard
def nhwc_to_nchw(x):
    """(B,H,W,C) contiguous -> (B,C,H,W) contiguous via the library kernel."""
    B, H, W, C = x.shape
    x = x.contiguous()
    out = torch.empty((B, C, H, W), dtype=torch.float32, device=x.device)
    check(_lib.load().brcnn_nhwc_to_nchw(x.data_ptr(), out.data_ptr(), B, C, H * W,
                                         _stream()), 'brcnn_nhwc_to_nchw')
    return out


@_device_guard
def _multi_transpose(fn_name, srcs, dst_shapes):
    """One launch for a whole pyramid: srcs are contiguous fp32 CUDA tensors
    sharing batch and channel counts."""
    lib = _lib.load()
    outs = [torch.empty(sh, dtype=torch.float32, device=srcs[0].device) for sh in dst_shapes]
    B = srcs[0].shape[0]
    if fn_name == 'brcnn_nchw_to_nhwc_multi':
        C = srcs[0].shape[1]
        hw = [s.shape[2] * s.shape[3] for s in srcs]
    else:
        C = srcs[0].shape[3]
        hw = [s.shape[1] * s.shape[2] for s in srcs]
    hw_arr = (c_int32 * len(hw))(*hw)
    check(getattr(lib, fn_name)(ptr_array([s.data_ptr() for s in srcs]),
                                ptr_array([o.data_ptr() for o in outs]), len(srcs), B, C,
                                hw_arr, _stream()), fn_name)
    return outs


def pyramid_to_nhwc(feats):
    """list of (B,C,H,W) maps -> list of NHWC-contiguous (B,H,W,C) tensors.
    channels_last maps are free views; all NCHW-contiguous ones are converted
    by ONE multi-map transpose launch."""
    out = [None] * len(feats)
    todo = []
    for i, f in enumerate(feats):
        assert f.dim() == 4
        if not f.is_cuda:
            raise RuntimeError('pyramid_to_nhwc: CUDA tensors required (no CPU path)')
        if f.dtype != torch.float32:
            f = f.float()
        fp = f.permute(0, 2, 3, 1)
        if fp.is_contiguous():
            out[i] = fp
        else:
            todo.append((i, f.contiguous()))
    same = len({(f.shape[0], f.shape[1]) for _, f in todo}) == 1
    if todo and same and len(todo) <= _lib.MAX_LEVELS:
        srcs = [f for _, f in todo]
        res = _multi_transpose('brcnn_nchw_to_nhwc_multi', srcs,
                               [(f.shape[0], f.shape[2], f.shape[3], f.shape[1]) for f in srcs])
        for (i, _), r in zip(todo, res):
            out[i] = r
    else:
        for i, f in todo:
            out[i] = to_nhwc(f)
    return out


def pyramid_to_nchw(grads_nhwc):
    """list of NHWC-contiguous (B,H,W,C) -> NCHW-contiguous (B,C,H,W), one launch."""
    if not grads_nhwc:
        return []
    return _multi_transpose('brcnn_nhwc_to_nchw_multi', grads_nhwc,
                            [(g.shape[0], g.shape[3], g.shape[1], g.shape[2]) for g in grads_nhwc])


@_device_guard
def bbox2roi_padded(proposals, num):
    """Padded proposals (B,cap,5) + num (B) int32 -> rois (B*cap,5) with
    b = -1 padding rows, prior (B*cap).  One kernel (bbox2roi, transforms.py:59-78)."""
    proposals = _f32c(proposals, 'proposals')
    assert proposals.dim() == 3 and proposals.size(2) == 5
    assert num.dtype == torch.int32 and num.is_cuda
    B, cap = proposals.shape[:2]
    rois = torch.empty((B * cap, 5), dtype=torch.float32, device=proposals.device)
    prior = torch.empty((B * cap,), dtype=torch.float32, device=proposals.device)
    check(_lib.load().brcnn_bbox2roi_padded(proposals.data_ptr(), num.contiguous().data_ptr(),
                                            B, cap, rois.data_ptr(), prior.data_ptr(),
                                            _stream()), 'brcnn_bbox2roi_padded')
    return rois, prior


@_device_guard
def map_roi_levels(rois, num_levels, finest_scale=56):
    rois = _f32c(rois, 'rois')
    out = torch.empty((rois.size(0),), dtype=torch.int64, device=rois.device)
    check(_lib.load().brcnn_map_roi_levels(rois.data_ptr(), rois.size(0),
                                           float(finest_scale), int(num_levels),
                                           out.data_ptr(), _stream()),
          'brcnn_map_roi_levels')
    return out


class _RoiExtractFunction(Function):
    """feats are (B,C,H,W)-shaped tensors (any memory format).

    ``channels_last_out``: the (R,C,oh,ow) result is stored (R,oh,ow,C) — a
    ``torch.channels_last`` tensor.  That is the RoI-feature hand-off to the first FC
    (``ProbConvFCBBoxHead`` reads it as a free (R, oh*ow*C) view against permuted weight
    columns); the forward kernel then stores straight from registers and the backward
    kernel reads ``grad_out`` bin-major with no transpose."""

    @staticmethod
    def forward(ctx, rois, spatial_scales, output_size, sampling_ratio, aligned,
                finest_scale, channels_last_out, *feats):
        lib = _lib.load()
        rois = _f32c(rois, 'rois')
        assert rois.dim() == 2 and rois.size(1) == 5, 'RoI must be (idx, x1, y1, x2, y2)!'
        B, C = feats[0].shape[:2]
        sizes = [tuple(f.shape[-2:]) for f in feats]
        p = make_roi_params(B, C, sizes, spatial_scales, output_size,
                            sampling_ratio, aligned, finest_scale,
                            out_layout=1 if channels_last_out else 0)
        nhwc = pyramid_to_nhwc(feats)
        R = rois.size(0)
        if channels_last_out:
            out = torch.empty((R, p.pooled_h, p.pooled_w, C), dtype=torch.float32,
                              device=rois.device).permute(0, 3, 1, 2)
        else:
            out = torch.empty((R, C, p.pooled_h, p.pooled_w), dtype=torch.float32,
                              device=rois.device)
        lvls = torch.empty((R,), dtype=torch.int32, device=rois.device)
        ws_bytes = lib.brcnn_roi_extract_forward_workspace_bytes(p)
        ws = _sched_scratch(rois.device, ws_bytes)
        check(lib.brcnn_roi_extract_forward(
            p, ptr_array([t.data_ptr() for t in nhwc]), rois.data_ptr(), R,
            out.data_ptr(), lvls.data_ptr(), ws.data_ptr(), ws_bytes, _stream()),
            'brcnn_roi_extract_forward')
        ctx.save_for_backward(rois)
        ctx.params = p
        ctx.channels_last = [f.permute(0, 2, 3, 1).is_contiguous() for f in feats]
        ctx.mark_non_differentiable(lvls)
        return out, lvls

    @staticmethod
    def backward(ctx, grad_out, _grad_lvls):
        (rois,) = ctx.saved_tensors
        grads = roi_extract_backward(ctx.params, grad_out, rois)
        # channels_last inputs get a channels_last gradient for free;
        # NCHW-contiguous inputs get an NCHW-contiguous one (one launch).
        need = [i for i, cl in enumerate(ctx.channels_last) if not cl]
        conv = dict(zip(need, pyramid_to_nchw([grads[i] for i in need])))
        outs = [conv[i] if i in conv else g.permute(0, 3, 1, 2) for i, g in enumerate(grads)]
        return (None, None, None, None, None, None, None, *outs)


@_device_guard
def roi_extract_backward(params, grad_out, rois):
    """Gradient of ``roi_extract`` w.r.t. every pyramid level, NHWC (B,H,W,C) each.
    ``grad_out`` is the logical (R,C,oh,ow) tensor; bin-major storage ((R,oh,ow,C), what the
    permuted-FC hand-off produces) is consumed as it is, anything else goes through the
    (R,C,oh*ow) layout and one transpose inside the library."""
    lib = _lib.load()
    p = RoiParams.from_buffer_copy(params)
    if not grad_out.is_cuda:
        raise RuntimeError('grad_out must be a CUDA tensor: no CPU path')
    if grad_out.dtype != torch.float32:
        grad_out = grad_out.float()
    gp = grad_out.permute(0, 2, 3, 1)
    if (gp.is_contiguous() and grad_out.data_ptr() % 16 == 0 and p.pooled_h <= 7
            and p.pooled_w <= 7):
        p.out_layout = 1
    else:
        p.out_layout = 0
        grad_out = _f32c(grad_out, 'grad_out')
    dev = grad_out.device
    R = rois.size(0)
    grads = [torch.empty((p.batch, p.feat_h[l], p.feat_w[l], p.channels),
                         dtype=torch.float32, device=dev)
             for l in range(p.num_levels)]
    ws = _ws(lib.brcnn_roi_extract_backward_workspace_bytes(p, R), dev)
    check(lib.brcnn_roi_extract_backward(
        p, grad_out.data_ptr(), rois.data_ptr(), R,
        ptr_array([g.data_ptr() for g in grads]), ws.data_ptr(), ws.numel(),
        _stream()), 'brcnn_roi_extract_backward')
    return grads


@_device_guard
def roi_extract(feats, rois, spatial_scales, output_size=7, sampling_ratio=0,
                aligned=True, finest_scale=56, return_levels=False, channels_last_out=False):
    """Fused SingleRoIExtractor.forward: level mapping + RoIAlign on every
    level in one launch.  Returns (R,C,oh,ow) [and int32 levels]; with
    ``channels_last_out`` the same logical tensor stored (R,oh,ow,C)."""
    oh, ow = _pair(output_size)
    if channels_last_out and (oh > 7 or ow > 7):
        channels_last_out = False       # the persistent kernel covers pooled sizes <= 7
    out, lvls = _RoiExtractFunction.apply(rois, tuple(spatial_scales), output_size,
                                          sampling_ratio, aligned, finest_scale,
                                          bool(channels_last_out), *feats)
    return (out, lvls) if return_levels else out


def roi_align(input, rois, output_size, spatial_scale=1.0, sampling_ratio=0,
              pool_mode='avg', aligned=True):
    """mmcv.ops.roi_align on one feature map."""
    if pool_mode != 'avg':
        raise NotImplementedError("pool_mode='max' is outside the hot path")
    return roi_extract([input], rois, [spatial_scale], output_size, sampling_ratio,
                       aligned, finest_scale=56)


class RoIAlign(nn.Module):
    """mmcv.ops.RoIAlign(output_size, spatial_scale, sampling_ratio, pool_mode,
    aligned, use_torchvision) with the same attributes (``output_size`` is read
    by single_level_roi_extractor.py:60)."""

    def __init__(self, output_size, spatial_scale=1.0, sampling_ratio=0,
                 pool_mode='avg', aligned=True, use_torchvision=False):
        super().__init__()
        self.output_size = _pair(output_size)
        self.spatial_scale = float(spatial_scale)
        self.sampling_ratio = int(sampling_ratio)
        self.pool_mode = pool_mode
        self.aligned = aligned
        self.use_torchvision = use_torchvision

    def forward(self, input, rois):
        return roi_align(input, rois, self.output_size, self.spatial_scale,
                         self.sampling_ratio, self.pool_mode, self.aligned)

    def __repr__(self):
        return (f'{self.__class__.__name__}(output_size={self.output_size}, '
                f'spatial_scale={self.spatial_scale}, '
                f'sampling_ratio={self.sampling_ratio}, pool_mode={self.pool_mode}, '
                f'aligned={self.aligned}, use_torchvision={self.use_torchvision})')


# --------------------------------------------------------------------------
# R-CNN training front-end: assign + sample + targets + prior (two launches)
# --------------------------------------------------------------------------
def sample_plan(counts, num, pos_fraction, neg_pos_ub=-1):
    """Host half of RandomSampler.sample (samplers/base_sampler.py:35-102,
    random_sampler.py:32-82) for a batch: from the per-image numbers of positive /
    negative candidates decide how many rows each image contributes and draw the
    permutations exactly like the reference — ``torch.randperm(n)[:k]`` on the CPU
    generator, only when a list is longer than its quota, positives before negatives,
    image by image.  Pure host code (unit-tested without a GPU).

    Returns plan (B,5) int32 [n_pos, n_neg, first_row, use_perm_pos, use_perm_neg],
    perm_pos / perm_neg (B, max(1,num)) int32, rows per image."""
    B = len(counts)
    n_exp_pos = int(num * pos_fraction)
    cap = max(1, num)
    perms = torch.zeros((2, B, cap), dtype=torch.int32)
    plan_rows, rows, base = [], [], 0
    for b in range(B):
        n_pos_c, n_neg_c = int(counts[b][0]), int(counts[b][1])
        if n_pos_c > n_exp_pos:
            perms[0, b, :n_exp_pos] = torch.randperm(n_pos_c)[:n_exp_pos]
            n_pos, use_p = n_exp_pos, 1
        else:
            n_pos, use_p = n_pos_c, 0
        n_exp_neg = num - n_pos
        if neg_pos_ub >= 0:
            n_exp_neg = min(n_exp_neg, int(neg_pos_ub * max(1, n_pos)))
        if n_neg_c > n_exp_neg:
            perms[1, b, :n_exp_neg] = torch.randperm(n_neg_c)[:n_exp_neg]
            n_neg, use_n = n_exp_neg, 1
        else:
            n_neg, use_n = n_neg_c, 0
        plan_rows.append([n_pos, n_neg, base, use_p, use_n])
        rows.append(n_pos + n_neg)
        base += n_pos + n_neg
    return torch.tensor(plan_rows, dtype=torch.int32).reshape(B, 5), perms[0], perms[1], rows


class RcnnAssigned:
    """Result of ``rcnn_assign``: the device-side state ``rcnn_sample_targets`` consumes plus
    the per-image (positive, negative) candidate counts on their way to the host.  The counts
    are copied into pinned memory behind an event, so a caller can keep queueing unrelated
    GPU work (the RPN loss) between ``rcnn_assign`` and the first ``counts()`` call — the
    reference's CPU ``randperm`` then runs while the GPU is busy instead of behind a drained
    stream."""

    def __init__(self, proposals, num_props, gtb, gtl, num_gt, gt_inds, counts_dev, num_gts_host,
                 is_static=False):
        self.proposals, self.num_props = proposals, num_props
        self.gtb, self.gtl, self.num_gt, self.gt_inds = gtb, gtl, num_gt, gt_inds
        self.num_gts_host = list(num_gts_host)
        self.is_static = is_static      # the tensors are a caller-owned, reused buffer set
        self._counts_host = torch.empty(counts_dev.shape, dtype=counts_dev.dtype, pin_memory=True)
        self._counts_host.copy_(counts_dev, non_blocking=True)
        self._event = torch.cuda.Event()
        self._event.record(torch.cuda.current_stream(counts_dev.device))
        self._counts_dev = counts_dev          # keeps the source alive until the copy is done

    def counts(self):
        """[[n_pos_candidates, n_neg_candidates]] per image: the one host sync of the
        training front-end (an event wait, not a stream drain)."""
        self._event.synchronize()
        return self._counts_host.tolist()


def rcnn_assign_buffers(batch, max_props, max_gts, device):
    """A reusable buffer set for ``rcnn_assign(..., static=...)``: the inputs of the captured
    training step live at fixed addresses, so nothing has to be copied between the assignment
    and the graph replay."""
    z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=device)
    return dict(proposals=z((batch, max_props, 5), torch.float32),
                num_props=z((batch,), torch.int32), gtb=z((batch, max_gts, 4), torch.float32),
                gtl=z((batch, max_gts), torch.int64), num_gt=z((batch,), torch.int32),
                gt_inds=z((batch, max_gts + max_props), torch.int32))


@_device_guard
def rcnn_assign(proposals, num_props, gt_bboxes, gt_labels, pos_iou_thr, neg_iou_thr,
                min_pos_iou=0., max_gts=None, static=None):
    """MaxIoUAssigner(match_low_quality=False) over ``cat[gt_bboxes, proposals]`` for a whole
    batch (max_iou_assigner.py:61-212 + assign_result.py:191-205 ``add_gt_``): one launch.
    ``max_gts`` pads the GT tensors to a fixed capacity (static shapes for a captured step);
    ``static``: a ``rcnn_assign_buffers`` set of exactly these shapes to work in (the proposals
    are copied into it; GT rows beyond an image's count keep stale values nobody reads)."""
    lib = _lib.load()
    proposals = _f32c(proposals, 'proposals')
    assert proposals.dim() == 3 and proposals.size(2) == 5
    dev = proposals.device
    B, M = proposals.shape[:2]
    num_props = num_props.to(torch.int32).contiguous()
    Gs = [int(g.size(0)) for g in gt_bboxes]
    Gmax = max(1, max(Gs))
    if max_gts is not None:
        assert max_gts >= Gmax
        Gmax = int(max_gts)
    if static is not None:
        assert tuple(static['proposals'].shape) == (B, M, 5) and static['gtb'].size(1) == Gmax
        static['proposals'].copy_(proposals)
        static['num_props'].copy_(num_props)
        proposals, num_props = static['proposals'], static['num_props']
        gtb, gtl, gt_inds = static['gtb'], static['gtl'], static['gt_inds']
        num_gt = static['num_gt']
        num_gt.copy_(torch.tensor(Gs, dtype=torch.int32, pin_memory=True), non_blocking=True)
    else:
        gtb = torch.zeros((B, Gmax, 4), dtype=torch.float32, device=dev)
        gtl = torch.zeros((B, Gmax), dtype=torch.int64, device=dev)
        num_gt = host_to_device(Gs, torch.int32, dev)
        gt_inds = torch.empty((B, Gmax + M), dtype=torch.int32, device=dev)
    for b in range(B):
        if Gs[b]:
            gtb[b, :Gs[b]] = gt_bboxes[b].float()
            gtl[b, :Gs[b]] = gt_labels[b].long()
    counts = torch.empty((B, 2), dtype=torch.int32, device=dev)
    ap = AssignParams(B, M, Gmax, float(pos_iou_thr), float(neg_iou_thr), float(min_pos_iou), 0)
    check(lib.brcnn_rcnn_assign(ap, proposals.data_ptr(), num_props.data_ptr(), gtb.data_ptr(),
                                num_gt.data_ptr(), gt_inds.data_ptr(), counts.data_ptr(),
                                _stream()), 'brcnn_rcnn_assign')
    return RcnnAssigned(proposals, num_props, gtb, gtl, num_gt, gt_inds, counts, Gs,
                        is_static=static is not None)


@_device_guard
def rcnn_sample_targets(proposals, num_props, gtb, gtl, num_gt, gt_inds, plan, perm_pos,
                        perm_neg, N, num_classes, means=(0., 0., 0., 0.),
                        stds=(1., 1., 1., 1.), pos_weight=-1):
    """RandomSampler row selection + BBoxHead.get_targets + the ProbRoIHead prior vector for a
    whole batch from device-resident inputs only (one launch, no host sync: capturable).
    plan (B,5) / perm_pos / perm_neg (B,cap) int32 come from ``sample_plan``; N = total rows.
    Returns rois (N,5), labels (N,), label_weights (N,), bbox_targets (N,4),
    bbox_weights (N,4), prior (N,)."""
    lib = _lib.load()
    dev = proposals.device
    B, M = proposals.shape[:2]
    Gmax = gtb.size(1)
    cap = perm_pos.size(1)
    rois = torch.empty((N, 5), dtype=torch.float32, device=dev)
    labels = torch.empty((N,), dtype=torch.int64, device=dev)
    label_weights = torch.empty((N,), dtype=torch.float32, device=dev)
    bbox_targets = torch.empty((N, 4), dtype=torch.float32, device=dev)
    bbox_weights = torch.empty((N, 4), dtype=torch.float32, device=dev)
    prior = torch.empty((N,), dtype=torch.float32, device=dev)
    if N > 0:
        sp = SampleParams()
        sp.batch, sp.max_props, sp.max_gts, sp.num_classes = B, M, Gmax, int(num_classes)
        sp.perm_cap, sp.max_sel = cap, cap
        for i in range(4):
            sp.means[i], sp.stds[i] = float(means[i]), float(stds[i])
        sp.pos_weight = 1.0 if pos_weight <= 0 else float(pos_weight)
        check(lib.brcnn_rcnn_sample_targets(
            sp, proposals.data_ptr(), num_props.data_ptr(), gtb.data_ptr(), gtl.data_ptr(),
            num_gt.data_ptr(), gt_inds.data_ptr(), plan.data_ptr(), perm_pos.data_ptr(),
            perm_neg.data_ptr(), rois.data_ptr(), labels.data_ptr(), label_weights.data_ptr(),
            bbox_targets.data_ptr(), bbox_weights.data_ptr(), prior.data_ptr(), _stream()),
            'brcnn_rcnn_sample_targets')
    return rois, labels, label_weights, bbox_targets, bbox_weights, prior


def rcnn_assign_sample(proposals, num_props, gt_bboxes, gt_labels, num_classes,
                       pos_iou_thr, neg_iou_thr, min_pos_iou=0., num=512, pos_fraction=0.25,
                       neg_pos_ub=-1, means=(0., 0., 0., 0.), stds=(1., 1., 1., 1.),
                       pos_weight=-1, assigned=None):
    """MaxIoUAssigner(match_low_quality=False) + RandomSampler(add_gt_as_proposals=True) +
    BBoxHead.get_targets + the ProbRoIHead prior vector for a whole batch.

    proposals (B,M,5) padded + num_props (B,) int32 (the layout of rpn_get_bboxes);
    gt_bboxes / gt_labels: per-image lists of (G_b,4) / (G_b,) device tensors.
    The permutations are drawn with torch.randperm on the CPU generator in the
    reference's order (random_sampler.py:58): per image positives, then negatives.
    ``assigned``: the result of an earlier ``rcnn_assign`` on the same inputs.
    Returns rois (N,5), labels (N,), label_weights (N,), bbox_targets (N,4),
    bbox_weights (N,4), prior (N,), rows_per_image (list)."""
    a = assigned if assigned is not None else rcnn_assign(
        proposals, num_props, gt_bboxes, gt_labels, pos_iou_thr, neg_iou_thr, min_pos_iou)
    dev = a.proposals.device
    cnt = a.counts()                     # the one host sync of the training front-end
    plan, perm_pos, perm_neg, rows = sample_plan(cnt, num, pos_fraction, neg_pos_ub)
    to = lambda t: host_to_device(t, torch.int32, dev)
    out = rcnn_sample_targets(a.proposals, a.num_props, a.gtb, a.gtl, a.num_gt, a.gt_inds,
                              to(plan), to(perm_pos), to(perm_neg), sum(rows), num_classes,
                              means, stds, pos_weight)
    return (*out, rows)


# --------------------------------------------------------------------------
# Boosting reweighted loss
# --------------------------------------------------------------------------
class _BoostLossFunction(Function):

    @staticmethod
    def forward(ctx, cls_score, bbox_pred, labels, label_weights, prior,
                bbox_targets, bbox_weights, num_classes, reg_class_agnostic, gamma,
                alpha, loss_cls_weight, loss_bbox_weight, reg_norm_mean):
        lib = _lib.load()
        cls_score = _f32c(cls_score, 'cls_score')
        bbox_pred = _f32c(bbox_pred, 'bbox_pred')
        prior = _f32c(prior, 'prior')
        bbox_targets = _f32c(bbox_targets, 'bbox_targets')
        bbox_weights = _f32c(bbox_weights, 'bbox_weights')
        labels = labels.to(torch.int64).contiguous()
        if label_weights is not None:
            label_weights = _f32c(label_weights, 'label_weights')
        N = cls_score.size(0)
        assert cls_score.size(1) == num_classes + 1
        p = LossParams(N, num_classes, int(reg_class_agnostic), float(gamma),
                       float(alpha), float(loss_cls_weight), float(loss_bbox_weight),
                       int(reg_norm_mean))
        out = torch.empty((8,), dtype=torch.float32, device=cls_score.device)
        g_cls = torch.empty_like(cls_score)
        g_box = torch.empty_like(bbox_pred)
        ws = _ws(lib.brcnn_boost_loss_workspace_bytes(p), cls_score.device)
        check(lib.brcnn_boost_loss(
            p, cls_score.data_ptr(), labels.data_ptr(),
            label_weights.data_ptr() if label_weights is not None else None,
            prior.data_ptr(), bbox_pred.data_ptr(), bbox_targets.data_ptr(),
            bbox_weights.data_ptr(), out.data_ptr(), g_cls.data_ptr(),
            g_box.data_ptr(), ws.data_ptr(), ws.numel(), _stream()), 'brcnn_boost_loss')
        ctx.save_for_backward(g_cls, g_box)
        loss_cls, loss_bbox, acc = out[0], out[1], out[2]
        ctx.mark_non_differentiable(acc)
        return loss_cls, loss_bbox, acc, out

    @staticmethod
    def backward(ctx, g_loss_cls, g_loss_bbox, _g_acc, _g_out):
        g_cls, g_box = ctx.saved_tensors
        gc = g_cls * g_loss_cls if g_loss_cls is not None else None
        gb = g_box * g_loss_bbox if g_loss_bbox is not None else None
        return (gc, gb) + (None,) * 12


@_device_guard
def boost_loss(cls_score, bbox_pred, labels, label_weights, prior, bbox_targets,
               bbox_weights, num_classes, reg_class_agnostic=False, gamma=0.5,
               alpha=0.0, loss_cls_weight=1.0, loss_bbox_weight=1.0,
               reg_norm_mean=False):
    """Fused ProbRoIHead boost loss.  Returns (loss_cls, loss_bbox, acc,
    scalars[8]); loss_cls / loss_bbox carry gradients to cls_score / bbox_pred."""
    return _BoostLossFunction.apply(
        cls_score, bbox_pred, labels, label_weights, prior, bbox_targets,
        bbox_weights, num_classes, reg_class_agnostic, gamma, alpha,
        loss_cls_weight, loss_bbox_weight, reg_norm_mean)


# --------------------------------------------------------------------------
# Score fusion + decode + class-wise NMS
# --------------------------------------------------------------------------
def make_rcnn_params(batch, rois_per_img, num_classes, score_thr, iou_threshold,
                     max_per_img, means=(0., 0., 0., 0.), stds=(.1, .1, .2, .2),
                     reg_class_agnostic=False, prob=True, rescale=False,
                     wh_ratio_clip=16 / 1000):
    p = RcnnParams()
    p.batch, p.rois_per_img, p.num_classes = int(batch), int(rois_per_img), int(num_classes)
    p.reg_class_agnostic, p.prob, p.rescale = int(reg_class_agnostic), int(prob), int(rescale)
    for i in range(4):
        p.means[i], p.stds[i] = float(means[i]), float(stds[i])
    p.max_ratio = max_ratio_f32(wh_ratio_clip)
    p.score_thr, p.iou_threshold = float(score_thr), float(iou_threshold)
    p.max_per_img = int(max_per_img)
    return p


def rcnn_workspace_layout(p):
    lay = RcnnWsLayout()
    check(_lib.load().brcnn_rcnn_workspace_layout(p, lay), 'brcnn_rcnn_workspace_layout')
    return lay


@_device_guard
def rcnn_get_bboxes(p, rois, prior, num_rois, cls_score, bbox_pred, img_hw,
                    scale_factor=None, return_workspace=False):
    """Batched fusion + ProbConvFCBBoxHead.get_bboxes + multiclass_nms on the
    padded RoI layout.  Returns (det_bboxes (B,M,5), det_labels (B,M) int64,
    num_dets (B,) int32)."""
    lib = _lib.load()
    rois = _f32c(rois, 'rois')
    cls_score = _f32c(cls_score, 'cls_score')
    bbox_pred = _f32c(bbox_pred, 'bbox_pred')
    img_hw = _f32c(img_hw, 'img_hw')
    dev = rois.device
    rows = p.batch * p.rois_per_img
    assert rois.shape == (rows, 5) and cls_score.shape == (rows, p.num_classes + 1)
    assert bbox_pred.shape == (rows, 4 if p.reg_class_agnostic else 4 * p.num_classes)
    if prior is not None:
        prior = _f32c(prior, 'prior')
        assert prior.numel() == rows
    if scale_factor is not None:
        scale_factor = _f32c(scale_factor, 'scale_factor')
        assert scale_factor.shape == (p.batch, 4)
    num_rois = num_rois.to(torch.int32).contiguous()
    nbytes = lib.brcnn_rcnn_workspace_bytes(p)
    if nbytes == 0:
        raise RuntimeError('brcnn_rcnn_workspace_bytes: bad parameters')
    ws = _ws(nbytes, dev)
    det = torch.empty((p.batch, p.max_per_img, 5), dtype=torch.float32, device=dev)
    lab = torch.empty((p.batch, p.max_per_img), dtype=torch.int64, device=dev)
    num = torch.empty((p.batch,), dtype=torch.int32, device=dev)
    check(lib.brcnn_rcnn_get_bboxes(
        p, rois.data_ptr(), prior.data_ptr() if prior is not None else None,
        num_rois.data_ptr(), cls_score.data_ptr(), bbox_pred.data_ptr(),
        img_hw.data_ptr(),
        scale_factor.data_ptr() if scale_factor is not None else None,
        det.data_ptr(), lab.data_ptr(), num.data_ptr(), ws.data_ptr(), ws.numel(),
        _stream()), 'brcnn_rcnn_get_bboxes')
    if return_workspace:
        return det, lab, num, ws
    return det, lab, num
