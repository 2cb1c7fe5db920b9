"""ctypes binding of ``libbrcnn.so`` (the C ABI declared in ``include/brcnn.h``).

The shared library is the product: every compute entry point of this package
goes through it.  There is no CPU / eager fallback — if the library is missing
or a call returns non-zero, a ``RuntimeError`` is raised.
"""
import ctypes
import os
import subprocess
import sys
from ctypes import (POINTER, Structure, c_char_p, c_float, c_int32, c_int64,
                    c_size_t, c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libbrcnn.so')
CSRC = os.path.join(_HERE, 'csrc')
MAX_LEVELS = 8

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo',
    '-fmad=false', '-std=c++17', '-shared', '-Xcompiler', '-fPIC'
]


class RpnParams(Structure):
    _fields_ = [
        ('batch', c_int32), ('num_levels', c_int32), ('num_anchors', c_int32),
        ('feat_h', c_int32 * MAX_LEVELS), ('feat_w', c_int32 * MAX_LEVELS),
        ('stride_w', c_int32 * MAX_LEVELS), ('stride_h', c_int32 * MAX_LEVELS),
        ('nms_pre', c_int32), ('max_per_img', c_int32),
        ('iou_threshold', c_float), ('min_bbox_size', c_float),
        ('means', c_float * 4), ('stds', c_float * 4), ('max_ratio', c_float),
    ]


class RpnLossParams(Structure):
    _fields_ = [
        ('batch', c_int32), ('num_levels', c_int32), ('num_anchors', c_int32),
        ('feat_h', c_int32 * MAX_LEVELS), ('feat_w', c_int32 * MAX_LEVELS),
        ('stride_w', c_int32 * MAX_LEVELS), ('stride_h', c_int32 * MAX_LEVELS),
        ('max_gts', c_int32),
        ('pos_iou_thr', c_float), ('neg_iou_thr', c_float), ('min_pos_iou', c_float),
        ('gamma', c_float), ('focal_gamma', c_float), ('focal_alpha', c_float),
        ('cls_loss_type', c_int32), ('loss_cls_weight', c_float), ('loss_bbox_weight', c_float),
        ('loss_iou_weight', c_float), ('loss_aug_weight', c_float), ('max_ratio', c_float),
    ]


class RpnWsLayout(Structure):
    _fields_ = [(n, c_int64) for n in (
        'cand_cap', 'keep_cap', 'cand_boxes', 'cand_key', 'cand_valid',
        'cand_count', 'img_maxc', 'kept_pos', 'kept_count', 'total_bytes')]


class RoiParams(Structure):
    _fields_ = [
        ('batch', c_int32), ('channels', c_int32), ('num_levels', c_int32),
        ('feat_h', c_int32 * MAX_LEVELS), ('feat_w', c_int32 * MAX_LEVELS),
        ('spatial_scale', c_float * MAX_LEVELS),
        ('pooled_h', c_int32), ('pooled_w', c_int32),
        ('sampling_ratio', c_int32), ('aligned', c_int32),
        ('finest_scale', c_float), ('out_layout', c_int32),
    ]


class LossParams(Structure):
    _fields_ = [
        ('num_rois', c_int32), ('num_classes', c_int32),
        ('reg_class_agnostic', c_int32), ('gamma', c_float), ('alpha', c_float),
        ('loss_cls_weight', c_float), ('loss_bbox_weight', c_float),
        ('reg_norm_mean', c_int32),
    ]


class AssignParams(Structure):
    _fields_ = [
        ('batch', c_int32), ('max_props', c_int32), ('max_gts', c_int32),
        ('pos_iou_thr', c_float), ('neg_iou_thr', c_float), ('min_pos_iou', c_float),
        ('match_low_quality', c_int32),
    ]


class SampleParams(Structure):
    _fields_ = [
        ('batch', c_int32), ('max_props', c_int32), ('max_gts', c_int32),
        ('num_classes', c_int32), ('perm_cap', c_int32), ('max_sel', c_int32),
        ('means', c_float * 4), ('stds', c_float * 4), ('pos_weight', c_float),
    ]


class RcnnParams(Structure):
    _fields_ = [
        ('batch', c_int32), ('rois_per_img', c_int32), ('num_classes', c_int32),
        ('reg_class_agnostic', c_int32), ('prob', c_int32), ('rescale', c_int32),
        ('means', c_float * 4), ('stds', c_float * 4), ('max_ratio', c_float),
        ('score_thr', c_float), ('iou_threshold', c_float),
        ('max_per_img', c_int32),
    ]


class RcnnWsLayout(Structure):
    _fields_ = [(n, c_int64) for n in (
        'scores', 'bboxes', 'img_maxc', 'seg_count', 'seg_key', 'kept_pos',
        'kept_count', 'total_bytes')]


# name -> (restype, argtypes); must list every symbol of include/brcnn.h
SIGNATURES = {
    'brcnn_version': (c_char_p, []),
    'brcnn_launch_count': (c_int64, []),
    'brcnn_rpn_workspace_layout': (c_int32, [POINTER(RpnParams), POINTER(RpnWsLayout)]),
    'brcnn_rpn_workspace_bytes': (c_size_t, [POINTER(RpnParams)]),
    'brcnn_rpn_get_bboxes': (c_int32, [
        POINTER(RpnParams), POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p),
        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'brcnn_rpn_loss_workspace_bytes': (c_size_t, [POINTER(RpnLossParams)]),
    'brcnn_rpn_loss_forward': (c_int32, [
        POINTER(RpnLossParams), POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p),
        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, POINTER(c_void_p), POINTER(c_void_p),
        POINTER(c_void_p), c_void_p, c_size_t, c_void_p]),
    'brcnn_rpn_loss_scale': (c_int32, [
        POINTER(RpnLossParams), POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p),
        c_void_p, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), c_void_p]),
    'brcnn_delta2bbox': (c_int32, [c_void_p, c_void_p, c_int32, c_int32,
                                   POINTER(c_float), POINTER(c_float), c_float,
                                   c_float, c_float, c_void_p, c_void_p]),
    'brcnn_nms_workspace_bytes': (c_size_t, [c_int32]),
    'brcnn_batched_nms': (c_int32, [
        c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_float, c_int32, c_int32, c_void_p,
        c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'brcnn_bbox2roi_padded': (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p,
                                        c_void_p]),
    'brcnn_map_roi_levels': (c_int32, [c_void_p, c_int32, c_float, c_int32,
                                       c_void_p, c_void_p]),
    'brcnn_roi_extract_forward_workspace_bytes': (c_size_t, [POINTER(RoiParams)]),
    'brcnn_roi_extract_forward': (c_int32, [
        POINTER(RoiParams), POINTER(c_void_p), c_void_p, c_int32, c_void_p,
        c_void_p, c_void_p, c_size_t, c_void_p]),
    'brcnn_roi_extract_backward_workspace_bytes': (c_size_t, [POINTER(RoiParams), c_int32]),
    'brcnn_roi_extract_backward': (c_int32, [
        POINTER(RoiParams), c_void_p, c_void_p, c_int32, POINTER(c_void_p),
        c_void_p, c_size_t, c_void_p]),
    'brcnn_nchw_to_nhwc': (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p]),
    'brcnn_nhwc_to_nchw': (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p]),
    'brcnn_nchw_to_nhwc_multi': (c_int32, [POINTER(c_void_p), POINTER(c_void_p), c_int32, c_int32,
                                           c_int32, POINTER(c_int32), c_void_p]),
    'brcnn_nhwc_to_nchw_multi': (c_int32, [POINTER(c_void_p), POINTER(c_void_p), c_int32, c_int32,
                                           c_int32, POINTER(c_int32), c_void_p]),
    'brcnn_boost_loss_workspace_bytes': (c_size_t, [POINTER(LossParams)]),
    'brcnn_boost_loss': (c_int32, [
        POINTER(LossParams), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'brcnn_rcnn_assign': (c_int32, [POINTER(AssignParams), c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p]),
    'brcnn_rcnn_sample_targets': (c_int32, [POINTER(SampleParams)] + [c_void_p] * 16),
    'brcnn_rcnn_workspace_layout': (c_int32, [POINTER(RcnnParams), POINTER(RcnnWsLayout)]),
    'brcnn_rcnn_workspace_bytes': (c_size_t, [POINTER(RcnnParams)]),
    'brcnn_rcnn_get_bboxes': (c_int32, [
        POINTER(RcnnParams), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
        c_void_p]),
}

_ERRORS = {-1: 'BRCNN_ERR_ARG (bad argument)',
           -2: 'BRCNN_ERR_WORKSPACE (workspace too small)',
           -3: 'BRCNN_ERR_UNSUPPORTED (size outside the supported range)'}

_lib = None


def build(verbose=False):
    """Compile ``csrc/libbrcnn.cu`` for sm_100a into ``libbrcnn.so`` (in-tree)."""
    src = os.path.join(CSRC, 'libbrcnn.cu')
    cmd = ['nvcc'] + NVCC_FLAGS + ['-o', LIB_PATH, src]
    if verbose:
        cmd.insert(1, '-Xptxas=-v')
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + res.stdout + res.stderr)
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB_PATH


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    for f in os.listdir(CSRC):
        if os.path.getmtime(os.path.join(CSRC, f)) > t:
            return True
    hdr = os.path.join(os.path.dirname(_HERE), 'include', 'brcnn.h')
    return os.path.exists(hdr) and os.path.getmtime(hdr) > t


def load():
    """Load libbrcnn.so and bind every symbol; raise if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f'{LIB_PATH} is missing: the sm_100a CUDA library is the only '
            'implementation of this package (no CPU fallback). Build it with '
            '`python -c "import __graft_entry__ as g; g.build()"`.')
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc == 0:
        return
    if rc < 0:
        raise RuntimeError(f'{what}: {_ERRORS.get(rc, rc)}')
    raise RuntimeError(f'{what}: CUDA error {rc}')


def ptr_array(ptrs):
    arr = (c_void_p * len(ptrs))()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return arr
