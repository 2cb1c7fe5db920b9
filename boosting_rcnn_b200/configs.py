"""Hot-path slices of the three named reference configs, as the dicts
``TwoStageDetector.__init__`` hands to ``build_head`` (two_stage.py:37-50 injects
``train_cfg.rpn`` / ``test_cfg.rpn`` into ``rpn_head`` and ``train_cfg.rcnn`` /
``test_cfg.rcnn`` into ``roi_head``).

Values are those of configs/boosting_rcnn/boosting_rcnn_r50_pafpn_1x_utdac.py,
..._pafpn_mstrain_2x_coco.py (BASELINE.json's "1x_coco" does not exist, see
SURVEY.md F2) and ..._pafpn_1x_voc.py.  Backbone / neck / dataset keys are not
reproduced: they are outside the ported path.
"""
import copy

_RPN_TRAIN = dict(
    assigner=dict(type='MaxIoUAssigner', pos_iou_thr=0.5, neg_iou_thr=0.5, min_pos_iou=0,
                  match_low_quality=True, ignore_iof_thr=-1),
    sampler=dict(type='PseudoSampler'),
    allowed_border=-1, pos_weight=-1, debug=False)
_RPN_PROPOSAL_TRAIN = dict(nms_pre=4000, max_per_img=2000,
                           nms=dict(type='nms', iou_threshold=0.7), min_bbox_size=0)
_RPN_TEST = dict(nms_pre=1000, max_per_img=256, nms=dict(type='nms', iou_threshold=0.7),
                 min_bbox_size=0)


def _rcnn_train(iou):
    return dict(
        assigner=dict(type='MaxIoUAssigner', pos_iou_thr=iou, neg_iou_thr=iou, min_pos_iou=iou,
                      match_low_quality=False, ignore_iof_thr=-1),
        sampler=dict(type='RandomSampler', num=512, pos_fraction=0.25, neg_pos_ub=-1,
                     add_gt_as_proposals=True),
        pos_weight=-1, debug=False)


def _model(num_classes, rpn_gamma, rpn_loss_cls, rpn_w, anchor, bbox_head_extra, roi_extra,
           rcnn_iou_train, rcnn_nms_iou):
    return dict(
        rpn_head=dict(
            type='ATSSRPNHead', in_channels=256, feat_channels=256, stacked_convs=4,
            reg_decoded_bbox=True, gamma=rpn_gamma, atss=False,
            anchor_generator=dict(type='AnchorGenerator', strides=[8, 16, 32, 64, 128], **anchor),
            bbox_coder=dict(type='DeltaXYWHBBoxCoder', target_means=[.0, .0, .0, .0],
                            target_stds=[1.0, 1.0, 1.0, 1.0]),
            loss_cls=rpn_loss_cls,
            loss_centerness=dict(type='CrossEntropyLoss', use_sigmoid=True, loss_weight=1.0),
            loss_bbox=dict(type='IoULoss', loss_weight=rpn_w),
            aug_reg_loss=dict(type='MSELoss', loss_weight=rpn_w)),
        roi_head=dict(
            type='ProbRoIHead', boost=True, gamma=0.5,
            bbox_roi_extractor=dict(
                type='SingleRoIExtractor',
                roi_layer=dict(type='RoIAlign', output_size=7, sampling_ratio=0),
                out_channels=256, featmap_strides=[8, 16, 32, 64, 128]),
            bbox_head=dict(
                type='ProbConvFCBBoxHead', in_channels=256, fc_out_channels=1024,
                roi_feat_size=7, num_classes=num_classes,
                bbox_coder=dict(type='DeltaXYWHBBoxCoder', target_means=[0., 0., 0., 0.],
                                target_stds=[0.1, 0.1, 0.2, 0.2]),
                reg_class_agnostic=False,
                loss_cls=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=2.0),
                loss_bbox=dict(type='L1Loss', loss_weight=2.0), **bbox_head_extra),
            **roi_extra),
        train_cfg=dict(rpn=copy.deepcopy(_RPN_TRAIN),
                       rpn_proposal=copy.deepcopy(_RPN_PROPOSAL_TRAIN),
                       rcnn=_rcnn_train(rcnn_iou_train)),
        test_cfg=dict(rpn=copy.deepcopy(_RPN_TEST),
                      rcnn=dict(score_thr=0.05, nms=dict(type='nms', iou_threshold=rcnn_nms_iou),
                                max_per_img=100)))


_FOCAL = dict(type='FocalLoss', use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=1.0)
_A9 = dict(octave_base_scale=4, scales_per_octave=3, ratios=[0.5, 1.0, 2.0])

MODELS = {
    # boosting_rcnn_r50_pafpn_1x_utdac.py
    'utdac': lambda: _model(4, 0.5, dict(_FOCAL), 1.0, dict(_A9), dict(num_shared_fcs=2), dict(),
                            0.6, 0.7),
    # boosting_rcnn_r50_pafpn_mstrain_2x_coco.py
    'coco': lambda: _model(80, 2, dict(_FOCAL), 2.0, dict(_A9), dict(num_shared_fcs=2), dict(),
                           0.6, 0.5),
    # boosting_rcnn_r50_pafpn_1x_voc.py
    'voc': lambda: _model(
        20, 2, dict(type='VarifocalLoss', use_sigmoid=True, alpha=0.75, gamma=2.0,
                    iou_weighted=True, loss_weight=1.0), 2.0,
        dict(ratios=[1.0], octave_base_scale=8, scales_per_octave=1),
        dict(num_cls_fcs=2, num_reg_convs=4,
             norm_cfg=dict(type='GN', num_groups=32, requires_grad=True)),
        dict(alpha=0, reg_norm='num_bbox', quality=False, iou_gamma=1), 0.5, 0.5),
}

# (img_shape, pad_shape) of the named workloads: Resize keep_ratio then
# Pad(size_divisor=32) (_base_/datasets/utdac_detection_coco.py; VOC :134-160)
IMAGE_GEOMETRY = {
    'utdac': dict(img_shape=(800, 1333, 3), pad_shape=(800, 1344, 3)),
    'coco': dict(img_shape=(800, 1333, 3), pad_shape=(800, 1344, 3)),
    'voc': dict(img_shape=(600, 1000, 3), pad_shape=(608, 1024, 3)),
}


def model_cfg(name):
    return MODELS[name]()


def build_hot_path(name, train=False):
    """Build (rpn_head, roi_head) the way TwoStageDetector.__init__ does."""
    from . import bbox_head, coder, roi_extractor, roi_head, rpn_head  # noqa: F401 (register)
    from .registry import build_head
    m = model_cfg(name)
    rpn_cfg, roi_cfg = dict(m['rpn_head']), dict(m['roi_head'])
    rpn_cfg.update(train_cfg=m['train_cfg']['rpn'] if train else None,
                   test_cfg=m['test_cfg']['rpn'])
    roi_cfg.update(train_cfg=m['train_cfg']['rcnn'] if train else None,
                   test_cfg=m['test_cfg']['rcnn'])
    return build_head(rpn_cfg), build_head(roi_cfg), m
