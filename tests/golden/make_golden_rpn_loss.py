#!/usr/bin/env python
"""Golden vectors for the RPN loss path, produced by EXECUTING the
reference's own Python — same mechanism as make_golden.py / make_golden_train.py.  Build
container only:

    python tests/golden/make_golden_rpn_loss.py     # rewrites reference_golden_rpn_loss.npz

Executed reference code (lifted with `ast`, unmodified):
  * `ATSSRPNHead.loss`, `loss_single`, `get_targets` (mmdet/models/dense_heads/
    atss_rpn_head.py:299-464,505-565) with `atss=False`, i.e. the dispatch of :582-603 to
    `AnchorHead._get_targets_single` / `get_anchors` (anchor_head.py:126-265);
  * `AnchorGenerator` (grid_anchors / valid_flags), `anchor_inside_flags`, `images_to_levels`,
    `unmap`, `multi_apply`; `MaxIoUAssigner` (match_low_quality=True, the RPN train_cfg of
    configs/boosting_rcnn/*), `PseudoSampler`, `AssignResult`, `SamplingResult`,
    `bbox_overlaps`; `delta2bbox` / `bbox2delta`;
  * the loss modules `FocalLoss` (on CPU its forward takes the reference's own
    `py_sigmoid_focal_loss` branch, focal_loss.py:159-177), `VarifocalLoss` + `varifocal_loss`
    (the VOC config's loss_cls; cases `varifocal_*`), `IoULoss` + `iou_loss`,
    `MSELoss` + `mse_loss`, `CrossEntropyLoss(use_sigmoid=True)` + `binary_cross_entropy`,
    `weighted_loss` / `weight_reduce_loss` / `reduce_loss`.
`reduce_mean` is the identity (single process), exactly as dist_utils.py:67-73 behaves
without an initialised process group.  Gradients are those of
sum(loss_rpn_cls) + sum(loss_rpn_bbox) + sum(loss_rpn_iou) (detectors/base.py:186-199).
Inputs are regenerated from a seed by the tests (tests/synth.py::rpn_loss_case).
"""
import ast
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [HERE, os.path.dirname(HERE)]
import synth  # noqa: E402
from make_golden import REF, AttrDict, lift  # noqa: E402
from make_golden_train import lift_classes, reference_namespace  # noqa: E402


def build_reference_head(cls_loss='focal'):
    ns = reference_namespace()
    ns['functools'] = __import__('functools')
    ns['nn'] = torch.nn
    ns['EPS'] = 1e-12
    ns['reduce_mean'] = lambda t: t                      # no process group (dist_utils.py:69-70)
    ns['force_fp32'] = lambda *a, **k: (lambda f: f)
    ns['GHMR'] = type('GHMR', (), {})
    ns['_pair'] = torch.nn.modules.utils._pair
    lift('mmdet/core/anchor/utils.py', ['images_to_levels', 'anchor_inside_flags'], ns)
    lift('mmdet/core/utils/misc.py', ['unmap'], ns)
    lift('mmdet/models/losses/utils.py', ['reduce_loss', 'weight_reduce_loss', 'weighted_loss'], ns)
    # decorated loss functions keep their @weighted_loss (it IS the reduction code)
    for path, fn in (('mmdet/models/losses/iou_loss.py', 'iou_loss'),
                     ('mmdet/models/losses/mse_loss.py', 'mse_loss')):
        raw = {}
        raw.update(ns)
        lift(path, [fn], raw)
        ns[fn] = ns['weighted_loss'](raw[fn])
    lift('mmdet/models/losses/focal_loss.py', ['py_sigmoid_focal_loss'], ns)
    ns['sigmoid_focal_loss'] = None                       # mmcv CUDA op: never reached on CPU
    lift('mmdet/models/losses/cross_entropy_loss.py',
         ['_expand_onehot_labels', 'binary_cross_entropy', 'cross_entropy', 'mask_cross_entropy'], ns)
    lift('mmdet/models/losses/varifocal_loss.py', ['varifocal_loss'], ns)
    lift_classes('mmdet/models/losses/varifocal_loss.py', ['VarifocalLoss'], ns)
    lift_classes('mmdet/models/losses/focal_loss.py', ['FocalLoss'], ns)
    lift_classes('mmdet/models/losses/iou_loss.py', ['IoULoss'], ns)
    lift_classes('mmdet/models/losses/mse_loss.py', ['MSELoss'], ns)
    lift_classes('mmdet/models/losses/cross_entropy_loss.py', ['CrossEntropyLoss'], ns)
    lift_classes('mmdet/core/anchor/anchor_generator.py', ['AnchorGenerator'], ns)

    hns = dict(ns)
    lift('mmdet/models/dense_heads/atss_rpn_head.py',
         ['loss', 'loss_single', 'get_targets', '_get_targets_single'], hns, cls='ATSSRPNHead')
    ans = dict(ns)
    lift('mmdet/models/dense_heads/anchor_head.py', ['get_anchors', '_get_targets_single'], ans,
         cls='AnchorHead')
    # the reference's class chain ATSSRPNHead -> RPNHead -> AnchorHead, reduced to the lifted
    # methods (ATSSRPNHead._get_targets_single dispatches with super(RPNHead, self), :594-603)
    AnchorHead = type('AnchorHead', (), {'get_anchors': ans['get_anchors'],
                                         '_get_targets_single': ans['_get_targets_single']})
    RPNHead = type('RPNHead', (AnchorHead,), {})
    hns['RPNHead'] = RPNHead
    Head = type('ATSSRPNHead', (RPNHead,), {k: hns[k] for k in
                                            ('loss', 'loss_single', 'get_targets',
                                             '_get_targets_single')})
    head = Head()
    head.anchor_generator = ns['AnchorGenerator'](strides=[8, 16, 32, 64, 128], octave_base_scale=4,
                                                  scales_per_octave=3, ratios=[0.5, 1.0, 2.0])
    head.num_classes, head.cls_out_channels, head.use_sigmoid_cls = 1, 1, True
    head.reg_decoded_bbox, head.gamma, head.atss, head.sampling = True, 0.5, False, False
    head.with_aug_loss = True
    head.train_cfg = AttrDict(allowed_border=-1, pos_weight=-1, debug=False)
    head.assigner = ns['MaxIoUAssigner'](pos_iou_thr=0.5, neg_iou_thr=0.5, min_pos_iou=0,
                                         match_low_quality=True, ignore_iof_thr=-1)
    head.sampler = ns['PseudoSampler']()
    head.bbox_coder = types.SimpleNamespace(
        encode=lambda b, g: ns['bbox2delta'](b, g, (0., 0., 0., 0.), (1., 1., 1., 1.)),
        decode=lambda b, p, max_shape=None: ns['delta2bbox'](b, p, (0., 0., 0., 0.),
                                                             (1., 1., 1., 1.), max_shape))
    if cls_loss == 'varifocal':      # the VOC config's RPN classification loss
        head.loss_cls = ns['VarifocalLoss'](use_sigmoid=True, alpha=0.75, gamma=2.0,
                                            iou_weighted=True, loss_weight=1.0)
    else:
        head.loss_cls = ns['FocalLoss'](use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=1.0)
    head.loss_centerness = ns['CrossEntropyLoss'](use_sigmoid=True, loss_weight=1.0)
    head.loss_bbox = ns['IoULoss'](loss_weight=1.0)
    head.aug_loss = ns['MSELoss'](loss_weight=1.0)
    return head


def run_case(head, case):
    c = synth.rpn_loss_case(case)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    cls = [t(a).requires_grad_(True) for a in c['cls']]
    box = [t(a).requires_grad_(True) for a in c['box']]
    iou = [t(a).requires_grad_(True) for a in c['iou']]
    gts = [t(g) for g in c['gt_bboxes']]
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        losses = head.loss(cls, box, iou, gts, c['img_metas'])
    total = sum(sum(v) for v in losses.values())
    total.backward()
    out = {k: np.array([float(x) for x in v], dtype=np.float32) for k, v in losses.items()}
    for l in range(len(cls)):
        out[f'grad_cls_{l}'] = cls[l].grad.numpy()
        out[f'grad_box_{l}'] = box[l].grad.numpy()
        out[f'grad_iou_{l}'] = iou[l].grad.numpy()
    return out


def main():
    head = build_reference_head()
    vhead = build_reference_head('varifocal')
    gold = {}
    for case in synth.RPN_LOSS_CASES + tuple('varifocal_' + c for c in ('basic', 'partial_valid')):
        vf = case.startswith('varifocal_')
        res = run_case(vhead if vf else head, case[len('varifocal_'):] if vf else case)
        for k, v in res.items():
            gold[f'{case}/{k}'] = v
        print(case, {k: res[k].round(4).tolist() for k in ('loss_rpn_cls', 'loss_rpn_bbox', 'loss_rpn_iou')})
    np.savez_compressed(os.path.join(HERE, 'reference_golden_rpn_loss.npz'), **gold)
    print('wrote reference_golden_rpn_loss.npz', len(gold), 'arrays',
          os.path.getsize(os.path.join(HERE, 'reference_golden_rpn_loss.npz')) // 1024, 'KB')


if __name__ == '__main__':
    if not os.path.isdir(REF):
        sys.exit('needs /root/reference (build container only)')
    main()
