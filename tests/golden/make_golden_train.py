#!/usr/bin/env python
"""Golden vectors for the R-CNN training front-end, produced by
EXECUTING the reference's own Python — same mechanism as make_golden.py (classes / methods
lifted out of /root/reference with `ast`, run unmodified in a namespace that provides what
they import).  Build container only:

    python tests/golden/make_golden_train.py     # rewrites reference_golden_train.npz

Executed reference code:
  * `ProbRoIHead.forward_train` (mmdet/models/roi_heads/prob_roi_head.py:23-88) — the
    assign / sample loop and the prior extraction :51-64; its `_bbox_forward_train_boost`
    is replaced by a recorder that runs the reference's `bbox2roi` (core/bbox/transforms.py
    :59-78) and `BBoxHead.get_targets` (bbox_heads/bbox_head.py:122-253);
  * `MaxIoUAssigner` (core/bbox/assigners/max_iou_assigner.py), `AssignResult`
    (assign_result.py incl. add_gt_), `BboxOverlaps2D` / `bbox_overlaps`
    (iou_calculators/iou2d_calculator.py), `RandomSampler` / `BaseSampler` /
    `SamplingResult` (samplers/*.py), `bbox2delta` (coder/delta_xywh_bbox_coder.py).
Inputs are regenerated from a seed by the tests (tests/synth.py::rcnn_train_case); the CPU
RNG is seeded with torch.manual_seed(123) before each case, only outputs are stored.
"""
import ast
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [HERE, os.path.dirname(HERE)]
import synth  # noqa: E402
from make_golden import REF, AttrDict, base_namespace, lift  # noqa: E402

SEED = 123
TRAIN_RCNN = dict(pos_iou_thr=0.6, neg_iou_thr=0.6, min_pos_iou=0.6, match_low_quality=False,
                  ignore_iof_thr=-1, num=512, pos_fraction=0.25, neg_pos_ub=-1,
                  add_gt_as_proposals=True, pos_weight=-1)
MEANS, STDS = (0., 0., 0., 0.), (0.1, 0.1, 0.2, 0.2)


def lift_classes(path, names, ns):
    """exec the named top-level classes of a reference file verbatim (decorators dropped)."""
    src = open(os.path.join(REF, path)).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name in names:
            node.decorator_list = []
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, 'exec'), ns)
    missing = [n for n in names if n not in ns]
    assert not missing, (path, missing)


def reference_namespace():
    ns = base_namespace()
    ns['ABCMeta'] = __import__('abc').ABCMeta
    ns['abstractmethod'] = __import__('abc').abstractmethod
    ns['util_mixins'] = types.SimpleNamespace(NiceRepr=type('NiceRepr', (), {}))
    # `from mmdet.core.bbox import demodata` inside RandomSampler.__init__
    demodata = types.ModuleType('mmdet.core.bbox.demodata')
    demodata.ensure_rng = lambda rng=None: np.random.mtrand._rand if rng is None else rng
    for name in ('mmdet', 'mmdet.core', 'mmdet.core.bbox'):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules['mmdet.core.bbox.demodata'] = demodata
    sys.modules['mmdet.core.bbox'].demodata = demodata
    lift('mmdet/core/bbox/iou_calculators/iou2d_calculator.py',
         ['cast_tensor_type', 'fp16_clamp', 'bbox_overlaps'], ns)
    lift_classes('mmdet/core/bbox/iou_calculators/iou2d_calculator.py', ['BboxOverlaps2D'], ns)
    ns['build_iou_calculator'] = lambda cfg: ns['BboxOverlaps2D']()
    lift_classes('mmdet/core/bbox/assigners/assign_result.py', ['AssignResult'], ns)
    lift_classes('mmdet/core/bbox/assigners/base_assigner.py', ['BaseAssigner'], ns)
    lift_classes('mmdet/core/bbox/assigners/max_iou_assigner.py', ['MaxIoUAssigner'], ns)
    lift_classes('mmdet/core/bbox/samplers/sampling_result.py', ['SamplingResult'], ns)
    lift_classes('mmdet/core/bbox/samplers/base_sampler.py', ['BaseSampler'], ns)
    lift_classes('mmdet/core/bbox/samplers/random_sampler.py', ['RandomSampler'], ns)
    lift_classes('mmdet/core/bbox/samplers/pseudo_sampler.py', ['PseudoSampler'], ns)
    lift('mmdet/core/bbox/coder/delta_xywh_bbox_coder.py', ['bbox2delta', 'delta2bbox'], ns)
    lift('mmdet/core/bbox/transforms.py', ['bbox2roi'], ns)
    lift('mmdet/core/utils/misc.py', ['multi_apply'], ns)
    ns['partial'] = __import__('functools').partial
    ns['map'] = map
    ns['six'] = types.SimpleNamespace(moves=types.SimpleNamespace(map=map, zip=zip))
    return ns


def run_case(ns, case):
    gts, labels, plist = synth.rcnn_train_case(case)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    gts_t, labels_t, plist_t = [t(g) for g in gts], [t(l) for l in labels], [t(p) for p in plist]
    c = TRAIN_RCNN
    assigner = ns['MaxIoUAssigner'](pos_iou_thr=c['pos_iou_thr'], neg_iou_thr=c['neg_iou_thr'],
                                    min_pos_iou=c['min_pos_iou'],
                                    match_low_quality=c['match_low_quality'],
                                    ignore_iof_thr=c['ignore_iof_thr'])
    sampler = ns['RandomSampler'](num=c['num'], pos_fraction=c['pos_fraction'],
                                  neg_pos_ub=c['neg_pos_ub'],
                                  add_gt_as_proposals=c['add_gt_as_proposals'])
    # BBoxHead.get_targets / _get_target_single executed against a minimal `self`
    hns = dict(ns)
    lift('mmdet/models/roi_heads/bbox_heads/bbox_head.py', ['get_targets', '_get_target_single'],
         hns, cls='BBoxHead')
    coder = types.SimpleNamespace(encode=lambda b, g: ns['bbox2delta'](b, g, MEANS, STDS))
    head = types.SimpleNamespace(num_classes=80, reg_decoded_bbox=False, bbox_coder=coder)
    head._get_target_single = types.MethodType(hns['_get_target_single'], head)
    head.get_targets = types.MethodType(hns['get_targets'], head)
    train_cfg = AttrDict(pos_weight=c['pos_weight'])
    rec = {}

    def recorder(self, x, sampling_results, gt_bboxes, gt_labels, img_metas, priors, ious=None):
        rec['rois'] = ns['bbox2roi']([res.bboxes for res in sampling_results])
        rec['targets'] = head.get_targets(sampling_results, gt_bboxes, gt_labels, train_cfg)
        rec['prior'] = priors
        rec['results'] = sampling_results
        return dict(loss_bbox=dict())

    rns = dict(ns)
    lift('mmdet/models/roi_heads/prob_roi_head.py', ['forward_train'], rns, cls='ProbRoIHead')
    roi_head = types.SimpleNamespace(with_bbox=True, with_mask=False, quality=False, boost=True,
                                     bbox_assigner=assigner, bbox_sampler=sampler)
    roi_head._bbox_forward_train_boost = types.MethodType(recorder, roi_head)
    torch.manual_seed(SEED)
    x = [torch.zeros(len(plist), 1, 2, 2)]
    rns['forward_train'](roi_head, x, [dict() for _ in plist], plist_t, gts_t, labels_t)
    lab, lw, bt, bw = rec['targets']
    out = {
        'rois': rec['rois'].numpy(), 'labels': lab.numpy(), 'label_weights': lw.numpy(),
        'bbox_targets': bt.numpy(), 'bbox_weights': bw.numpy(), 'prior': rec['prior'].numpy(),
        'rows': np.array([r.bboxes.size(0) for r in rec['results']], dtype=np.int64),
        'num_pos': np.array([r.pos_inds.numel() for r in rec['results']], dtype=np.int64),
    }
    for b, r in enumerate(rec['results']):
        out[f'pos_inds_{b}'] = r.pos_inds.numpy()
        out[f'neg_inds_{b}'] = r.neg_inds.numpy()
    return out


def main():
    ns = reference_namespace()
    gold = {}
    for case in synth.RCNN_TRAIN_CASES:
        for k, v in run_case(ns, case).items():
            gold[f'{case}/{k}'] = v
        print(case, 'rows', gold[f'{case}/rows'], 'num_pos', gold[f'{case}/num_pos'])
    # reference KAT of the assigner with match_low_quality=True (the RPN-stage setting,
    # tests/test_utils/test_assigner.py:16-37) plus a random larger case
    rng = np.random.RandomState(5)
    a = ns['MaxIoUAssigner'](pos_iou_thr=0.5, neg_iou_thr=0.5, min_pos_iou=0.0,
                             match_low_quality=True, ignore_iof_thr=-1)
    bx = synth.random_boxes(3000, 250, 317, seed=90)
    gb = synth.random_boxes(12, 250, 317, seed=91)
    gb[3] = [400, 400, 410, 410]           # a GT no box overlaps: min_pos_iou=0 quirk
    res = a.assign(torch.from_numpy(bx), torch.from_numpy(gb))
    gold['mlq/boxes'], gold['mlq/gts'] = bx, gb
    gold['mlq/gt_inds'] = res.gt_inds.numpy()
    gold['mlq/max_overlaps'] = res.max_overlaps.numpy()
    np.savez_compressed(os.path.join(HERE, 'reference_golden_train.npz'), **gold)
    print('wrote reference_golden_train.npz', len(gold), 'arrays')


if __name__ == '__main__':
    if not os.path.isdir(REF):
        sys.exit('needs /root/reference (build container only)')
    main()
