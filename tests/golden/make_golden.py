#!/usr/bin/env python
"""Generate golden vectors by EXECUTING the reference's own Python.

Run in the build container only (needs /root/reference; the GPU box has none):

    python tests/golden/make_golden.py          # rewrites tests/golden/*.npz

`import mmdet` is impossible here (mmcv-full 1.3.8-1.4.0 is not installable,
SURVEY.md F3), so the reference FUNCTIONS on the hot path are lifted out of their
source files with `ast` and executed unmodified in a namespace that provides
only what they import: torch, numpy, a no-op `mmcv.jit`, and — for the three
call sites of mmcv's native ops — stand-ins built on torchvision
(`torchvision.ops.nms` uses the same division-form test as mmcv `nms_cpu`;
`batched_nms` is mmcv 1.4.0's Python restated).  What these
fixtures therefore pin: every line of reference Python on the path
(anchor_generator, delta2bbox, _get_bboxes_single, map_roi_levels,
multiclass_nms, cross_entropy/weight_reduce_loss/accuracy, norm_loss, the
fusion lines).  What they do not pin: mmcv's native kernels themselves.
"""
import ast
import os
import sys
import types

import numpy as np
import torch
import torchvision

REF = '/root/reference'
OUT = os.path.dirname(os.path.abspath(__file__))


def lift(path, names, namespace, cls=None):
    """exec the named top-level functions (or methods of `cls`) of a reference
    file, verbatim, inside `namespace`; decorators are dropped."""
    src = open(os.path.join(REF, path)).read()
    tree = ast.parse(src)
    body = tree.body
    if cls is not None:
        body = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls).body
    for node in body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            node.decorator_list = []
            code = ast.get_source_segment(src, node)
            # re-indent methods, strip decorators textually
            lines = code.split('\n')
            indent = len(lines[0]) - len(lines[0].lstrip())
            code = '\n'.join(l[indent:] if l[:indent].strip() == '' else l for l in lines)
            exec(compile(code, path, 'exec'), namespace)
    missing = [n for n in names if n not in namespace]
    assert not missing, (path, missing)


# ---- stand-ins for the un-vendored mmcv pieces --------------------------------
def nms(boxes, scores, iou_threshold, offset=0, score_threshold=0, max_num=-1):
    inds = torchvision.ops.nms(boxes, scores, float(iou_threshold))
    if max_num > 0:
        inds = inds[:max_num]
    return torch.cat((boxes[inds], scores[inds].reshape(-1, 1)), dim=1), inds


def batched_nms(boxes, scores, idxs, nms_cfg, class_agnostic=False):
    nms_cfg_ = nms_cfg.copy()
    class_agnostic = nms_cfg_.pop('class_agnostic', class_agnostic)
    if class_agnostic:
        boxes_for_nms = boxes
    else:
        max_coordinate = boxes.max()
        offsets = idxs.to(boxes) * (max_coordinate + torch.tensor(1).to(boxes))
        boxes_for_nms = boxes + offsets[:, None]
    nms_cfg_.pop('type', 'nms')
    split_thr = nms_cfg_.pop('split_thr', 10000)
    if boxes_for_nms.shape[0] < split_thr:
        dets, keep = nms(boxes_for_nms, scores, **nms_cfg_)
        boxes = boxes[keep]
        scores = dets[:, 4]
    else:
        max_num = nms_cfg_.pop('max_num', -1)
        total_mask = scores.new_zeros(scores.size(), dtype=torch.bool)
        scores_after_nms = scores.new_zeros(scores.size())
        for id in torch.unique(idxs):
            mask = (idxs == id).nonzero(as_tuple=False).view(-1)
            dets, keep = nms(boxes_for_nms[mask], scores[mask], **nms_cfg_)
            total_mask[mask[keep]] = True
            scores_after_nms[mask[keep]] = dets[:, -1]
        keep = total_mask.nonzero(as_tuple=False).view(-1)
        scores, inds = scores_after_nms[keep].sort(descending=True)
        keep = keep[inds]
        boxes = boxes[keep]
        if max_num > 0:
            keep, boxes, scores = keep[:max_num], boxes[:max_num], scores[:max_num]
    return torch.cat([boxes, scores[:, None]], -1), keep


class AttrDict(dict):

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


def base_namespace():
    mmcv = types.SimpleNamespace(jit=lambda *a, **k: (lambda f: f))
    return dict(torch=torch, np=np, mmcv=mmcv, batched_nms=batched_nms, nms=nms,
                warnings=__import__('warnings'), copy=__import__('copy'),
                F=torch.nn.functional)


def main():
    torch.manual_seed(0)
    rng = np.random.RandomState(0)
    gold = {}

    # ---- delta2bbox / bbox2delta (delta_xywh_bbox_coder.py:98-272) ----
    ns = base_namespace()
    lift('mmdet/core/bbox/coder/delta_xywh_bbox_coder.py', ['delta2bbox', 'bbox2delta'], ns)
    rois = torch.tensor(rng.uniform(0, 300, (64, 2)), dtype=torch.float32)
    rois = torch.cat([rois, rois + torch.tensor(rng.uniform(1, 200, (64, 2)), dtype=torch.float32)], 1)
    d1 = torch.tensor(rng.normal(0, 1.5, (64, 4)), dtype=torch.float32)
    d4 = torch.tensor(rng.normal(0, 1.5, (64, 16)), dtype=torch.float32)
    gold['d2b_rois'], gold['d2b_deltas1'], gold['d2b_deltas4'] = rois.numpy(), d1.numpy(), d4.numpy()
    gold['d2b_out1'] = ns['delta2bbox'](rois, d1, max_shape=(320, 400, 3)).numpy()
    gold['d2b_out4'] = ns['delta2bbox'](rois, d4, (0., 0., 0., 0.), (.1, .1, .2, .2),
                                        max_shape=(320, 400, 3)).numpy()
    gold['d2b_out1_noclip'] = ns['delta2bbox'](rois, d1).numpy()
    gold['b2d_out'] = ns['bbox2delta'](rois, rois.flip(0), (0., 0., 0., 0.), (.1, .1, .2, .2)).numpy()

    # ---- AnchorGenerator (anchor_generator.py) ----
    ns = base_namespace()
    ns['_pair'] = torch.nn.modules.utils._pair
    src = open(os.path.join(REF, 'mmdet/core/anchor/anchor_generator.py')).read()
    tree = ast.parse(src)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == 'AnchorGenerator')
    cls.decorator_list = []
    exec(compile(ast.Module(body=[cls], type_ignores=[]), 'anchor_generator.py', 'exec'), ns)
    for tag, kw in (('a9', dict(octave_base_scale=4, scales_per_octave=3, ratios=[0.5, 1.0, 2.0])),
                    ('a1', dict(octave_base_scale=8, scales_per_octave=1, ratios=[1.0]))):
        g = ns['AnchorGenerator'](strides=[8, 16, 32, 64, 128], **kw)
        gold[f'anchor_base_{tag}'] = torch.stack(g.base_anchors).numpy()
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            grid = g.grid_anchors([(5, 7), (3, 4), (2, 2), (1, 1), (1, 1)], device='cpu')
        gold[f'anchor_grid_{tag}'] = torch.cat(grid).numpy()

    # ---- ATSSRPNHead._get_bboxes_single (atss_rpn_head.py:688-760) ----
    ns = base_namespace()
    lift('mmdet/models/dense_heads/atss_rpn_head.py', ['_get_bboxes_single'], ns, cls='ATSSRPNHead')
    dns = base_namespace()
    lift('mmdet/core/bbox/coder/delta_xywh_bbox_coder.py', ['delta2bbox'], dns)
    coder = types.SimpleNamespace(decode=lambda b, p, max_shape=None: dns['delta2bbox'](
        b, p, (0., 0., 0., 0.), (1., 1., 1., 1.), max_shape))
    gen = ns_gen = None
    ans = base_namespace()
    ans['_pair'] = torch.nn.modules.utils._pair
    exec(compile(ast.Module(body=[cls], type_ignores=[]), 'anchor_generator.py', 'exec'), ans)
    gen = ans['AnchorGenerator'](strides=[8, 16, 32, 64, 128], octave_base_scale=4,
                                 scales_per_octave=3, ratios=[0.5, 1.0, 2.0])
    sizes = [(12, 16), (6, 8), (3, 4), (2, 2), (1, 1)]
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        mlvl_anchors = gen.grid_anchors(sizes, device='cpu')
    cls_s = [torch.tensor(rng.normal(0, 1.5, (9, h, w)), dtype=torch.float32) for h, w in sizes]
    box_p = [torch.tensor(rng.normal(0, 0.3, (36, h, w)), dtype=torch.float32) for h, w in sizes]
    iou_p = [torch.tensor(rng.normal(0, 1.5, (9, h, w)), dtype=torch.float32) for h, w in sizes]
    self = types.SimpleNamespace(use_sigmoid_cls=True, bbox_coder=coder, test_cfg=None)
    cfg = AttrDict(nms_pre=60, max_per_img=40, nms=dict(type='nms', iou_threshold=0.7),
                   min_bbox_size=0)
    props = ns['_get_bboxes_single'](self, cls_s, box_p, iou_p, mlvl_anchors, (90, 125, 3),
                                     1.0, cfg)
    for l in range(5):
        gold[f'rpn_cls_{l}'], gold[f'rpn_box_{l}'], gold[f'rpn_iou_{l}'] = (
            cls_s[l].numpy(), box_p[l].numpy(), iou_p[l].numpy())
    gold['rpn_proposals'] = props.numpy()
    gold['rpn_img_shape'] = np.array([90, 125], dtype=np.float32)

    # ---- the same at training size: 61 380 anchors, nms_pre 4000 -> 11 780 candidates, so
    # the (restated) mmcv batched_nms takes its split_thr >= 10000 path (per-level nms +
    # re-sort).  Inputs are re-generated from the seed by the test
    # (synth.rpn_outputs), only the 2000 x 5 result is stored.
    big_sizes = [(64, 80), (32, 40), (16, 20), (8, 10), (4, 5)]
    sys.path.insert(0, os.path.join(os.path.dirname(OUT)))
    import synth
    bcls, bbox, biou = synth.rpn_outputs(1, big_sizes, 9, seed=4242)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        big_anchors = gen.grid_anchors(big_sizes, device='cpu')
    cfg_big = AttrDict(nms_pre=4000, max_per_img=2000, nms=dict(type='nms', iou_threshold=0.7),
                       min_bbox_size=0)
    props_big = ns['_get_bboxes_single'](
        self, [torch.from_numpy(c[0]) for c in bcls], [torch.from_numpy(c[0]) for c in bbox],
        [torch.from_numpy(c[0]) for c in biou], big_anchors, (500, 633, 3), 1.0, cfg_big)
    assert props_big.shape == (2000, 5)
    np.savez_compressed(os.path.join(OUT, 'reference_golden_rpn_train.npz'),
                        proposals=props_big.numpy(), img_shape=np.array([500, 633], np.float32),
                        seed=np.array(4242), sizes=np.array(big_sizes))

    # ---- map_roi_levels (single_level_roi_extractor.py:36-55) ----
    ns = base_namespace()
    lift('mmdet/models/roi_heads/roi_extractors/single_level_roi_extractor.py',
         ['map_roi_levels'], ns, cls='SingleRoIExtractor')
    self = types.SimpleNamespace(finest_scale=56)
    wh = np.exp(rng.uniform(np.log(2), np.log(1500), (400, 2)))
    xy = rng.uniform(0, 300, (400, 2))
    lr = np.concatenate([np.zeros((400, 1)), xy, xy + wh], 1).astype(np.float32)
    edge = []
    for s in (112.0, 224.0, 448.0, 896.0):
        for d in (-1e-2, -1e-3, 0.0, 1e-3, 1e-2):
            edge.append([0, 3, 3, 3 + s + d, 3 + s + d])
    edge += [[0, 5, 5, 5, 5], [0, 0, 0, 1, 1]]
    lr = np.concatenate([lr, np.array(edge, dtype=np.float32)], 0)
    gold['lvl_rois'] = lr
    gold['lvl_out'] = ns['map_roi_levels'](self, torch.from_numpy(lr), 5).numpy()

    # ---- multiclass_nms (bbox_nms.py:8-95) ----
    ns = base_namespace()
    lift('mmdet/core/post_processing/bbox_nms.py', ['multiclass_nms'], ns)
    # COCO scale: 256 RoIs x 80 classes, score_thr 0.001 -> ~20 000 candidates, so the
    # restated mmcv batched_nms runs its split path; inputs are re-generated from the seed
    # by the test (tests/synth.py::multiclass_inputs), only the 100 detections are stored
    import synth as _synth
    mbb, mss = _synth.multiclass_inputs(256, 80, seed=777)
    dd, ll = ns['multiclass_nms'](torch.from_numpy(mbb), torch.from_numpy(mss), 0.001,
                                  dict(type='nms', iou_threshold=0.5), 100)
    assert int((mss[:, :-1] > 0.001).sum()) >= 10000
    np.savez_compressed(os.path.join(OUT, 'reference_golden_multiclass_coco.npz'),
                        dets=dd.numpy(), labels=ll.numpy(), seed=np.array(777))
    R, C = 60, 4
    ctr = rng.uniform(20, 280, (R, 1, 2)) + rng.normal(0, 6, (R, C, 2))
    whc = rng.uniform(20, 120, (R, C, 2))
    mb = np.concatenate([ctr - whc / 2, ctr + whc / 2], -1).reshape(R, C * 4).astype(np.float32)
    ms = rng.uniform(0, 1, (R, C + 1)).astype(np.float32)
    dets, labels = ns['multiclass_nms'](torch.from_numpy(mb), torch.from_numpy(ms), 0.05,
                                        dict(type='nms', iou_threshold=0.5), 30)
    gold['mc_bboxes'], gold['mc_scores'] = mb, ms
    gold['mc_dets'], gold['mc_labels'] = dets.numpy(), labels.numpy()

    # ---- fusion lines (prob_roi_head.py:232-240), restated verbatim ----
    cls_score = torch.tensor(rng.normal(0, 2, (50, 5)), dtype=torch.float32)
    prior = torch.tensor(rng.uniform(0, 1, (50,)), dtype=torch.float32)
    fused = cls_score.softmax(1)
    fused = fused * prior.reshape(-1, 1)
    fused = fused ** 0.5
    gold['fuse_cls'], gold['fuse_prior'], gold['fuse_out'] = cls_score.numpy(), prior.numpy(), fused.numpy()

    # ---- boost loss: cross_entropy + weight_reduce_loss + accuracy + norm_loss ----
    ns = base_namespace()
    lift('mmdet/models/losses/utils.py', ['reduce_loss', 'weight_reduce_loss'], ns)
    lift('mmdet/models/losses/cross_entropy_loss.py', ['cross_entropy'], ns)
    lift('mmdet/models/losses/accuracy.py', ['accuracy'], ns)
    lift('mmdet/models/roi_heads/prob_roi_head.py', ['norm_loss'], ns, cls='ProbRoIHead')
    N, C = 96, 4
    cs = torch.tensor(rng.normal(0, 2, (N, C + 1)), dtype=torch.float32, requires_grad=True)
    labels = torch.full((N,), C, dtype=torch.long)
    labels[:24] = torch.tensor(rng.randint(0, C, 24))
    pri = torch.tensor(rng.uniform(0, 1, (N,)), dtype=torch.float32)
    pri[:3] = 0
    lw = torch.ones(N)
    loss_none = 2.0 * ns['cross_entropy'](cs, labels, lw, reduction='none', avg_factor=float(N))
    w = (1 - pri) ** 0.5
    loss_cls = ns['norm_loss'](None, loss_none, w, w.shape[0])
    loss_cls.backward()
    gold['loss_cls_score'], gold['loss_labels'], gold['loss_prior'] = (
        cs.detach().numpy(), labels.numpy(), pri.numpy())
    gold['loss_cls'] = np.array(loss_cls.item(), dtype=np.float32)
    gold['loss_grad_cls'] = cs.grad.numpy()
    gold['loss_acc'] = np.array(ns['accuracy'](cs.detach(), labels).item(), dtype=np.float32)

    np.savez_compressed(os.path.join(OUT, 'reference_golden.npz'), **gold)
    print('wrote', os.path.join(OUT, 'reference_golden.npz'), len(gold), 'arrays')


if __name__ == '__main__':
    if not os.path.isdir(REF):
        sys.exit('needs /root/reference (build container only)')
    main()
