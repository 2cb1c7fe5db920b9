"""CPU: pin the oracle (oracle/brcnn_oracle.c) before trusting it.

1. against tests/golden/reference_golden.npz — outputs of the reference's OWN
   Python functions executed by tests/golden/make_golden.py;
2. against the known-answer vectors held by the reference's test-suite
   (tests/test_utils/test_anchor.py:548-646, tests/test_utils/test_coder.py:
   27-75, tests/test_metrics/test_losses.py:186-240);
3. against torchvision.ops.{nms,roi_align} — the algorithm mmcv delegates to
   under use_torchvision=True — for the two mmcv-native ops.
"""
import os

import numpy as np
import pytest
import torch
import torchvision

import synth
from boosting_rcnn_b200.anchors import AnchorGenerator
from oracle import oracle

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_golden.npz'))


# ------------------------------------------------------------------ pinned math
def test_pinned_expf_accuracy_and_edges():
    x = np.concatenate([np.linspace(-104, 88.7, 400001), [0.0, -0.0, 1e-8, -1e-8]]).astype(np.float32)
    e = oracle.expf(x).astype(np.float64)
    ref = np.exp(x.astype(np.float64))
    ok = ref > 1e-37  # normal range
    ulp = np.abs(e[ok] - ref[ok]) / np.spacing(ref[ok].astype(np.float32)).astype(np.float64)
    assert ulp.max() < 1.1, ulp.max()  # measured 1.004 ulp worst case
    sub = ~ok
    assert np.all(np.abs(e[sub] - ref[sub]) <= 2e-45 + 1e-6 * ref[sub])
    assert oracle.expf(np.array([200.0], np.float32))[0] == np.inf
    assert oracle.expf(np.array([-200.0], np.float32))[0] == 0.0
    assert np.isnan(oracle.expf(np.array([np.nan], np.float32))[0])
    assert oracle.expf(np.array([0.0], np.float32))[0] == 1.0


def test_pinned_sigmoid_matches_torch_within_ulps():
    x = np.random.RandomState(0).normal(0, 4, 200000).astype(np.float32)
    mine = oracle.sigmoid(x)
    ref = torch.from_numpy(x).sigmoid().numpy()
    rel = np.abs(mine.astype(np.float64) - ref) / np.maximum(ref, 1e-30)
    assert rel.max() < 4e-7  # <= ~3 ulp of the reference's own (libm-dependent) sigmoid


# ------------------------------------------------------------------ anchors
EXPECTED_BASE_ANCHORS_L0 = np.array(  # tests/test_utils/test_anchor.py:582-590
    [[-22.6274, -11.3137, 22.6274, 11.3137], [-28.5088, -14.2544, 28.5088, 14.2544],
     [-35.9188, -17.9594, 35.9188, 17.9594], [-16.0000, -16.0000, 16.0000, 16.0000],
     [-20.1587, -20.1587, 20.1587, 20.1587], [-25.3984, -25.3984, 25.3984, 25.3984],
     [-11.3137, -22.6274, 11.3137, 22.6274], [-14.2544, -28.5088, 14.2544, 28.5088],
     [-17.9594, -35.9188, 17.9594, 35.9188]], dtype=np.float32)


def test_base_anchors_match_reference_kat_and_executed_reference():
    g = AnchorGenerator(strides=[8, 16, 32, 64, 128], ratios=[0.5, 1.0, 2.0],
                        octave_base_scale=4, scales_per_octave=3)
    assert g.num_base_anchors == [9] * 5  # test_anchor.py:640
    t = g.base_anchor_table().numpy()
    for l in range(5):  # every level is level 0 scaled by 2**l (test_anchor.py:581-627)
        assert np.allclose(t[l], EXPECTED_BASE_ANCHORS_L0 * 2 ** l, rtol=1e-5, atol=1e-4 * 2 ** l)
    np.testing.assert_array_equal(t, G['anchor_base_a9'])  # bit-exact vs executed reference
    g1 = AnchorGenerator(strides=[8, 16, 32, 64, 128], ratios=[1.0], octave_base_scale=8,
                         scales_per_octave=1)
    np.testing.assert_array_equal(g1.base_anchor_table().numpy(), G['anchor_base_a1'])
    grid = torch.cat(g.grid_anchors_cpu([(5, 7), (3, 4), (2, 2), (1, 1), (1, 1)])).numpy()
    np.testing.assert_array_equal(grid, G['anchor_grid_a9'])


# ------------------------------------------------------------------ delta2bbox
def test_delta2bbox_reference_kats():
    # tests/test_utils/test_coder.py:27-40 and the doctest delta_xywh_bbox_coder.py:191-204
    rois = np.array([[0., 0., 1., 1.], [0., 0., 1., 1.], [0., 0., 1., 1.], [5., 5., 5., 5.]], np.float32)
    deltas = np.array([[0., 0., 0., 0.], [1., 1., 1., 1.], [0., 0., 2., -1.],
                       [0.7, -1.9, -0.5, 0.3]], np.float32)
    expected = np.array([[0.0000, 0.0000, 1.0000, 1.0000], [0.1409, 0.1409, 2.8591, 2.8591],
                         [0.0000, 0.3161, 4.1945, 0.6839], [5.0000, 5.0000, 5.0000, 5.0000]], np.float32)
    out = oracle.delta2bbox(rois, deltas, max_shape=(32, 32))
    assert np.allclose(out, expected, atol=1e-4)
    assert oracle.delta2bbox(np.zeros((0, 4), np.float32), np.zeros((0, 4), np.float32),
                             max_shape=(32, 32)).shape == (0, 4)


def test_delta2bbox_vs_executed_reference():
    out1 = oracle.delta2bbox(G['d2b_rois'], G['d2b_deltas1'], max_shape=(320, 400))
    out4 = oracle.delta2bbox(G['d2b_rois'], G['d2b_deltas4'], stds=(.1, .1, .2, .2), max_shape=(320, 400))
    outn = oracle.delta2bbox(G['d2b_rois'], G['d2b_deltas1'])
    # the reference's exp() is libm's; the oracle pins it (<= 1 ulp) -> 1e-5 relative
    for mine, ref in ((out1, G['d2b_out1']), (out4, G['d2b_out4']), (outn, G['d2b_out1_noclip'])):
        assert np.allclose(mine, ref, rtol=1e-5, atol=1e-4)


# ------------------------------------------------------------------ RPN path
def test_rpn_get_bboxes_single_vs_executed_reference():
    g = AnchorGenerator(strides=[8, 16, 32, 64, 128], ratios=[0.5, 1.0, 2.0],
                        octave_base_scale=4, scales_per_octave=3)
    props = oracle.rpn_get_bboxes_single(
        [G[f'rpn_cls_{l}'] for l in range(5)], [G[f'rpn_box_{l}'] for l in range(5)],
        [G[f'rpn_iou_{l}'] for l in range(5)], g.base_anchor_table().numpy(), synth.STRIDES,
        tuple(G['rpn_img_shape']), nms_pre=60, max_per_img=40, iou_threshold=0.7)
    ref = G['rpn_proposals']
    assert props.shape == ref.shape  # same number of proposals, same order
    assert np.allclose(props[:, 4], ref[:, 4], rtol=1e-6, atol=1e-7)
    assert np.allclose(props[:, :4], ref[:, :4], rtol=1e-5, atol=1e-4)


def test_map_roi_levels_vs_executed_reference():
    lv = oracle.map_roi_levels(G['lvl_rois'], 5, 56)
    np.testing.assert_array_equal(lv, G['lvl_out'])


def test_map_roi_levels_dense_boundary_sweep_vs_torch_log2():
    # compare-form level mapping == floor(log2(.)) of the reference on a dense
    # set of scales straddling every threshold by a few ulp (SURVEY.md hard part vi)
    rois = []
    for s in (112.0, 224.0, 448.0, 896.0):
        base = np.float32(s)
        for k in range(-40, 41):
            v = base
            for _ in range(abs(k)):
                v = np.nextafter(v, np.float32(np.inf if k > 0 else -np.inf), dtype=np.float32)
            rois.append([0, 0, 0, v, v])
    rois = np.array(rois, dtype=np.float32)
    t = torch.from_numpy(rois)
    scale = torch.sqrt((t[:, 3] - t[:, 1]) * (t[:, 4] - t[:, 2]))
    ref = torch.floor(torch.log2(scale / 56 + 1e-6)).clamp(min=0, max=4).long().numpy()
    mine = oracle.map_roi_levels(rois, 5, 56)
    bad = np.nonzero(mine != ref)[0]
    assert len(bad) == 0, f'{len(bad)} boundary disagreements, first at roi {rois[bad[0]]}'


# ------------------------------------------------------------------ mmcv-native stand-ins
@pytest.mark.parametrize('n,clustered', [(200, False), (1000, True), (3000, True)])
def test_nms_cpu_vs_torchvision(n, clustered):
    boxes = synth.random_boxes(n, 800, 1333, seed=n, clustered=clustered)
    scores = np.random.RandomState(n).permutation(n).astype(np.float32) / n  # unique
    for thr in (0.5, 0.7):
        keep = oracle.nms_cpu(boxes, scores, thr)
        ref = torchvision.ops.nms(torch.from_numpy(boxes), torch.from_numpy(scores), thr).numpy()
        np.testing.assert_array_equal(keep, ref)


def test_batched_nms_split_equals_unsplit():
    boxes = synth.random_boxes(2500, 600, 1000, seed=3, clustered=True)
    scores = (np.round(np.random.RandomState(4).rand(2500) * 100) / 100).astype(np.float32)
    ids = np.random.RandomState(5).randint(0, 6, 2500)
    d1, k1 = oracle.batched_nms(boxes, scores, ids, 0.6, split_thr=10000)
    d2, k2 = oracle.batched_nms(boxes, scores, ids, 0.6, split_thr=100)  # forces the split path
    np.testing.assert_array_equal(k1, k2)
    np.testing.assert_array_equal(d1, d2)


def test_multiclass_nms_vs_executed_reference():
    mb, ms = G['mc_bboxes'], G['mc_scores']
    R, C = ms.shape[0], ms.shape[1] - 1
    rois = np.zeros((R, 5), np.float32)
    # decode is bypassed: feed already-decoded boxes through the NMS half
    import ctypes
    cb, cs, cl = [], [], []
    for r in range(R):
        for c in range(C):
            if ms[r, c] > 0.05:
                cb.append(mb[r, c * 4:c * 4 + 4]); cs.append(ms[r, c]); cl.append(c)
    dets, keep = oracle.batched_nms(np.array(cb), np.array(cs), np.array(cl), 0.5)
    dets, labels = dets[:30], np.array(cl)[keep][:30]
    np.testing.assert_array_equal(labels, G['mc_labels'])
    np.testing.assert_array_equal(dets, G['mc_dets'])


@pytest.mark.parametrize('scale,H,W', [(1 / 8, 50, 84), (1 / 32, 13, 21)])
def test_roi_align_forward_backward_vs_torchvision(scale, H, W):
    rng = np.random.RandomState(1)
    feat = rng.normal(0, 1, (2, 8, H, W)).astype(np.float32)
    rois = synth.random_rois(2, 40, 400, 672, seed=2)
    extra = np.array([[0, -30, -30, 20, 20], [1, 600, 350, 700, 420], [0, 5, 5, 5, 5],
                      [1, 0, 0, 672, 400]], np.float32)
    rois = np.concatenate([rois, extra])
    out = oracle.roi_align_forward(feat, rois, 7, scale)
    tf = torch.from_numpy(feat).requires_grad_(True)
    ref = torchvision.ops.roi_align(tf, torch.from_numpy(rois), 7, scale, sampling_ratio=0, aligned=True)
    assert np.allclose(out, ref.detach().numpy(), rtol=1e-5, atol=1e-5)
    g = rng.normal(0, 1, out.shape).astype(np.float32)
    ref.backward(torch.from_numpy(g))
    gi = oracle.roi_align_backward(g, rois, feat.shape, scale)
    assert np.allclose(gi, tf.grad.numpy(), rtol=1e-4, atol=1e-4)


# ------------------------------------------------------------------ fusion + loss
def test_fusion_vs_executed_reference():
    out = oracle.fuse_scores(G['fuse_cls'], G['fuse_prior'])
    assert np.allclose(out, G['fuse_out'], rtol=2e-6, atol=1e-7)


def test_boost_loss_vs_executed_reference():
    N, C1 = G['loss_cls_score'].shape
    z4 = np.zeros((N, 4), np.float32)
    o = oracle.boost_loss(G['loss_cls_score'], G['loss_labels'], G['loss_prior'],
                          np.zeros((N, 4 * (C1 - 1)), np.float32), z4, z4, C1 - 1, gamma=0.5,
                          loss_cls_weight=2.0, loss_bbox_weight=2.0)
    assert abs(o['loss_cls'] - G['loss_cls']) <= 1e-5 * abs(G['loss_cls'])
    assert abs(o['acc'] - G['loss_acc']) <= 1e-4
    scale = np.abs(G['loss_grad_cls']).max()
    assert np.abs(o['grad_cls'] - G['loss_grad_cls']).max() <= 1e-5 * scale


def test_accuracy_reference_kat():
    # tests/test_metrics/test_losses.py:193-201: top-1 accuracy 100 on this fixture
    pred = np.array([[0.2, 0.3, 0.6, 0.5], [0.1, 0.1, 0.2, 0.6], [0.9, 0.0, 0.0, 0.1],
                     [0.4, 0.7, 0.1, 0.1], [0.0, 0.0, 0.99, 0]], np.float32)
    label = np.array([2, 3, 0, 1, 2])
    z = np.zeros((5, 4), np.float32)
    o = oracle.boost_loss(pred, label, np.zeros(5, np.float32), np.zeros((5, 12), np.float32), z, z, 3)
    assert o['acc'] == 100.0


def _match_rows_allowing_near_tie_swaps(mine, ref, window=2):
    """Every reference row must appear in `mine` within `window` positions: the reference's
    torch.sigmoid and the pinned exp differ by <= 2 ulp and torch.sort is unstable, so rows
    whose scores agree to ~1e-7 may swap places; nothing else may differ."""
    assert mine.shape == ref.shape
    assert np.allclose(mine[:, 4], ref[:, 4], rtol=1e-6, atol=1e-7)
    used = np.zeros(len(mine), bool)
    swapped = 0
    for i, r in enumerate(ref):
        lo, hi = max(0, i - window), min(len(mine), i + window + 1)
        ok = [j for j in range(lo, hi)
              if not used[j] and np.allclose(mine[j, :4], r[:4], rtol=1e-5, atol=1e-3)]
        assert ok, f'reference row {i} has no counterpart'
        j = min(ok, key=lambda j: abs(j - i))
        used[j] = True
        swapped += (j != i)
    return swapped


def _rpn_train_golden_inputs():
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_golden_rpn_train.npz'))
    sizes = [tuple(int(v) for v in s) for s in g['sizes']]
    cls, box, iou = synth.rpn_outputs(1, sizes, 9, seed=int(g['seed']))
    return g, sizes, cls, box, iou


def test_rpn_train_size_split_path_vs_executed_reference():
    """61 380 anchors, nms_pre 4000 / max 2000 (11 780 candidates -> mmcv batched_nms split
    path) executed by the reference's own _get_bboxes_single: same 2000 proposals."""
    g, sizes, cls, box, iou = _rpn_train_golden_inputs()
    gen = AnchorGenerator(strides=[8, 16, 32, 64, 128], ratios=[0.5, 1.0, 2.0],
                          octave_base_scale=4, scales_per_octave=3)
    props = oracle.rpn_get_bboxes_single(
        [c[0] for c in cls], [c[0] for c in box], [c[0] for c in iou],
        gen.base_anchor_table().numpy(), synth.STRIDES, tuple(g['img_shape']), 4000, 2000, 0.7, 0.0)
    swapped = _match_rows_allowing_near_tie_swaps(props, g['proposals'])
    assert swapped <= 20, swapped


def _multiclass_coco_candidates():
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden',
                             'reference_golden_multiclass_coco.npz'))
    mb, ms = synth.multiclass_inputs(256, 80, seed=int(g['seed']))
    R, C = ms.shape[0], ms.shape[1] - 1
    valid = ms[:, :C] > 0.001                      # bbox_nms.py:55, row-major (roi, class)
    rr, cc = np.nonzero(valid)
    boxes = mb.reshape(R, C, 4)[rr, cc]
    return g, boxes, ms[rr, cc], cc


def test_multiclass_nms_coco_scale_split_path_vs_executed_reference():
    """256 RoIs x 80 classes above score_thr (> 10 000 candidates -> mmcv batched_nms split
    path) executed by the reference's multiclass_nms: identical 100 detections + labels."""
    g, boxes, scores, labels = _multiclass_coco_candidates()
    assert len(scores) >= 10000
    dets, keep = oracle.batched_nms(boxes, scores, labels, 0.5)
    np.testing.assert_array_equal(labels[keep][:100], g['labels'])
    np.testing.assert_array_equal(dets[:100].view(np.uint32), g['dets'].view(np.uint32))


# ---------------------------------------------------------------------------
# R-CNN training front-end: the numpy oracle against goldens produced by
# executing the reference's MaxIoUAssigner / RandomSampler / get_targets / forward_train
# (tests/golden/make_golden_train.py)
# ---------------------------------------------------------------------------
TRAIN_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden',
                          'reference_golden_train.npz')


def _train_prep_oracle(case, seed=123):
    import torch
    gts, labels, plist = synth.rcnn_train_case(case)
    torch.manual_seed(seed)      # the reference draws torch.randperm on the CPU generator
    return oracle.rcnn_train_prep(
        plist, gts, labels, 80, 0.6, 0.6, 0.6, 512, 0.25, -1, (0., 0., 0., 0.),
        (0.1, 0.1, 0.2, 0.2), randperm=lambda n: torch.randperm(n).numpy())


@pytest.mark.parametrize('case', synth.RCNN_TRAIN_CASES)
def test_train_prep_oracle_equals_executed_reference(case):
    g = np.load(TRAIN_GOLD)
    o = _train_prep_oracle(case)
    assert list(o['rows']) == list(g[f'{case}/rows'])
    for b in range(len(o['rows'])):
        np.testing.assert_array_equal(o['pos_inds'][b], g[f'{case}/pos_inds_{b}'])
        np.testing.assert_array_equal(o['neg_inds'][b], g[f'{case}/neg_inds_{b}'])
    for k in ('rois', 'label_weights', 'bbox_weights', 'prior'):
        np.testing.assert_array_equal(o[k].view(np.uint32), g[f'{case}/{k}'].view(np.uint32), k)
    np.testing.assert_array_equal(o['labels'], g[f'{case}/labels'])
    np.testing.assert_allclose(o['bbox_targets'], g[f'{case}/bbox_targets'], rtol=1e-6, atol=1e-6)
    if case == 'many_positives':
        # GTs were dropped by randperm: the reference's pos_inds[num_gts:] slice is misaligned
        # (prior 0 on non-GT rows) and the oracle reproduces exactly that
        assert all(n == 128 for n in g[f'{case}/num_pos'])
        G = 6
        assert (g[f'{case}/pos_inds_0'][:G] >= G).any()


def test_max_iou_assigner_match_low_quality_equals_executed_reference():
    """RPN-stage assigner setting (pos=neg=0.5, min_pos_iou=0, match_low_quality=True): a GT
    that overlaps nothing claims every zero-overlap box (gt_max == 0 >= min_pos_iou), later GTs
    override earlier ones — reproduced from the executed reference."""
    g = np.load(TRAIN_GOLD)
    gi, mo = oracle.max_iou_assign(g['mlq/boxes'], g['mlq/gts'], 0.5, 0.5, 0.0, True)
    np.testing.assert_array_equal(gi, g['mlq/gt_inds'])
    np.testing.assert_array_equal(mo.view(np.uint32), g['mlq/max_overlaps'].view(np.uint32))
    assert (gi == 4).sum() > 100     # the isolated GT (index 3) grabbed the zero-overlap boxes


@pytest.mark.parametrize('case', synth.RCNN_TRAIN_CASES)
def test_python_fallback_train_prep_equals_executed_reference(case):
    """The torch fallback of the training front-end (boosting_rcnn_b200/sampling.py +
    ProbConvFCBBoxHead.get_targets + the prior lines of ProbRoIHead.forward_train; used when the
    fused kernels' limits are exceeded) on CPU tensors against the executed-reference golden."""
    import torch
    from boosting_rcnn_b200 import configs
    from boosting_rcnn_b200.roi_head import bbox2roi
    g = np.load(TRAIN_GOLD)
    torch.manual_seed(3)
    _, roi, _ = configs.build_hot_path('coco', train=True)
    gts, labels, plist = synth.rcnn_train_case(case)
    gts, labels, plist = ([torch.from_numpy(a) for a in xs] for xs in (gts, labels, plist))
    torch.manual_seed(123)
    results, priors = [], []
    for i in range(len(plist)):
        ar = roi.bbox_assigner.assign(plist[i], gts[i], None, labels[i])
        res = roi.bbox_sampler.sample(ar, plist[i], gts[i], labels[i])
        results.append(res)
        G = ar.num_gts
        pos_prior = plist[i][res.pos_inds[G:] - G, -1]
        neg_prior = 1 - plist[i][res.neg_inds - G, -1]
        priors.append(torch.cat([pos_prior.new_zeros(G), pos_prior, neg_prior]))
    lab, lw, bt, bw = roi.bbox_head.get_targets(results, gts, labels, roi.train_cfg)
    np.testing.assert_array_equal(bbox2roi([r.bboxes for r in results]).numpy(), g[f'{case}/rois'])
    np.testing.assert_array_equal(lab.numpy(), g[f'{case}/labels'])
    np.testing.assert_array_equal(lw.numpy(), g[f'{case}/label_weights'])
    np.testing.assert_array_equal(bw.numpy(), g[f'{case}/bbox_weights'])
    np.testing.assert_array_equal(torch.cat(priors).numpy(), g[f'{case}/prior'])
    np.testing.assert_allclose(bt.numpy(), g[f'{case}/bbox_targets'], rtol=1e-6, atol=1e-6)


# ---------------------------------------------------------------------------
# RPN loss path: oracle vs goldens produced by executing the reference's
# ATSSRPNHead.loss / loss_single / get_targets + its loss modules
# (tests/golden/make_golden_rpn_loss.py)
# ---------------------------------------------------------------------------
RPN_LOSS_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden',
                             'reference_golden_rpn_loss.npz')


RPN_LOSS_GOLD_CASES = synth.RPN_LOSS_CASES + ('varifocal_basic', 'varifocal_partial_valid')


def _rpn_loss_oracle(case):
    vf = case.startswith('varifocal_')
    c = synth.rpn_loss_case(case[len('varifocal_'):] if vf else case)
    gen = AnchorGenerator(strides=[8, 16, 32, 64, 128], ratios=[0.5, 1.0, 2.0],
                          octave_base_scale=4, scales_per_octave=3)
    kw = dict(cls_loss='varifocal', focal_alpha=0.75, focal_gamma=2.0) if vf else {}
    return c, oracle.rpn_loss(c['cls'], c['box'], c['iou'], c['gt_bboxes'], c['img_metas'],
                              gen.base_anchor_table().numpy(), synth.STRIDES, **kw)


def _rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max()) / \
        max(float(np.abs(b).max()), 1e-12)


@pytest.mark.parametrize('case', RPN_LOSS_GOLD_CASES)
def test_rpn_loss_oracle_equals_executed_reference(case):
    g = np.load(RPN_LOSS_GOLD)
    c, o = _rpn_loss_oracle(case)
    for k in ('loss_rpn_cls', 'loss_rpn_bbox', 'loss_rpn_iou'):
        np.testing.assert_allclose(o[k], g[f'{case}/{k}'], rtol=1e-5, atol=1e-7, err_msg=k)
    for l in range(len(c['cls'])):
        for name, key in (('grad_cls', 'grad_cls'), ('grad_box', 'grad_box'), ('grad_iou', 'grad_iou')):
            ref = g[f'{case}/{key}_{l}']
            if np.abs(ref).max() == 0:
                assert np.abs(o[name][l]).max() == 0
            else:
                assert _rel(o[name][l], ref) <= 1e-5, (name, l, _rel(o[name][l], ref))
    if case == 'isolated_gt':
        # the far-away GT (index 2 of image 0) claims every anchor it shares its zero maximum with
        assert (o['targets'][0]['assigned'] == 2).sum() > 1000
