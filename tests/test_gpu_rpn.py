"""GPU parity: brcnn_rpn_get_bboxes (C ABI) vs the CPU oracle.

Bar (BASELINE.json north_star): per-level top-k indices, candidate boxes,
NMS keep lists and final proposals BIT-EXACT (the oracle pins exp() and the
tie-break, see oracle/brcnn_oracle.c header).
"""
import numpy as np
import pytest
import torch

import synth
from boosting_rcnn_b200 import ops
from boosting_rcnn_b200.anchors import AnchorGenerator
from oracle import oracle

pytestmark = pytest.mark.gpu


def _anchor_gen(num_scales=3, ratios=(0.5, 1.0, 2.0)):
    return AnchorGenerator(strides=list(synth.STRIDES), ratios=list(ratios),
                           octave_base_scale=4, scales_per_octave=num_scales)


def _run_case(dev, batch, pad_hw, img_hw, nms_pre, max_per_img, iou_thr, seed,
              num_scales=3, ratios=(0.5, 1.0, 2.0), duplicate_frac=0.0, box_std=0.3,
              min_bbox_size=0.0, cls_std=1.5, mutate=None):
    gen = _anchor_gen(num_scales, ratios)
    A = gen.num_base_anchors[0]
    sizes = synth.featmap_sizes(*pad_hw)
    cls, box, iou = synth.rpn_outputs(batch, sizes, A, seed=seed, cls_std=cls_std,
                                      duplicate_frac=duplicate_frac, box_std=box_std)
    if mutate is not None:
        mutate(cls, iou)
    base = gen.base_anchor_table()
    img_hw_arr = np.array([img_hw] * batch, dtype=np.float32)
    p = ops.make_rpn_params(batch, sizes, synth.STRIDES, A, nms_pre, max_per_img,
                            iou_thr, min_bbox_size)
    lay = ops.rpn_workspace_layout(p)
    t = lambda arrs: [torch.from_numpy(a).to(dev) for a in arrs]
    props, num, ws = ops.rpn_get_bboxes(p, t(cls), t(box), t(iou), base.to(dev),
                                        torch.from_numpy(img_hw_arr).to(dev),
                                        return_workspace=True)
    torch.cuda.synchronize()
    props, num, ws = props.cpu().numpy(), num.cpu().numpy(), ws.cpu().numpy()
    L, Kc = len(sizes), int(lay.cand_cap)
    cand_boxes = ws[lay.cand_boxes:lay.cand_boxes + batch * L * Kc * 16].view(np.float32).reshape(batch, L, Kc, 4)
    cand_key = ws[lay.cand_key:lay.cand_key + batch * L * Kc * 8].view(np.uint64).reshape(batch, L, Kc)
    cand_count = ws[lay.cand_count:lay.cand_count + batch * L * 4].view(np.int32).reshape(batch, L)
    level_base = np.cumsum([0] + [h * w * A for (h, w) in sizes])
    for b in range(batch):
        ref, dbg = oracle.rpn_get_bboxes_single(
            [c[b] for c in cls], [c[b] for c in box], [c[b] for c in iou], base.numpy(),
            synth.STRIDES, img_hw, nms_pre, max_per_img, iou_thr, min_bbox_size,
            debug=True)
        off = 0
        for l in range(L):
            k = int(dbg['cand_n'][l])
            assert cand_count[b, l] == k
            keys = cand_key[b, l, :k]
            idx = (0xFFFFFFFF - (keys & np.uint64(0xFFFFFFFF))).astype(np.int64) - level_base[l]
            sc = (keys >> np.uint64(32)).astype(np.uint32).view(np.float32)
            ref_idx = dbg['topk_idx'][off:off + k].astype(np.int64)
            ref_sc = dbg['cand_scores'][off:off + k]
            n_l = sizes[l][0] * sizes[l][1] * A
            if n_l > nms_pre > 0:
                # sorted levels: rank-for-rank identical
                np.testing.assert_array_equal(idx, ref_idx, err_msg=f'top-k idx b{b} l{l}')
                np.testing.assert_array_equal(sc.view(np.uint32), ref_sc.view(np.uint32))
                np.testing.assert_array_equal(cand_boxes[b, l, :k].view(np.uint32),
                                              dbg['cand_boxes'][off:off + k].view(np.uint32))
            else:
                # the reference keeps small levels unsorted; same set, and
                # the sorted order must be (score desc, idx asc)
                order = np.lexsort((ref_idx, -ref_sc.astype(np.float64)))
                np.testing.assert_array_equal(idx, ref_idx[order])
                np.testing.assert_array_equal(cand_boxes[b, l, :k].view(np.uint32),
                                              dbg['cand_boxes'][off:off + k][order].view(np.uint32))
            off += k
        assert num[b] == ref.shape[0], f'image {b}: {num[b]} vs {ref.shape[0]} proposals'
        np.testing.assert_array_equal(props[b, :num[b]].view(np.uint32), ref.view(np.uint32),
                                      err_msg=f'proposals image {b}')
        assert not props[b, num[b]:].any()
    return props, num


def test_rpn_small_unique(cuda):
    _run_case(cuda, batch=2, pad_hw=(256, 320), img_hw=(250, 317), nms_pre=300,
              max_per_img=100, iou_thr=0.7, seed=1)


def test_rpn_small_duplicates(cuda):
    # heavy exact ties in the scores: tie-break (score desc, index asc)
    _run_case(cuda, batch=3, pad_hw=(192, 256), img_hw=(192, 250), nms_pre=200,
              max_per_img=64, iou_thr=0.7, seed=2, duplicate_frac=0.9)


def test_rpn_concentrated_scores_slow_path(cuda):
    # degenerate score distributions over-populate the threshold bin of the value
    # histogram (> smem candidate slots): the top-k kernel must fall back to the exact
    # radix selection.  (a) every logit identical, (b) a narrow spread, (c) one level
    # constant and the others random
    def const(cls, iou):
        for c, u in zip(cls, iou):
            c[...] = 0.25
            u[...] = -0.5
    _run_case(cuda, batch=2, pad_hw=(256, 320), img_hw=(250, 317), nms_pre=300,
              max_per_img=100, iou_thr=0.7, seed=21, mutate=const)
    _run_case(cuda, batch=1, pad_hw=(512, 640), img_hw=(500, 600), nms_pre=1000,
              max_per_img=300, iou_thr=0.7, seed=22, cls_std=1e-4)

    def level0_const(cls, iou):
        cls[0][...] = 1.0
        iou[0][...] = 1.0
    _run_case(cuda, batch=2, pad_hw=(256, 320), img_hw=(250, 317), nms_pre=300,
              max_per_img=100, iou_thr=0.7, seed=23, mutate=level0_const)


def test_rpn_scores_clustered_on_histogram_bin_edges(cuda):
    """Candidate selection runs on approximate scores (2048 value bins), ranking on the
    pinned ones: pile the scores of every level onto bin edges d/2048 (+- a few ulp, where the
    approximate and the exact score fall into different bins) so that the k-th score sits in
    such a pile; indices, scores and proposals must still equal the oracle bit for bit."""
    def on_edges(cls, iou):
        rng = np.random.RandomState(77)
        for c, u in zip(cls, iou):
            u[...] = 30.0                                  # sigmoid(iou) == 1: s = sqrt(sigmoid(cls))
            edges = rng.randint(900, 1100, size=c.shape).astype(np.float64) / 2048.0
            p = edges * edges                              # sigmoid(cls) that lands ON the edge
            x = np.log(p / (1.0 - p))
            x32 = x.astype(np.float32)
            # +-3 ulp jitter around the edge logit
            jit = rng.randint(-3, 4, size=c.shape).astype(np.int32)
            c[...] = (x32.view(np.int32) + jit).view(np.float32)
    _run_case(cuda, batch=2, pad_hw=(256, 320), img_hw=(250, 317), nms_pre=300,
              max_per_img=100, iou_thr=0.7, seed=31, mutate=on_edges)
    _run_case(cuda, batch=1, pad_hw=(512, 640), img_hw=(500, 600), nms_pre=1000,
              max_per_img=300, iou_thr=0.7, seed=32, mutate=on_edges)


def test_rpn_extreme_logits(cuda):
    """Saturated and infinite logits: the approximate pre-selection score and the pinned score
    must agree on 0 / 1 and on everything in between (exp overflow, denormal sigmoids)."""
    def extreme(cls, iou):
        rng = np.random.RandomState(91)
        vals = np.array([np.inf, -np.inf, 95.0, -95.0, 88.5, -88.5, 40.0, -40.0, 17.0, -17.0],
                        dtype=np.float32)
        for c, u in zip(cls, iou):
            m = rng.rand(*c.shape) < 0.03
            c[m] = vals[rng.randint(0, len(vals), int(m.sum()))]
            m = rng.rand(*u.shape) < 0.03
            u[m] = vals[rng.randint(0, len(vals), int(m.sum()))]
    _run_case(cuda, batch=2, pad_hw=(256, 320), img_hw=(250, 317), nms_pre=300,
              max_per_img=100, iou_thr=0.7, seed=41, mutate=extreme)
    _run_case(cuda, batch=1, pad_hw=(512, 640), img_hw=(500, 600), nms_pre=1000,
              max_per_img=300, iou_thr=0.7, seed=42, mutate=extreme)


def test_rpn_single_anchor_voc_like(cuda):
    _run_case(cuda, batch=2, pad_hw=(608, 1024), img_hw=(600, 1000), nms_pre=1000,
              max_per_img=256, iou_thr=0.7, seed=3, num_scales=1, ratios=(1.0,))


def test_rpn_utdac_test_cfg_full_size(cuda):
    # configs[1] shapes: 1333x800 padded to 1344x800, nms_pre 1000, max 256
    _run_case(cuda, batch=2, pad_hw=(800, 1344), img_hw=(800, 1333), nms_pre=1000,
              max_per_img=256, iou_thr=0.7, seed=4)


def test_rpn_train_cfg_full_size(cuda):
    # configs[2] proposal settings: nms_pre 4000, max 2000 (split path, K=15150)
    _run_case(cuda, batch=1, pad_hw=(800, 1344), img_hw=(800, 1333), nms_pre=4000,
              max_per_img=2000, iou_thr=0.7, seed=5)


def test_rpn_degenerate_boxes_filtered(cuda):
    # large deltas push many boxes outside -> clipped to zero width (w<=0 filter)
    _run_case(cuda, batch=2, pad_hw=(128, 160), img_hw=(100, 150), nms_pre=500,
              max_per_img=300, iou_thr=0.7, seed=6, box_std=3.0)


def test_rpn_low_iou_threshold(cuda):
    _run_case(cuda, batch=1, pad_hw=(256, 256), img_hw=(256, 256), nms_pre=1000,
              max_per_img=1000, iou_thr=0.3, seed=7)


def test_rpn_determinism(cuda):
    a = _run_case(cuda, batch=2, pad_hw=(256, 320), img_hw=(250, 317), nms_pre=300,
                  max_per_img=100, iou_thr=0.7, seed=11, duplicate_frac=0.5)
    b = _run_case(cuda, batch=2, pad_hw=(256, 320), img_hw=(250, 317), nms_pre=300,
                  max_per_img=100, iou_thr=0.7, seed=11, duplicate_frac=0.5)
    np.testing.assert_array_equal(a[0].view(np.uint32), b[0].view(np.uint32))


def test_rpn_image_kernel_equals_segment_path(cuda):
    """BRCNN_RPN_NMS=segments (per-level NMS to completion + merge) vs image1 (one CTA
    per image, global score order, early stop) vs the cluster version with 8, 4 and 2 CTAs per
    image (the launcher drops to 4 when a batch's 8-CTA clusters cannot all be resident):
    identical proposals."""
    import os
    import subprocess
    import sys
    import tempfile
    here = os.path.dirname(os.path.abspath(__file__))
    code = '''
import sys, numpy as np, torch
sys.path[:0] = [%r, %r]
import synth
from boosting_rcnn_b200 import ops
from boosting_rcnn_b200.anchors import AnchorGenerator
gen = AnchorGenerator(strides=list(synth.STRIDES), ratios=[0.5, 1.0, 2.0], octave_base_scale=4,
                      scales_per_octave=3)
sizes = synth.featmap_sizes(800, 1344)
cls, box, iou = synth.rpn_outputs(2, sizes, 9, seed=77, duplicate_frac=0.2)
t = lambda arrs: [torch.from_numpy(a).cuda() for a in arrs]
hw = torch.tensor([[800., 1333.]] * 2).cuda()
res = []
for nms_pre, mx in ((1000, 256), (4000, 2000), (300, 3000)):
    p = ops.make_rpn_params(2, sizes, synth.STRIDES, 9, nms_pre, mx, 0.7, 0.0)
    pr, n = ops.rpn_get_bboxes(p, t(cls), t(box), t(iou), gen.base_anchor_table().cuda(), hw)
    res += [pr.cpu().numpy().reshape(-1), n.cpu().numpy().astype(np.float32)]
np.save(sys.argv[1], np.concatenate(res))
''' % (os.path.dirname(here), here)
    outs = []
    with tempfile.TemporaryDirectory() as d:
        for mode, cs in (('segments', ''), ('image1', ''), ('cluster', '8'), ('cluster', '4'),
                         ('cluster', '2')):
            path = os.path.join(d, mode + cs + '.npy')
            env = dict(os.environ, BRCNN_RPN_NMS=mode)
            if cs:
                env['BRCNN_RNI_CS'] = cs
            subprocess.run([sys.executable, '-c', code, path], check=True, env=env)
            outs.append(np.load(path))
    for o in outs[1:]:
        np.testing.assert_array_equal(outs[0].view(np.uint32), o.view(np.uint32))
