"""GPU: size-independent properties at BASELINE.json's FULL sizes (configs[1]:
16 images of 1333x800, UTDAC test cfg; configs[4]: 4000 pre-NMS / 2000 post-NMS),
where the CPU oracle would take minutes: sortedness, box range, NMS idempotence,
agreement of independent CUDA code paths, RoIAlign linearity / partition of unity
/ adjointness of forward and backward."""
import numpy as np
import pytest
import torch

import synth
from boosting_rcnn_b200 import configs, ops
from boosting_rcnn_b200.anchors import AnchorGenerator

pytestmark = pytest.mark.gpu

PAD_HW, IMG_HW = (800, 1344), (800, 1333)


def _rpn_inputs(dev, B, seed):
    sizes = synth.featmap_sizes(*PAD_HW)
    g = torch.Generator().manual_seed(seed)
    mk = lambda ch, std: [(torch.randn(B, ch, h, w, generator=g) * std).to(dev) for h, w in sizes]
    return sizes, mk(9, 1.5), mk(36, 0.3), mk(9, 1.5)


@pytest.mark.parametrize('B,nms_pre,max_per_img', [(16, 1000, 256), (4, 4000, 2000)])
def test_rpn_full_size_properties_and_generic_nms_agreement(cuda, B, nms_pre, max_per_img):
    gen = AnchorGenerator(strides=list(synth.STRIDES), ratios=[0.5, 1.0, 2.0],
                          octave_base_scale=4, scales_per_octave=3)
    sizes, cls, box, iou = _rpn_inputs(cuda, B, seed=100 + B)
    hw = torch.tensor([IMG_HW] * B, dtype=torch.float32, device=cuda)
    p = ops.make_rpn_params(B, sizes, synth.STRIDES, 9, nms_pre, max_per_img, 0.7, 0.0)
    lay = ops.rpn_workspace_layout(p)
    props, num, ws = ops.rpn_get_bboxes(p, cls, box, iou, gen.base_anchor_table().to(cuda), hw,
                                        return_workspace=True)
    L, Kc = 5, int(lay.cand_cap)
    n = num.cpu().numpy()
    assert (n == max_per_img).all()          # dense random anchors always fill the quota
    P = props.cpu().numpy()
    for b in range(B):
        s = P[b, :n[b], 4]
        assert (np.diff(s) <= 0).all(), 'scores not descending'
        bx = P[b, :n[b], :4]
        assert (bx[:, 0::2] >= 0).all() and (bx[:, 0::2] <= IMG_HW[1]).all()
        assert (bx[:, 1::2] >= 0).all() and (bx[:, 1::2] <= IMG_HW[0]).all()
        assert ((bx[:, 2] - bx[:, 0]) > 0).all() and ((bx[:, 3] - bx[:, 1]) > 0).all()
        assert not P[b, n[b]:].any()
    # second path: the generic mmcv-style batched_nms operator on the very same candidates
    # (its own id-rank sort + list segmentation + keep-index output, all keeps instead of
    # an early stop) must give the same first max_per_img rows as brcnn_rpn_get_bboxes;
    # tests/test_gpu_nms.py and test_gpu_rpn.py pin the clustered walk against the
    # single-segment and per-level kernels
    wsb = ws.cpu().numpy()
    cb = wsb[lay.cand_boxes:lay.cand_boxes + B * L * Kc * 16].view(np.float32).reshape(B, L, Kc, 4)
    ck = wsb[lay.cand_key:lay.cand_key + B * L * Kc * 8].view(np.uint64).reshape(B, L, Kc)
    cc = wsb[lay.cand_count:lay.cand_count + B * L * 4].view(np.int32).reshape(B, L)
    for b in (0, B - 1):
        boxes = np.concatenate([cb[b, l, :cc[b, l]] for l in range(L)])
        keys = np.concatenate([ck[b, l, :cc[b, l]] for l in range(L)])
        scores = (keys >> np.uint64(32)).astype(np.uint32).view(np.float32)
        ids = np.concatenate([np.full(cc[b, l], l) for l in range(L)])
        # the reference drops w <= 0 / h <= 0 boxes before batched_nms (atss_rpn_head.py:747-754)
        ok = ((boxes[:, 2] - boxes[:, 0]) > 0) & ((boxes[:, 3] - boxes[:, 1]) > 0)
        boxes, keys, scores, ids = boxes[ok], keys[ok], scores[ok], ids[ok]
        # the operator breaks ties by input order; feed it in the kernel's key order
        order = np.argsort(-keys.astype(np.float64), kind='stable')
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
        dets, _ = ops.batched_nms(t(boxes[order]), t(scores[order]), t(ids[order]),
                                  dict(type='nms', iou_threshold=0.7))
        np.testing.assert_array_equal(dets[:max_per_img].cpu().numpy().view(np.uint32),
                                      P[b, :n[b]].view(np.uint32))


def _roi_setup(dev, B, n_per_img, C=256, seed=5):
    sizes = synth.featmap_sizes(*PAD_HW)
    g = torch.Generator().manual_seed(seed)
    feats = [torch.randn(B, C, h, w, generator=g).to(dev).contiguous(memory_format=torch.channels_last)
             for h, w in sizes]
    rois = torch.from_numpy(synth.random_rois(B, n_per_img, IMG_HW[0], IMG_HW[1], seed=seed + 1)).to(dev)
    return sizes, feats, rois, [1.0 / s for s in synth.STRIDES]


@pytest.mark.parametrize('B,n', [(16, 256), (4, 2000)])
def test_roi_align_full_size_linearity_unity_adjoint(cuda, B, n):
    sizes, f1, rois, scales = _roi_setup(cuda, B, n)
    _, f2, _, _ = _roi_setup(cuda, B, n, seed=9)
    a, b = 0.75, -1.5
    o1 = ops.roi_extract(f1, rois, scales, 7)
    o2 = ops.roi_extract(f2, rois, scales, 7)
    o12 = ops.roi_extract([a * x + b * y for x, y in zip(f1, f2)], rois, scales, 7)
    ref = a * o1 + b * o2
    assert (o12 - ref).abs().max().item() <= 1e-5 * ref.abs().max().item() * 4
    # partition of unity: a constant map pools to the constant for RoIs inside the image
    ones = [torch.full_like(x, 3.25) for x in f1]
    oc = ops.roi_extract(ones, rois, scales, 7)
    r = rois.cpu().numpy()
    inside = (r[:, 1] >= 1) & (r[:, 2] >= 1) & (r[:, 3] <= IMG_HW[1] - 1) & (r[:, 4] <= IMG_HW[0] - 1)
    inside = torch.from_numpy(inside).to(cuda)
    assert inside.sum().item() > n  # the sweep has plenty of interior RoIs
    assert (oc[inside] - 3.25).abs().max().item() <= 3.25e-5
    # adjointness of forward and backward: <A f, g> == <f, A^T g>
    fr = [x.detach().clone().requires_grad_(True) for x in f1]
    out = ops.roi_extract(fr, rois, scales, 7)
    g = torch.randn(out.shape, generator=torch.Generator().manual_seed(3)).to(cuda)
    lhs = (out.double() * g.double()).sum().item()
    out.backward(g)
    rhs = sum((x.double() * x.grad.double()).sum().item() for x in fr)
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), abs(rhs), 1.0) * 10, (lhs, rhs)
    # every pyramid level received a gradient tensor, deterministic across runs
    out2 = ops.roi_extract(fr, rois, scales, 7)
    grads = torch.autograd.grad(out2, fr, g)
    for x, g2 in zip(fr, grads):
        assert torch.equal(x.grad, g2)


def test_detections_full_size_properties_and_nms_idempotence(cuda):
    torch.manual_seed(0)
    rpn, roi, model = configs.build_hot_path('utdac')
    rpn, roi = rpn.to(cuda).eval(), roi.to(cuda).eval()
    B = 16
    sizes, cls, box, iou = _rpn_inputs(cuda, B, seed=321)
    g = torch.Generator().manual_seed(8)
    feats = [torch.randn(B, 256, h, w, generator=g).to(cuda) for h, w in sizes]
    metas = [dict(img_shape=(*IMG_HW, 3), pad_shape=(*PAD_HW, 3),
                  scale_factor=np.array([1.6662, 1.6667, 1.6662, 1.6667], dtype=np.float32))
             for _ in range(B)]
    cfg = model['test_cfg']['rcnn']
    with torch.no_grad():
        props = rpn.get_bboxes_padded(cls, box, iou, metas)
        det, lab, num = roi.simple_test_bboxes_padded(feats, metas, props, cfg, rescale=True)
    n = num.cpu().numpy()
    assert (n <= cfg['max_per_img']).all() and n.sum() > 0
    for b in range(B):
        d, l = det[b, :n[b]], lab[b, :n[b]]
        s = d[:, 4].cpu().numpy()
        assert (np.diff(s) <= 0).all() and (s > cfg['score_thr']).all()
        assert ((l >= 0) & (l < 4)).all()
        # class-wise NMS of the detections themselves suppresses nothing
        if n[b]:
            kept, _ = ops.batched_nms(d[:, :4].contiguous(), d[:, 4].contiguous(), l,
                                      dict(type='nms', iou_threshold=cfg['nms']['iou_threshold']))
            assert kept.size(0) == n[b]
        assert not det[b, n[b]:].any() and (lab[b, n[b]:] == -1).all()


def test_level_map_monotone_in_scale(cuda):
    s = torch.logspace(0, 3.2, 4000, device=cuda)
    rois = torch.stack([torch.zeros_like(s), torch.zeros_like(s), torch.zeros_like(s), s, s], 1)
    lv = ops.map_roi_levels(rois, 5, 56).cpu().numpy()
    assert (np.diff(lv) >= 0).all() and lv.min() == 0 and lv.max() == 4
