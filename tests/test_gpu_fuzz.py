"""GPU: seeded randomised small configurations of every entry point against the
oracle — odd level counts, single-anchor heads, tiny maps, ragged images, extreme
thresholds — the shapes the hand-picked cases do not enumerate.  Same bars as
the dedicated files: bit-exact indices / boxes / keep lists, <= 1e-5 features."""
import numpy as np
import pytest
import torch

import synth
from boosting_rcnn_b200 import ops
from boosting_rcnn_b200.anchors import AnchorGenerator
from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('seed', range(8))
def test_fuzz_rpn_get_bboxes(cuda, seed):
    rng = np.random.RandomState(1000 + seed)
    L = int(rng.randint(1, 7))
    strides = [4 * 2 ** l for l in range(L)]
    nsc = int(rng.choice([1, 2, 3]))
    ratios = [[1.0], [0.5, 1.0, 2.0], [0.5, 2.0]][int(rng.randint(0, 3))]
    gen = AnchorGenerator(strides=strides, ratios=ratios, octave_base_scale=4, scales_per_octave=nsc)
    A = gen.num_base_anchors[0]
    B = int(rng.randint(1, 4))
    pad_h, pad_w = int(rng.randint(2, 12)) * 32, int(rng.randint(2, 12)) * 32
    img_hw = (pad_h - int(rng.randint(0, 31)), pad_w - int(rng.randint(0, 31)))
    sizes = [(-(-pad_h // s), -(-pad_w // s)) for s in strides]
    nms_pre = int(rng.choice([1, 7, 64, 300, 1000, 5000]))
    max_per_img = int(rng.choice([1, 3, 100, 700]))
    iou_thr = float(rng.choice([0.0, 0.3, 0.7, 0.95]))
    cls, box, iou = synth.rpn_outputs(B, sizes, A, seed=seed, cls_std=float(rng.choice([0.01, 1.5, 6.0])),
                                      box_std=float(rng.choice([0.1, 0.5, 2.5])),
                                      duplicate_frac=float(rng.choice([0.0, 0.5])))
    p = ops.make_rpn_params(B, sizes, strides, A, nms_pre, max_per_img, iou_thr, 0.0)
    t = lambda arrs: [torch.from_numpy(a).to(cuda) for a in arrs]
    hw = torch.tensor([img_hw] * B, dtype=torch.float32, device=cuda)
    props, num = ops.rpn_get_bboxes(p, t(cls), t(box), t(iou), gen.base_anchor_table().to(cuda), hw)
    props, num = props.cpu().numpy(), num.cpu().numpy()
    for b in range(B):
        ref = oracle.rpn_get_bboxes_single([c[b] for c in cls], [c[b] for c in box],
                                           [c[b] for c in iou], gen.base_anchor_table().numpy(),
                                           strides, img_hw, nms_pre, max_per_img, iou_thr, 0.0)
        assert num[b] == ref.shape[0], (seed, b, num[b], ref.shape[0])
        np.testing.assert_array_equal(props[b, :num[b]].view(np.uint32), ref.view(np.uint32))
        assert not props[b, num[b]:].any()


@pytest.mark.parametrize('seed', range(8))
def test_fuzz_nms_operators(cuda, seed):
    rng = np.random.RandomState(2000 + seed)
    K = int(rng.choice([1, 2, 63, 64, 65, 129, 777, 2500]))
    nid = int(rng.choice([1, 2, 5, 8, 9, 30]))
    boxes = synth.random_boxes(K, 600, 900, seed=seed, clustered=bool(rng.randint(0, 2)))
    scores = (np.round(rng.rand(K) * 200) / 200 - 0.1).astype(np.float32)
    ids = rng.randint(0, nid, K)
    thr = float(rng.choice([0.1, 0.5, 0.7]))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    dets, keep = ops.batched_nms(t(boxes), t(scores), t(ids), dict(type='nms', iou_threshold=thr))
    rd, rk = oracle.batched_nms(boxes, scores, ids, thr)
    np.testing.assert_array_equal(keep.cpu().numpy(), rk)
    np.testing.assert_array_equal(dets.cpu().numpy().view(np.uint32), rd.view(np.uint32))
    off = int(rng.randint(0, 2))
    dets, keep = ops.nms(t(boxes), t(scores), thr, offset=off)
    np.testing.assert_array_equal(keep.cpu().numpy(), oracle.nms_cpu(boxes, scores, thr, offset=off))


@pytest.mark.parametrize('seed', range(6))
def test_fuzz_roi_extract_forward_backward(cuda, seed):
    rng = np.random.RandomState(3000 + seed)
    L = int(rng.randint(1, 6))
    strides = [8 * 2 ** l for l in range(L)]
    B, C = int(rng.randint(1, 4)), int(rng.choice([4, 12, 64, 132, 260]))
    pad_h, pad_w = int(rng.randint(2, 10)) * 32, int(rng.randint(2, 10)) * 32
    sizes = [(-(-pad_h // s), -(-pad_w // s)) for s in strides]
    scales = [1.0 / s for s in strides]
    osz = [(7, 7), (7, 7), (3, 5), (1, 1), (14, 14)][int(rng.randint(0, 5))]
    sr = int(rng.choice([0, 0, 2]))
    feats = synth.fpn_feats(B, C, sizes, seed=seed)
    rois = synth.random_rois(B, int(rng.randint(1, 60)), pad_h, pad_w, seed=seed + 1,
                             clustered=bool(rng.randint(0, 2)))
    rois = np.concatenate([rois, [[-1, 0, 0, 0, 0], [0, 3, 3, 3, 3]]]).astype(np.float32)
    tf = [torch.from_numpy(f).to(cuda).requires_grad_(True) for f in feats]
    out, lv = ops.roi_extract(tf, torch.from_numpy(rois).to(cuda), scales, osz, sr, True, 56,
                              return_levels=True)
    ref, rlv = oracle.roi_extract_forward(feats, rois, scales, osz, sr)
    np.testing.assert_array_equal(lv.cpu().numpy().astype(np.int64), rlv)
    tol = lambda r: 1e-5 * max(float(np.abs(r).max()), 1e-6)
    assert np.abs(out.detach().cpu().numpy() - ref).max() <= tol(ref)
    g = rng.normal(0, 1, ref.shape).astype(np.float32)
    out.backward(torch.from_numpy(g).to(cuda))
    refg = oracle.roi_extract_backward(g, rois, [f.shape for f in feats], scales, sr)
    for x, r in zip(tf, refg):
        assert np.abs(x.grad.cpu().numpy() - r).max() <= tol(r) * 4


@pytest.mark.parametrize('seed', range(6))
def test_fuzz_rcnn_get_bboxes(cuda, seed):
    from test_gpu_rcnn import _case
    rng = np.random.RandomState(4000 + seed)
    B = int(rng.randint(1, 5))
    Rc = int(rng.choice([1, 7, 64, 200, 513]))
    C = int(rng.choice([1, 2, 4, 20, 33, 80]))
    counts = [int(rng.randint(0, Rc + 1)) for _ in range(B)]
    _case(cuda, B, Rc, C, (int(rng.randint(200, 900)), int(rng.randint(200, 1400))), seed=seed,
          score_thr=float(rng.choice([0.0, 0.01, 0.05, 0.3])), iou_thr=float(rng.choice([0.3, 0.5, 0.7])),
          max_per_img=int(rng.choice([1, 10, 100, 300])), rescale=bool(rng.randint(0, 2)),
          counts=counts, logit_std=float(rng.choice([0.5, 2.0, 5.0])))


@pytest.mark.parametrize('seed', range(6))
def test_fuzz_boost_loss(cuda, seed):
    from test_gpu_loss import _inputs, _rel, _torch_reference
    rng = np.random.RandomState(5000 + seed)
    N = int(rng.choice([1, 2, 31, 257, 1024, 3000]))
    C = int(rng.choice([1, 4, 20, 80, 150]))
    gamma = float(rng.choice([0.5, 1.0, 0.3]))
    cls, labels, prior, bp, bt, bw = _inputs(N, C, seed, pos_frac=float(rng.choice([0.0, 0.25, 1.0])))
    prior = np.minimum(prior, 0.999).astype(np.float32)
    tc = torch.from_numpy(cls).to(cuda).requires_grad_(True)
    tb = torch.from_numpy(bp).to(cuda).requires_grad_(True)
    t = lambda a: torch.from_numpy(a).to(cuda)
    loss_cls, loss_bbox, acc, _ = ops.boost_loss(tc, tb, t(labels), None, t(prior), t(bt), t(bw), C,
                                                 False, gamma, 0.0, 2.0, 2.0, False)
    (loss_cls + loss_bbox).backward()
    rc, rb, ra, gc, gb = _torch_reference(cls, labels, prior, bp, bt, bw, C, gamma, 2.0, 2.0)
    assert abs(loss_cls.item() - rc) <= 1e-5 * max(abs(rc), 1e-6)
    assert abs(loss_bbox.item() - rb) <= 1e-5 * max(abs(rb), 1e-6) + 1e-9
    assert abs(acc.item() - ra) <= 1e-4
    assert _rel(tc.grad.cpu().numpy(), gc) <= 1e-5
    if np.abs(gb).max() > 0:
        assert _rel(tb.grad.cpu().numpy(), gb) <= 1e-5
    else:
        assert tb.grad.abs().max().item() == 0
