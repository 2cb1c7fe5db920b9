"""GPU parity: probabilistic score fusion + per-class decode + class-wise NMS
(brcnn_rcnn_get_bboxes) vs the oracle.  Bar: fused scores, decoded boxes, keep
lists, labels and final detections bit-exact."""
import numpy as np
import pytest
import torch

import synth
from boosting_rcnn_b200 import ops
from oracle import oracle

pytestmark = pytest.mark.gpu


def _case(dev, batch, Rc, C, img_hw, seed, score_thr=0.05, iou_thr=0.7, max_per_img=100,
          rescale=False, counts=None, logit_std=2.0, prior_lo=0.0):
    rng = np.random.RandomState(seed)
    rois = synth.random_rois(batch, Rc, img_hw[0], img_hw[1], seed=seed, clustered=True)
    counts = np.array(counts if counts is not None else [Rc] * batch, dtype=np.int32)
    cls = rng.normal(0, logit_std, (batch * Rc, C + 1)).astype(np.float32)
    bp = rng.normal(0, 1.0, (batch * Rc, 4 * C)).astype(np.float32)
    prior = (prior_lo + (1 - prior_lo) * rng.rand(batch * Rc)).astype(np.float32)
    hw = np.array([img_hw] * batch, dtype=np.float32)
    sf = np.array([[1.666, 1.5, 1.666, 1.5]] * batch, dtype=np.float32)
    p = ops.make_rcnn_params(batch, Rc, C, score_thr, iou_thr, max_per_img, rescale=rescale)
    lay = ops.rcnn_workspace_layout(p)
    t = lambda a: torch.from_numpy(a).to(dev)
    det, lab, num, ws = ops.rcnn_get_bboxes(p, t(rois), t(prior), t(counts), t(cls), t(bp), t(hw),
                                            t(sf) if rescale else None, return_workspace=True)
    torch.cuda.synchronize()
    det, lab, num, ws = det.cpu().numpy(), lab.cpu().numpy(), num.cpu().numpy(), ws.cpu().numpy()
    scores_gpu = ws[lay.scores:lay.scores + batch * Rc * (C + 1) * 4].view(np.float32).reshape(batch, Rc, C + 1)
    boxes_gpu = ws[lay.bboxes:lay.bboxes + batch * Rc * C * 16].view(np.float32).reshape(batch, Rc, C * 4)
    for b in range(batch):
        n = counts[b]
        sl = slice(b * Rc, b * Rc + n)
        fused = oracle.fuse_scores(cls[sl], prior[sl])
        np.testing.assert_array_equal(scores_gpu[b, :n].view(np.uint32), fused.view(np.uint32))
        rd, rl, dbg = oracle.rcnn_get_bboxes_single(rois[sl], fused, bp[sl], img_hw, sf[b], C, score_thr,
                                                    iou_thr, max_per_img, rescale=rescale, debug=True)
        np.testing.assert_array_equal(boxes_gpu[b, :n].view(np.uint32), dbg['decoded'].view(np.uint32))
        assert num[b] == rd.shape[0], f'image {b}: {num[b]} vs {rd.shape[0]}'
        np.testing.assert_array_equal(lab[b, :num[b]], rl)
        np.testing.assert_array_equal(det[b, :num[b]].view(np.uint32), rd.view(np.uint32))
    return det, lab, num


def test_rcnn_utdac(cuda):
    _case(cuda, 4, 256, 4, (800, 1333), seed=0)


def test_rcnn_coco_classes_split_path(cuda):
    # 256 x 80 = 20480 cells; low threshold -> > 10000 candidates (mmcv split path)
    _case(cuda, 2, 256, 80, (800, 1333), seed=1, score_thr=0.001, iou_thr=0.5, logit_std=1.0,
          prior_lo=0.5)


def test_rcnn_voc_1000_rois_rescale(cuda):
    _case(cuda, 2, 1000, 20, (600, 1000), seed=2, iou_thr=0.5, rescale=True)


def test_rcnn_ragged_and_empty_images(cuda):
    _case(cuda, 4, 128, 4, (400, 600), seed=3, counts=[128, 0, 5, 77])


def test_rcnn_nothing_above_threshold(cuda):
    det, lab, num = _case(cuda, 2, 64, 4, (400, 600), seed=4, score_thr=0.999)
    assert (num == 0).all()
