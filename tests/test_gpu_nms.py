"""GPU parity: mmcv-style nms / batched_nms operators vs the oracle's restated
mmcv nms_cpu + batched_nms (keep lists bit-exact)."""
import numpy as np
import pytest
import torch

import synth
from boosting_rcnn_b200 import ops
from oracle import oracle

pytestmark = pytest.mark.gpu


def _scores(n, seed, dup=False):
    rng = np.random.RandomState(seed)
    s = rng.rand(n).astype(np.float32)
    if dup:
        s = np.round(s * 50) / 50
    return s.astype(np.float32)


@pytest.mark.parametrize('n,clustered,dup', [(1, False, False), (63, False, False),
                                             (64, True, False), (65, True, True),
                                             (1000, True, True), (4693, False, False),
                                             (5000, True, True)])
def test_nms_matches_oracle(cuda, n, clustered, dup):
    boxes = synth.random_boxes(n, 800, 1333, seed=n, clustered=clustered)
    scores = _scores(n, n + 1, dup)
    dets, keep = ops.nms(torch.from_numpy(boxes).to(cuda), torch.from_numpy(scores).to(cuda), 0.7)
    ref = oracle.nms_cpu(boxes, scores, 0.7)
    np.testing.assert_array_equal(keep.cpu().numpy(), ref)
    np.testing.assert_array_equal(dets.cpu().numpy()[:, :4], boxes[ref])
    np.testing.assert_array_equal(dets.cpu().numpy()[:, 4], scores[ref])


def test_nms_offset_one_and_thresholds(cuda):
    boxes = synth.random_boxes(700, 600, 1000, seed=3, clustered=True)
    scores = _scores(700, 4)
    for thr in (0.3, 0.5, 0.7):
        for off in (0, 1):
            _, keep = ops.nms(torch.from_numpy(boxes).to(cuda), torch.from_numpy(scores).to(cuda), thr, offset=off)
            np.testing.assert_array_equal(keep.cpu().numpy(), oracle.nms_cpu(boxes, scores, thr, off))


def test_nms_score_threshold_and_max_num(cuda):
    boxes = synth.random_boxes(500, 600, 1000, seed=5, clustered=True)
    scores = _scores(500, 6)
    dets, keep = ops.nms(torch.from_numpy(boxes).to(cuda), torch.from_numpy(scores).to(cuda), 0.5,
                         score_threshold=0.3, max_num=20)
    m = scores > 0.3
    ref = np.nonzero(m)[0][oracle.nms_cpu(boxes[m], scores[m], 0.5)][:20]
    np.testing.assert_array_equal(keep.cpu().numpy(), ref)


@pytest.mark.parametrize('n,nid', [(300, 5), (4693, 5), (12000, 80)])
def test_batched_nms_matches_oracle(cuda, n, nid):
    # n=12000 takes mmcv's split path (split_thr=10000) in the oracle
    boxes = synth.random_boxes(n, 800, 1333, seed=n + 7, clustered=True)
    scores = _scores(n, n + 8, dup=True)
    ids = np.random.RandomState(n).randint(0, nid, n).astype(np.int64)
    dets, keep = ops.batched_nms(torch.from_numpy(boxes).to(cuda), torch.from_numpy(scores).to(cuda),
                                 torch.from_numpy(ids).to(cuda), dict(type='nms', iou_threshold=0.7))
    rdets, rkeep = oracle.batched_nms(boxes, scores, ids, 0.7)
    np.testing.assert_array_equal(keep.cpu().numpy(), rkeep)
    np.testing.assert_array_equal(dets.cpu().numpy().view(np.uint32), rdets.view(np.uint32))


def test_batched_nms_class_agnostic_and_empty(cuda):
    boxes = synth.random_boxes(200, 300, 300, seed=1, clustered=True)
    scores = _scores(200, 2)
    ids = np.random.RandomState(0).randint(0, 3, 200).astype(np.int64)
    _, keep = ops.batched_nms(torch.from_numpy(boxes).to(cuda), torch.from_numpy(scores).to(cuda),
                              torch.from_numpy(ids).to(cuda),
                              dict(type='nms', iou_threshold=0.5, class_agnostic=True))
    np.testing.assert_array_equal(keep.cpu().numpy(), oracle.nms_cpu(boxes, scores, 0.5))
    dets, keep = ops.batched_nms(torch.zeros((0, 4), device=cuda), torch.zeros((0,), device=cuda),
                                 torch.zeros((0,), dtype=torch.long, device=cuda),
                                 dict(type='nms', iou_threshold=0.5))
    assert dets.shape == (0, 5) and keep.shape == (0,)


def test_nms_rejects_cpu_tensors():
    with pytest.raises(RuntimeError):
        ops.nms(torch.zeros((4, 4)), torch.zeros((4,)), 0.5)


def test_nms_operator_cluster_and_segment_paths_agree(cuda):
    """BRCNN_NMS_OP=old (one offset-box segment) vs the default clustered list walk
    of brcnn_batched_nms (<= 8 ids) and the per-id segmented path (> 8 ids, K <= 8192):
    identical keep lists for plain nms, offset=1, few and many ids."""
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
res = []
for K, nid, clustered in ((700, 1, True), (4693, 5, False), (9000, 8, True), (3000, 3, True),
                          (5000, 80, True), (8192, 9, False), (900, 1000, True)):
    b = torch.from_numpy(synth.random_boxes(K, 800, 1333, seed=K, clustered=clustered)).cuda()
    g = torch.Generator().manual_seed(K)
    s = (torch.randint(0, 400, (K,), generator=g).float() / 400 - 0.2).cuda()   # ties, negatives
    ids = torch.randint(0, nid, (K,), generator=g).cuda()
    d, k = ops.batched_nms(b, s, ids, dict(type='nms', iou_threshold=0.6))
    res += [k.cpu().numpy().astype(np.float64), d.cpu().numpy().reshape(-1).astype(np.float64)]
    d, k = ops.nms(b, s, 0.5, offset=1)
    res += [k.cpu().numpy().astype(np.float64)]
np.save(sys.argv[1], np.concatenate(res))
''' % (os.path.dirname(here), here)
    outs = []
    with tempfile.TemporaryDirectory() as d:
        for mode in ('old', 'cluster'):
            path = os.path.join(d, mode + '.npy')
            subprocess.run([sys.executable, '-c', code, path], check=True,
                           env=dict(os.environ, BRCNN_NMS_OP=mode))
            outs.append(np.load(path))
    np.testing.assert_array_equal(outs[0], outs[1])


@pytest.mark.parametrize('n,nid,clustered', [(20480, 80, True), (20480, 80, False), (17000, 20, True)])
def test_batched_nms_coco_scale_many_ids(cuda, n, nid, clustered):
    """The reference's COCO multiclass_nms call shape (<= 20 480 candidates, 80 class ids,
    bbox_nms.py:86): per-id fused kernels + rank-counting merge over the whole GPU (more kept
    keys than one CTA sorts in shared memory); `num_ids` in the cfg avoids the host read."""
    boxes = synth.random_boxes(n, 800, 1333, seed=n + nid, clustered=clustered)
    scores = _scores(n, n + 3, dup=True)
    ids = np.random.RandomState(n + 1).randint(0, nid, n).astype(np.int64)
    t = lambda a: torch.from_numpy(a).to(cuda)
    _, ref = oracle.batched_nms(boxes, scores, ids, 0.5)
    for cfg in (dict(type='nms', iou_threshold=0.5), dict(type='nms', iou_threshold=0.5, num_ids=nid)):
        dets, keep = ops.batched_nms(t(boxes), t(scores), t(ids), cfg)
        np.testing.assert_array_equal(keep.cpu().numpy(), ref)
        np.testing.assert_array_equal(dets.cpu().numpy()[:, 4], scores[ref])


@pytest.mark.parametrize('nid', [12, 4])
def test_batched_nms_one_id_with_more_keeps_than_shared_memory(cuda, nid):
    """One id holds 9000 boxes and keeps them all: its list overflows the per-id sorting CTA
    (> 8192: the operator falls back to the rank-counting sort) and, with more than 8 ids (the
    per-id NMS path), its kept list spills from shared to global memory inside the kernel; with
    4 ids the clustered walk is taken."""
    rng = np.random.RandomState(3)
    n0 = 9000                     # a 100 x 90 lattice of disjoint 6x6 boxes: all kept
    gx, gy = np.meshgrid(np.arange(100) * 10.0, np.arange(90) * 8.0)
    b0 = np.stack([gx.ravel(), gy.ravel(), gx.ravel() + 6, gy.ravel() + 6], 1)
    b1 = synth.random_boxes(3000, 800, 1333, seed=11, clustered=True)
    boxes = np.concatenate([b0, b1]).astype(np.float32)
    ids = np.concatenate([np.zeros(n0), rng.randint(1, nid, 3000)]).astype(np.int64)
    scores = rng.permutation(len(boxes)).astype(np.float32) / len(boxes)
    t = lambda a: torch.from_numpy(a).to(cuda)
    dets, keep = ops.batched_nms(t(boxes), t(scores), t(ids), dict(type='nms', iou_threshold=0.5))
    _, ref = oracle.batched_nms(boxes, scores, ids, 0.5)
    assert (ids[ref] == 0).sum() == n0
    np.testing.assert_array_equal(keep.cpu().numpy(), ref)


@pytest.mark.parametrize('n,nid,max_num', [(4693, 5, 256), (20000, 5, 2000), (20480, 80, 100),
                                           (5000, 80, 100), (300, 3, 1000)])
def test_batched_nms_max_num_stops_early_with_the_same_prefix(cuda, n, nid, max_num):
    """nms_cfg['max_num'] (mmcv slices keep[:max_num] after the full NMS): the library stops its
    sweeps at max_num keeps; the result is the same prefix of the full keep list."""
    boxes = synth.random_boxes(n, 800, 1333, seed=3 * n + nid, clustered=True)
    scores = _scores(n, n + 5, dup=True)
    ids = np.random.RandomState(n + 2).randint(0, nid, n).astype(np.int64)
    t = lambda a: torch.from_numpy(a).to(cuda)
    ref = oracle.batched_nms(boxes, scores, ids, 0.7)[1][:max_num]
    dets, keep = ops.batched_nms(t(boxes), t(scores), t(ids),
                                 dict(type='nms', iou_threshold=0.7, max_num=max_num, num_ids=nid))
    np.testing.assert_array_equal(keep.cpu().numpy(), ref)
    np.testing.assert_array_equal(dets.cpu().numpy()[:, :4], boxes[ref])
