"""CPU, world_size 2, gloo: the N>1 plumbing (image sharding, the fused scalar
all-reduce that replaces the reference's 2 + 7 scalar all-reduces, max-over-
ranks timing)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from boosting_rcnn_b200 import dist as bdist


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    lo, hi = bdist.shard_range(33)
    losses = {k: torch.tensor(float(i + 1 + 10 * rank)) for i, k in enumerate(
        ['loss_rpn_cls', 'loss_rpn_bbox', 'loss_rpn_iou', 'loss_cls', 'acc', 'loss_bbox', 'loss'])}
    fused = bdist.fused_scalar_allreduce(losses)
    each = {k: bdist.reduce_mean(v) for k, v in losses.items()}
    ms = bdist.max_over_ranks(1.5 + rank)
    # result wire: rank r owns images [lo, hi) of 5; image i has i detections, rows tagged i
    lo5, hi5 = bdist.shard_range(5)
    M = 4
    det = torch.zeros(hi5 - lo5, M, 5)
    lab = torch.zeros(hi5 - lo5, M, dtype=torch.int64)
    num = torch.zeros(hi5 - lo5, dtype=torch.int32)
    for j, i in enumerate(range(lo5, hi5)):
        det[j, :i] = float(i)
        lab[j, :i] = i
        num[j] = min(i, M)
    col = bdist.collect_detections(det, lab, num, size=5)
    col = None if col is None else [(tuple(c.shape), float(c.sum())) for c in col]
    out.put((rank, lo, hi, {k: float(v) for k, v in fused.items()},
             {k: float(v) for k, v in each.items()}, ms, col))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_fused_allreduce():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, f0, e0, m0, c0), (r1, lo1, hi1, f1, e1, m1, c1) = res
    # collect_detections: rank 0 holds all 5 images in dataset order, rank 1 nothing
    assert c1 is None and [c[0] for c in c0] == [(0, 6), (1, 6), (2, 6), (3, 6), (4, 6)]
    assert [c[1] for c in c0] == [0.0, 6.0, 24.0, 54.0, 96.0]
    assert (lo0, hi0, lo1, hi1) == (0, 17, 17, 33)  # every image exactly once
    assert f0 == f1 == e0 == e1  # one fused collective == 7 separate reduce_means
    assert f0['loss_cls'] == (4 + 14) / 2 and list(f0) == list(e0)
    assert m0 == m1 == 2.5


def test_single_process_paths_are_identity():
    assert bdist.shard_range(16) == (0, 16)
    assert bdist.get_dist_info() == (0, 1)
    d = {'a': torch.tensor(2.0)}
    assert bdist.fused_scalar_allreduce(d)['a'].item() == 2.0
    assert [bdist.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    det = torch.arange(2 * 3 * 5, dtype=torch.float32).reshape(2, 3, 5)
    col = bdist.collect_detections(det, torch.tensor([[1, 2, 3], [4, 5, 6]]), torch.tensor([2, 0]))
    assert [tuple(c.shape) for c in col] == [(2, 6), (0, 6)] and col[0][1, 5] == 2


RESULTS_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden',
                            'reference_golden_results.npz')


def _wire_worker(rank, world, port, out):
    import numpy as np
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    g = np.load(RESULTS_GOLD)
    mine = g['shards'][rank]                        # DistributedSampler shard (round-robin, padded)
    det = torch.from_numpy(g['det'][mine])
    lab = torch.from_numpy(g['lab'][mine])
    num = torch.from_numpy(g['num'][mine])
    col = bdist.collect_detections(det, lab, num, size=int(g['det'].shape[0]), interleaved=True)
    out.put((rank, None if col is None else [c.numpy() for c in col]))
    dist.destroy_process_group()


def test_result_wire_equals_executed_reference():
    """``dist.collect_detections(interleaved=True)`` + ``bbox2result`` give,
    on rank 0, exactly what the reference's ``collect_results_cpu`` (mmdet/apis/test.py:273-313)
    returns for the same per-rank ``bbox2result`` lists (transforms.py:100-117) — golden produced
    by executing both reference functions (tests/golden/make_golden_results.py): dataset order
    restored from the round-robin shards, the sampler's padding image dropped, per-class arrays
    bit-identical (incl. an image without detections)."""
    import numpy as np
    from boosting_rcnn_b200.roi_head import bbox2result
    g = np.load(RESULTS_GOLD)
    world, C = int(g['world']), int(g['num_classes'])
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_wire_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(res[r] is None for r in range(1, world))
    col = res[0]
    assert len(col) == g['det'].shape[0]
    for i, rows in enumerate(col):
        per_class = bbox2result(rows[:, :5], rows[:, 5].astype(np.int64), C)
        assert len(per_class) == C
        for c in range(C):
            ref = g[f'out/{i}/{c}']
            assert per_class[c].shape == ref.shape, (i, c)
            np.testing.assert_array_equal(per_class[c].view(np.uint32), ref.view(np.uint32))
