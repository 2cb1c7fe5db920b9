"""GPU parity: fused boosting-reweighted loss (forward value, accuracy and the
gradients w.r.t. cls_score / bbox_pred) vs the oracle and vs a literal torch
restatement of prob_roi_head.py:107-154.  Bar: <= 1e-5 relative."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from boosting_rcnn_b200 import ops
from oracle import oracle

pytestmark = pytest.mark.gpu


def _inputs(N, C, seed, pos_frac=0.25, agnostic=False):
    rng = np.random.RandomState(seed)
    cls = rng.normal(0, 2, (N, C + 1)).astype(np.float32)
    labels = np.full((N,), C, dtype=np.int64)
    npos = int(N * pos_frac)
    labels[:npos] = rng.randint(0, C, npos)
    prior = rng.rand(N).astype(np.float32)
    prior[:3] = 0.0  # GT rows carry prior 0 (prob_roi_head.py:57)
    bp = rng.normal(0, 0.5, (N, 4 if agnostic else 4 * C)).astype(np.float32)
    bt = rng.normal(0, 0.5, (N, 4)).astype(np.float32)
    bw = np.zeros((N, 4), np.float32)
    bw[:npos] = 1.0
    return cls, labels, prior, bp, bt, bw


def _torch_reference(cls, labels, prior, bp, bt, bw, C, gamma, wcls, wbbox, agnostic=False):
    """Literal restatement with torch ops (fp64 for a tight yardstick)."""
    cls = torch.tensor(cls, dtype=torch.float64, requires_grad=True)
    bp = torch.tensor(bp, dtype=torch.float64, requires_grad=True)
    labels_t = torch.tensor(labels)
    prior_t = torch.tensor(prior, dtype=torch.float64)
    w = (1 - prior_t) ** gamma
    loss = wcls * F.cross_entropy(cls, labels_t, reduction='none')
    new_w = w * (loss.sum() / (w * loss).sum())
    loss_cls = (loss * new_w.detach()).sum() / w.shape[0]
    pos = (labels_t >= 0) & (labels_t < C)
    if pos.any():
        pred = bp.view(bp.size(0), 4)[pos] if agnostic else bp.view(bp.size(0), -1, 4)[pos, labels_t[pos]]
        lb = wbbox * (torch.abs(pred - torch.tensor(bt, dtype=torch.float64)[pos]) *
                      torch.tensor(bw, dtype=torch.float64)[pos])
        loss_bbox = lb.sum() / bp.size(0)
    else:
        loss_bbox = bp[pos].sum()
    (loss_cls + loss_bbox).backward()
    acc = 100.0 * (cls.argmax(1) == labels_t).double().mean()
    return (loss_cls.item(), loss_bbox.item(), acc.item(), cls.grad.numpy(), bp.grad.numpy())


def _rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - b).max() / max(np.abs(b).max(), 1e-12))


@pytest.mark.parametrize('N,C,seed', [(1024, 80, 0), (1024, 4, 1), (37, 20, 2), (2048, 80, 3)])
def test_boost_loss_matches_references(cuda, N, C, seed):
    cls, labels, prior, bp, bt, bw = _inputs(N, C, seed)
    tc = torch.from_numpy(cls).to(cuda).requires_grad_(True)
    tb = torch.from_numpy(bp).to(cuda).requires_grad_(True)
    lc, lb, acc, sc = ops.boost_loss(tc, tb, torch.from_numpy(labels).to(cuda), None,
                                     torch.from_numpy(prior).to(cuda), torch.from_numpy(bt).to(cuda),
                                     torch.from_numpy(bw).to(cuda), C, gamma=0.5,
                                     loss_cls_weight=2.0, loss_bbox_weight=2.0)
    (lc + lb).backward()
    r_lc, r_lb, r_acc, r_gc, r_gb = _torch_reference(cls, labels, prior, bp, bt, bw, C, 0.5, 2.0, 2.0)
    assert abs(lc.item() - r_lc) <= 1e-5 * abs(r_lc)
    assert abs(lb.item() - r_lb) <= 1e-5 * abs(r_lb)
    assert abs(acc.item() - r_acc) <= 1e-4
    assert _rel(tc.grad.cpu().numpy(), r_gc) <= 1e-5
    assert _rel(tb.grad.cpu().numpy(), r_gb) <= 1e-5
    o = oracle.boost_loss(cls, labels, prior, bp, bt, bw, C, gamma=0.5, loss_cls_weight=2.0,
                          loss_bbox_weight=2.0)
    assert abs(lc.item() - o['loss_cls']) <= 1e-5 * abs(o['loss_cls'])
    assert _rel(tc.grad.cpu().numpy(), o['grad_cls'].astype(np.float64)) <= 1e-5
    assert _rel(tb.grad.cpu().numpy(), o['grad_bbox'].astype(np.float64)) <= 1e-5
    # F4: the forward VALUE equals the plain mean CE (times loss_weight)
    plain = 2.0 * F.cross_entropy(torch.tensor(cls, dtype=torch.float64), torch.tensor(labels)).item()
    assert abs(lc.item() - plain) <= 1e-5 * plain


def test_boost_loss_no_positives_and_agnostic(cuda):
    cls, labels, prior, bp, bt, bw = _inputs(256, 4, 5, pos_frac=0.0)
    tc = torch.from_numpy(cls).to(cuda).requires_grad_(True)
    tb = torch.from_numpy(bp).to(cuda).requires_grad_(True)
    lc, lb, acc, _ = ops.boost_loss(tc, tb, torch.from_numpy(labels).to(cuda), None,
                                    torch.from_numpy(prior).to(cuda), torch.from_numpy(bt).to(cuda),
                                    torch.from_numpy(bw).to(cuda), 4, loss_cls_weight=2.0,
                                    loss_bbox_weight=2.0)
    (lc + lb).backward()
    assert lb.item() == 0.0 and tb.grad.abs().max().item() == 0.0
    cls, labels, prior, bp, bt, bw = _inputs(300, 4, 6, agnostic=True)
    tc = torch.from_numpy(cls).to(cuda).requires_grad_(True)
    tb = torch.from_numpy(bp).to(cuda).requires_grad_(True)
    lc, lb, acc, _ = ops.boost_loss(tc, tb, torch.from_numpy(labels).to(cuda), None,
                                    torch.from_numpy(prior).to(cuda), torch.from_numpy(bt).to(cuda),
                                    torch.from_numpy(bw).to(cuda), 4, reg_class_agnostic=True,
                                    loss_cls_weight=1.0, loss_bbox_weight=1.0)
    (lc + lb).backward()
    r = _torch_reference(cls, labels, prior, bp, bt, bw, 4, 0.5, 1.0, 1.0, agnostic=True)
    assert abs(lb.item() - r[1]) <= 1e-5 * abs(r[1])
    assert _rel(tb.grad.cpu().numpy(), r[4]) <= 1e-5


def test_boost_loss_deterministic(cuda):
    cls, labels, prior, bp, bt, bw = _inputs(1024, 80, 9)
    outs = []
    for _ in range(2):
        tc = torch.from_numpy(cls).to(cuda).requires_grad_(True)
        tb = torch.from_numpy(bp).to(cuda).requires_grad_(True)
        lc, lb, acc, _ = ops.boost_loss(tc, tb, torch.from_numpy(labels).to(cuda), None,
                                        torch.from_numpy(prior).to(cuda), torch.from_numpy(bt).to(cuda),
                                        torch.from_numpy(bw).to(cuda), 80)
        (lc + lb).backward()
        outs.append((lc.item(), tc.grad.cpu().numpy()))
    assert outs[0][0] == outs[1][0]
    np.testing.assert_array_equal(outs[0][1].view(np.uint32), outs[1][1].view(np.uint32))
