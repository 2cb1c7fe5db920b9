"""GPU parity of the fused RPN loss path (brcnn_rpn_loss_forward /
brcnn_rpn_loss_scale behind ATSSRPNHead.loss) against
  (1) goldens produced by EXECUTING the reference's ATSSRPNHead.loss / loss_single /
      get_targets + its loss modules (tests/golden/make_golden_rpn_loss.py), and
  (2) the numpy oracle (oracle.rpn_loss), incl. at the full 201 600-anchor size.
Bar: loss values and gradients <= 1e-5 relative (fp32); anchor targets bit-exact (they decide
which terms exist at all)."""
import os

import numpy as np
import pytest
import torch

import synth
from boosting_rcnn_b200 import configs, ops
from boosting_rcnn_b200.anchors import AnchorGenerator
from oracle import oracle

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden',
                    'reference_golden_rpn_loss.npz')
RTOL = 1e-5


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max()) / max(float(np.abs(b).max()), 1e-12)


def _run_head(dev, c, cfg='utdac', varifocal=False):
    torch.manual_seed(0)
    rpn, _, _ = configs.build_hot_path(cfg, train=True)
    if varifocal:     # the VOC config's RPN classification loss on the 9-anchor head
        from boosting_rcnn_b200.registry import build_loss
        rpn.loss_cls = build_loss(dict(type='VarifocalLoss', use_sigmoid=True, alpha=0.75,
                                       gamma=2.0, iou_weighted=True, loss_weight=1.0))
    rpn = rpn.to(dev)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    cls = [t(a).requires_grad_(True) for a in c['cls']]
    box = [t(a).requires_grad_(True) for a in c['box']]
    iou = [t(a).requires_grad_(True) for a in c['iou']]
    gts = [t(g) for g in c['gt_bboxes']]
    losses = rpn.loss(cls, box, iou, gts, c['img_metas'])
    assert set(losses) == {'loss_rpn_cls', 'loss_rpn_bbox', 'loss_rpn_iou'}
    assert all(len(v) == len(cls) and all(x.dim() == 0 for x in v) for v in losses.values())
    sum(sum(v) for v in losses.values()).backward()        # detectors/base.py:186-199
    return rpn, losses, cls, box, iou


@pytest.mark.parametrize('case', synth.RPN_LOSS_CASES + ('varifocal_basic', 'varifocal_partial_valid'))
def test_rpn_loss_equals_executed_reference_and_oracle(cuda, case):
    g = np.load(GOLD)
    vf = case.startswith('varifocal_')
    c = synth.rpn_loss_case(case[len('varifocal_'):] if vf else case)
    rpn, losses, cls, box, iou = _run_head(cuda, c, varifocal=vf)
    gen = AnchorGenerator(strides=[8, 16, 32, 64, 128], ratios=[0.5, 1.0, 2.0],
                          octave_base_scale=4, scales_per_octave=3)
    kw = dict(cls_loss='varifocal', focal_alpha=0.75, focal_gamma=2.0) if vf else {}
    o = oracle.rpn_loss(c['cls'], c['box'], c['iou'], c['gt_bboxes'], c['img_metas'],
                        gen.base_anchor_table().numpy(), synth.STRIDES, **kw)
    for k in ('loss_rpn_cls', 'loss_rpn_bbox', 'loss_rpn_iou'):
        got = np.array([float(v) for v in losses[k]], dtype=np.float32)
        np.testing.assert_allclose(got, g[f'{case}/{k}'], rtol=RTOL, atol=1e-7, err_msg=k)
        np.testing.assert_allclose(got, o[k], rtol=RTOL, atol=1e-7, err_msg=k + ' (oracle)')
    for l in range(len(cls)):
        for name, ts in (('grad_cls', cls), ('grad_box', box), ('grad_iou', iou)):
            got = ts[l].grad.cpu().numpy()
            ref = g[f'{case}/{name}_{l}']
            if np.abs(ref).max() == 0:
                assert np.abs(got).max() == 0, (name, l)
            else:
                assert _rel(got, ref) <= RTOL, (name, l, _rel(got, ref))
                assert _rel(got, o[name][l]) <= RTOL, (name, l, 'oracle')


def test_rpn_loss_full_size_vs_oracle(cuda):
    """configs[2] geometry: 2 images x 201 600 anchors (1344x800 pad), 1-20 GTs per image; COCO
    head settings (gamma 2, loss weights 2)."""
    rng = np.random.RandomState(5)
    B, pad_hw, img_hw = 2, (800, 1344), (800, 1333)
    sizes = synth.featmap_sizes(*pad_hw)
    cls, box, iou = synth.rpn_outputs(B, sizes, 9, seed=77)
    gts = []
    for b in range(B):
        n = [17, 3][b]
        wh = np.exp(rng.uniform(np.log(16), np.log(500), (n, 2)))
        ctr = rng.uniform(0.1, 0.9, (n, 2)) * [img_hw[1], img_hw[0]]
        gb = np.concatenate([ctr - wh / 2, ctr + wh / 2], 1)
        gb[:, 0::2] = gb[:, 0::2].clip(0, img_hw[1])
        gb[:, 1::2] = gb[:, 1::2].clip(0, img_hw[0])
        gts.append(gb.astype(np.float32))
    metas = [dict(img_shape=img_hw + (3,), pad_shape=pad_hw + (3,)) for _ in range(B)]
    c = dict(cls=cls, box=box, iou=iou, gt_bboxes=gts, img_metas=metas)
    rpn, losses, tc, tb, ti = _run_head(cuda, c, cfg='coco')
    gen = AnchorGenerator(strides=[8, 16, 32, 64, 128], ratios=[0.5, 1.0, 2.0],
                          octave_base_scale=4, scales_per_octave=3)
    o = oracle.rpn_loss(cls, box, iou, gts, metas, gen.base_anchor_table().numpy(), synth.STRIDES,
                        gamma=2, w_bbox=2.0, w_aug=2.0)
    assert o['num_pos'] > 20
    for k in ('loss_rpn_cls', 'loss_rpn_bbox', 'loss_rpn_iou'):
        got = np.array([float(v) for v in losses[k]], dtype=np.float32)
        np.testing.assert_allclose(got, o[k], rtol=RTOL, atol=1e-7, err_msg=k)
    for l in range(5):
        for name, ts in (('grad_cls', tc), ('grad_box', tb), ('grad_iou', ti)):
            got, ref = ts[l].grad.cpu().numpy(), o[name][l]
            if np.abs(ref).max() == 0:
                assert np.abs(got).max() == 0
            else:
                assert _rel(got, ref) <= RTOL, (name, l, _rel(got, ref))


def test_rpn_loss_sums_and_determinism(cuda):
    """Raw sums expose num_total_pos / sum(iou_target); two runs are bit-identical (fixed-order
    reductions, no float atomics)."""
    c = synth.rpn_loss_case('basic')
    gen = AnchorGenerator(strides=[8, 16, 32, 64, 128], ratios=[0.5, 1.0, 2.0],
                          octave_base_scale=4, scales_per_octave=3)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    B = 2
    G = max(len(g) for g in c['gt_bboxes'])
    gtb = torch.zeros(B, G, 4, device=cuda)
    for b, gb in enumerate(c['gt_bboxes']):
        gtb[b, :len(gb)] = t(gb)
    num_gt = torch.tensor([len(g) for g in c['gt_bboxes']], dtype=torch.int32, device=cuda)
    pad = torch.tensor([m['pad_shape'][:2] for m in c['img_metas']], dtype=torch.float32, device=cuda)
    p = ops.make_rpn_loss_params(B, c['sizes'], synth.STRIDES, 9, G)
    outs = []
    for _ in range(2):
        lc, lb, li, sums = ops.rpn_loss(p, [t(a) for a in c['cls']], [t(a) for a in c['box']],
                                        [t(a) for a in c['iou']], gen.base_anchor_table().to(cuda),
                                        gtb, num_gt, pad)
        outs.append(torch.cat([torch.stack(lc), torch.stack(lb), torch.stack(li), sums]).cpu().numpy())
    np.testing.assert_array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))
    o = oracle.rpn_loss(c['cls'], c['box'], c['iou'], c['gt_bboxes'], c['img_metas'],
                        gen.base_anchor_table().numpy(), synth.STRIDES)
    sums = outs[0][15:]
    assert sums[15] == o['num_pos']
    np.testing.assert_allclose(sums[16], o['iou_sum'], rtol=1e-5)


def test_forward_train_returns_reference_keys(cuda):
    """ATSSRPNHead.forward_train (atss_rpn_head.py:270-294): losses + proposals, gradients reach
    the conv tower."""
    torch.manual_seed(0)
    rpn, _, model = configs.build_hot_path('utdac', train=True)
    rpn = rpn.to(cuda).train()
    sizes = synth.featmap_sizes(128, 160)
    x = [torch.from_numpy(f).to(cuda) for f in synth.fpn_feats(2, 256, sizes, seed=3)]
    c = synth.rpn_loss_case('basic')
    gts = [torch.from_numpy(g).to(cuda) for g in c['gt_bboxes']]
    losses, props = rpn.forward_train(x, c['img_metas'], gts, None,
                                      proposal_cfg=model['train_cfg']['rpn_proposal'])
    assert set(losses) == {'loss_rpn_cls', 'loss_rpn_bbox', 'loss_rpn_iou'}
    assert len(props) == 2 and props[0].shape[1] == 5
    sum(sum(v) for v in losses.values()).backward()
    assert rpn.rpn_cls.weight.grad is not None and rpn.rpn_reg.weight.grad.abs().sum() > 0
    assert rpn.rpn_iou.weight.grad.abs().sum() > 0 and rpn.scales[0].scale.grad is not None


def test_voc_config_rpn_loss_runs_and_unsupported_variants_raise(cuda):
    """VOC head (1 anchor / location, VarifocalLoss, gamma 2, loss weights 2) against the oracle;
    atss=True is refused with a clear message."""
    rpn, _, _ = configs.build_hot_path('voc', train=True)
    c = synth.rpn_loss_case('basic', num_anchors=1)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    cls = [t(a).requires_grad_(True) for a in c['cls']]
    losses = rpn.to(cuda).loss(cls, [t(a) for a in c['box']], [t(a) for a in c['iou']],
                               [t(g) for g in c['gt_bboxes']], c['img_metas'])
    gen = AnchorGenerator(strides=[8, 16, 32, 64, 128], ratios=[1.0], octave_base_scale=8,
                          scales_per_octave=1)
    o = oracle.rpn_loss(c['cls'], c['box'], c['iou'], c['gt_bboxes'], c['img_metas'],
                        gen.base_anchor_table().numpy(), synth.STRIDES, gamma=2, w_bbox=2.0,
                        w_aug=2.0, cls_loss='varifocal', focal_alpha=0.75)
    for k in ('loss_rpn_cls', 'loss_rpn_bbox', 'loss_rpn_iou'):
        got = np.array([float(v.detach()) for v in losses[k]], dtype=np.float32)
        np.testing.assert_allclose(got, o[k], rtol=RTOL, atol=1e-7, err_msg=k)
    rpn.atss = True
    with pytest.raises(NotImplementedError, match='atss=True'):
        rpn.loss(cls, [t(a) for a in c['box']], [t(a) for a in c['iou']],
                 [t(g) for g in c['gt_bboxes']], c['img_metas'])
