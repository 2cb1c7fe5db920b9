"""GPU: the drop-in module classes (registry names of the reference) and the
executed-reference golden vectors, through the CUDA path."""
import os

import numpy as np
import pytest
import torch

import synth
from boosting_rcnn_b200 import configs, ops
from boosting_rcnn_b200.anchors import AnchorGenerator
from boosting_rcnn_b200.roi_head import bbox2roi, pad_proposals
from boosting_rcnn_b200.rpn_head import PaddedProposals, unpad_proposals
from oracle import oracle

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_golden.npz'))


def _metas(batch, img_hw=(250, 317), pad_hw=(256, 320), sf=(1.25, 1.2, 1.25, 1.2)):
    return [dict(img_shape=(*img_hw, 3), pad_shape=(*pad_hw, 3),
                 scale_factor=np.array(sf, dtype=np.float32)) for _ in range(batch)]


# ------------------------------------------------------------- golden vectors
def test_golden_delta2bbox_cuda(cuda):
    t = lambda k: torch.from_numpy(G[k]).to(cuda)
    out1 = ops.delta2bbox(t('d2b_rois'), t('d2b_deltas1'), max_shape=(320, 400, 3)).cpu().numpy()
    out4 = ops.delta2bbox(t('d2b_rois'), t('d2b_deltas4'), stds=(.1, .1, .2, .2),
                          max_shape=(320, 400, 3)).cpu().numpy()
    outn = ops.delta2bbox(t('d2b_rois'), t('d2b_deltas1')).cpu().numpy()
    assert np.allclose(out1, G['d2b_out1'], rtol=1e-5, atol=1e-4)
    assert np.allclose(out4, G['d2b_out4'], rtol=1e-5, atol=1e-4)
    assert np.allclose(outn, G['d2b_out1_noclip'], rtol=1e-5, atol=1e-4)
    # and bit-exact against the oracle (pinned exp)
    np.testing.assert_array_equal(
        out4.view(np.uint32),
        oracle.delta2bbox(G['d2b_rois'], G['d2b_deltas4'], stds=(.1, .1, .2, .2),
                          max_shape=(320, 400)).view(np.uint32))


def test_golden_rpn_get_bboxes_cuda(cuda):
    rpn, _, _ = configs.build_hot_path('utdac')
    rpn = rpn.to(cuda)
    t = lambda k: torch.from_numpy(G[k])[None].to(cuda)
    props = rpn.get_bboxes([t(f'rpn_cls_{l}') for l in range(5)], [t(f'rpn_box_{l}') for l in range(5)],
                           [t(f'rpn_iou_{l}') for l in range(5)], [dict(img_shape=(90, 125, 3))],
                           cfg=dict(nms_pre=60, max_per_img=40, nms=dict(type='nms', iou_threshold=0.7),
                                    min_bbox_size=0))
    assert len(props) == 1
    p, ref = props[0].cpu().numpy(), G['rpn_proposals']
    assert p.shape == ref.shape
    assert np.allclose(p[:, 4], ref[:, 4], rtol=1e-6, atol=1e-7)
    assert np.allclose(p[:, :4], ref[:, :4], rtol=1e-5, atol=1e-4)


def test_golden_levels_and_fusion_cuda(cuda):
    lv = ops.map_roi_levels(torch.from_numpy(G['lvl_rois']).to(cuda), 5, 56).cpu().numpy()
    np.testing.assert_array_equal(lv, G['lvl_out'])
    R, C1 = G['fuse_cls'].shape
    p = ops.make_rcnn_params(1, R, C1 - 1, 0.05, 0.5, 100)
    lay = ops.rcnn_workspace_layout(p)
    rois = torch.from_numpy(synth.random_rois(1, R, 200, 300, seed=0)).to(cuda)
    _, _, _, ws = ops.rcnn_get_bboxes(p, rois, torch.from_numpy(G['fuse_prior']).to(cuda),
                                      torch.tensor([R], dtype=torch.int32, device=cuda),
                                      torch.from_numpy(G['fuse_cls']).to(cuda),
                                      torch.zeros(R, 4 * (C1 - 1), device=cuda),
                                      torch.tensor([[200., 300.]], device=cuda), return_workspace=True)
    fused = ws.cpu().numpy()[lay.scores:lay.scores + R * C1 * 4].view(np.float32).reshape(R, C1)
    assert np.allclose(fused, G['fuse_out'], rtol=2e-6, atol=1e-7)


def test_golden_boost_loss_cuda(cuda):
    N, C1 = G['loss_cls_score'].shape
    cs = torch.from_numpy(G['loss_cls_score']).to(cuda).requires_grad_(True)
    bp = torch.zeros(N, 4 * (C1 - 1), device=cuda, requires_grad=True)
    z4 = torch.zeros(N, 4, device=cuda)
    lc, lb, acc, _ = ops.boost_loss(cs, bp, torch.from_numpy(G['loss_labels']).to(cuda), None,
                                    torch.from_numpy(G['loss_prior']).to(cuda), z4, z4, C1 - 1,
                                    gamma=0.5, loss_cls_weight=2.0, loss_bbox_weight=2.0)
    lc.backward()
    assert abs(lc.item() - G['loss_cls']) <= 1e-5 * abs(G['loss_cls'])
    assert abs(acc.item() - G['loss_acc']) <= 1e-4
    assert np.abs(cs.grad.cpu().numpy() - G['loss_grad_cls']).max() <= 1e-5 * np.abs(G['loss_grad_cls']).max()


# ------------------------------------------------------------- module classes
def _rpn_inputs(dev, batch, seed):
    sizes = synth.featmap_sizes(256, 320)
    cls, box, iou = synth.rpn_outputs(batch, sizes, 9, seed=seed)
    t = lambda arrs: [torch.from_numpy(a).to(dev) for a in arrs]
    return sizes, cls, box, iou, t(cls), t(box), t(iou)


def test_rpn_head_get_bboxes_list_api(cuda):
    rpn, _, m = configs.build_hot_path('utdac')
    rpn = rpn.to(cuda)
    sizes, cls, box, iou, tc, tb, tu = _rpn_inputs(cuda, 3, 0)
    props = rpn.get_bboxes(tc, tb, tu, _metas(3), rescale=True)  # rescale is ignored like the reference
    base = rpn.anchor_generator.base_anchor_table().numpy()
    cfg = m['test_cfg']['rpn']
    for b in range(3):
        ref = oracle.rpn_get_bboxes_single([c[b] for c in cls], [c[b] for c in box], [c[b] for c in iou],
                                           base, synth.STRIDES, (250, 317), cfg['nms_pre'],
                                           cfg['max_per_img'], 0.7, 0)
        np.testing.assert_array_equal(props[b].cpu().numpy().view(np.uint32), ref.view(np.uint32))
    with pytest.raises(AssertionError):
        rpn.get_bboxes(tc, tb, tu, _metas(3), with_nms=False)


def test_rpn_head_conv_tower_and_simple_test(cuda):
    torch.manual_seed(0)
    rpn, _, _ = configs.build_hot_path('utdac')
    rpn = rpn.to(cuda).eval()
    feats = [torch.randn(2, 256, h, w, device=cuda) for h, w in synth.featmap_sizes(256, 320)]
    with torch.no_grad():
        cls, box, iou = rpn(feats)
        assert cls[0].shape == (2, 9, 32, 40) and box[0].shape == (2, 36, 32, 40) and iou[4].shape == (2, 9, 2, 3)
        props = rpn.simple_test_rpn(feats, _metas(2))
        padded = rpn.simple_test_rpn(feats, _metas(2), padded=True)
    assert len(props) == 2 and props[0].shape[1] == 5 and props[0].shape[0] <= 256
    assert isinstance(padded, PaddedProposals)
    for a, b in zip(props, unpad_proposals(padded)):
        assert torch.equal(a, b)
    # scores sorted descending, boxes inside the image
    s = props[0][:, 4]
    assert torch.all(s[:-1] >= s[1:]) and props[0][:, :4].min() >= 0 and props[0][:, 2].max() <= 317


def test_roi_head_simple_test_matches_oracle(cuda):
    torch.manual_seed(1)
    rpn, roi, m = configs.build_hot_path('utdac')
    rpn, roi = rpn.to(cuda).eval(), roi.to(cuda).eval()
    # make the classifier non-trivial so that scores straddle the threshold
    torch.nn.init.normal_(roi.bbox_head.fc_cls.weight, 0, 0.05)
    torch.nn.init.normal_(roi.bbox_head.fc_reg.weight, 0, 0.02)
    B = 3
    sizes, cls, box, iou, tc, tb, tu = _rpn_inputs(cuda, B, 5)
    feats_np = synth.fpn_feats(B, 256, sizes, seed=6)
    feats = [torch.from_numpy(f).to(cuda) for f in feats_np]
    metas = _metas(B)
    with torch.no_grad():
        padded = rpn.get_bboxes_padded(tc, tb, tu, metas)
        plist = unpad_proposals(padded)
        for rescale in (False, True):
            det_b, det_l = roi.simple_test_bboxes(feats, metas, padded, roi.test_cfg, rescale=rescale)
            det_b2, det_l2 = roi.simple_test_bboxes(feats, metas, plist, roi.test_cfg, rescale=rescale)
            rois = bbox2roi(plist)
            res = roi._bbox_forward(feats, rois)
            cs, bp = res['cls_score'].cpu().numpy(), res['bbox_pred'].cpu().numpy()
            off = 0
            for b in range(B):
                n = plist[b].shape[0]
                pr = plist[b].cpu().numpy()
                r = np.concatenate([np.zeros((n, 1), np.float32), pr[:, :4]], 1)
                fused = oracle.fuse_scores(cs[off:off + n], pr[:, 4])
                rd, rl = oracle.rcnn_get_bboxes_single(r, fused, bp[off:off + n], (250, 317),
                                                       metas[b]['scale_factor'], 4, 0.05, 0.7, 100,
                                                       rescale=rescale)
                off += n
                for db, dl in ((det_b, det_l), (det_b2, det_l2)):
                    np.testing.assert_array_equal(db[b].cpu().numpy().view(np.uint32), rd.view(np.uint32))
                    np.testing.assert_array_equal(dl[b].cpu().numpy(), rl)
        results = roi.simple_test(feats, padded, metas, rescale=True)
    assert len(results) == B and len(results[0]) == 4
    assert all(r.shape[1] == 5 for img in results for r in img)
    assert sum(r.shape[0] for r in results[0]) == det_b[0].shape[0]


def test_roi_head_edge_cases(cuda):
    _, roi, _ = configs.build_hot_path('utdac')
    roi = roi.to(cuda).eval()
    sizes = synth.featmap_sizes(256, 320)
    feats = [torch.from_numpy(f).to(cuda) for f in synth.fpn_feats(2, 256, sizes, seed=2)]
    metas = _metas(2)
    empty = torch.zeros(0, 5, device=cuda)
    some = torch.tensor([[10., 10., 80., 90., 0.9], [100., 40., 200., 200., 0.7]], device=cuda)
    with torch.no_grad():
        # whole batch without proposals (prob_roi_head.py:216-225)
        db, dl = roi.simple_test_bboxes(feats, metas, [empty, empty], roi.test_cfg)
        assert all(d.shape == (0, 5) for d in db) and all(l.shape == (0,) and l.dtype == torch.long for l in dl)
        db, dl = roi.simple_test_bboxes(feats, metas, [empty, empty], None)
        assert db[0].shape == (0, 4) and dl[0].shape == (0, 5)
        # one image without proposals (:264-271)
        db, dl = roi.simple_test_bboxes(feats, metas, [some, empty], roi.test_cfg)
        assert db[1].shape == (0, 5) and dl[1].shape == (0,) and db[0].shape[1] == 5
        # rcnn_test_cfg is None: raw decoded boxes and fused scores (convfc_bbox_head.py:323-324)
        bb, sc = roi.simple_test_bboxes(feats, metas, [some, some[:1]], None)
        assert bb[0].shape == (2, 16) and sc[0].shape == (2, 5) and bb[1].shape == (1, 16)
        # extractor with zero rois
        assert roi.bbox_roi_extractor(feats, torch.zeros(0, 5, device=cuda)).shape == (0, 256, 7, 7)


def test_roi_head_forward_train_boost(cuda):
    torch.manual_seed(3)
    _, roi, m = configs.build_hot_path('coco', train=True)
    roi = roi.to(cuda).train()
    B = 2
    sizes = synth.featmap_sizes(256, 320)
    feats = [torch.from_numpy(f).to(cuda).requires_grad_(True) for f in synth.fpn_feats(B, 256, sizes, seed=4)]
    rng = np.random.RandomState(5)
    gts, labels, plist = [], [], []
    for b in range(B):
        g = synth.random_boxes(4, 250, 317, seed=10 + b)
        gts.append(torch.from_numpy(g).to(cuda))
        labels.append(torch.from_numpy(rng.randint(0, 80, 4)).to(cuda))
        jit = np.concatenate([g + rng.normal(0, 3, g.shape) for _ in range(30)]).astype(np.float32)
        rnd = synth.random_boxes(380, 250, 317, seed=20 + b)
        bx = np.concatenate([jit, rnd])
        sc = np.sort(rng.rand(len(bx)).astype(np.float32))[::-1].copy()
        plist.append(torch.from_numpy(np.concatenate([bx, sc[:, None]], 1)).to(cuda))
    losses = roi.forward_train(feats, _metas(B), plist, gts, labels)
    assert set(losses) == {'loss_cls', 'acc', 'loss_bbox'}
    assert all(torch.isfinite(v) for v in losses.values())
    (losses['loss_cls'] + losses['loss_bbox']).backward()
    assert all(f.grad is not None and torch.isfinite(f.grad).all() for f in feats)
    assert sum(f.grad.abs().sum().item() for f in feats) > 0
    assert roi.bbox_head.fc_cls.weight.grad.abs().sum().item() > 0
    assert roi.bbox_head.shared_fcs[0].weight.grad.abs().sum().item() > 0


def test_cuda_graph_replay_equals_eager_and_pipeline(cuda):
    """HotPathGraph / HostPipeline (graph.py): the captured step is bit-identical
    to the eager one, replays follow in-place input updates, and the double
    buffered host pipeline returns the same detections."""
    from boosting_rcnn_b200.graph import HostPipeline, HotPathGraph
    torch.manual_seed(0)
    rpn, roi, model = configs.build_hot_path('utdac')
    rpn, roi = rpn.to(cuda).eval(), roi.to(cuda).eval()
    B, pad_hw = 2, (256, 320)
    sizes = synth.featmap_sizes(*pad_hw)
    metas = _metas(B)
    host = []
    for seed in (0, 1):
        cls, box, iou = synth.rpn_outputs(B, sizes, rpn.num_anchors, seed=seed)
        feats = synth.fpn_feats(B, 256, sizes, seed=seed + 10)
        host.append(tuple([torch.from_numpy(a).pin_memory() for a in ts]
                          for ts in (feats, cls, box, iou)))
    dev0 = tuple([t.to(cuda) for t in ts] for ts in host[0])
    g = HotPathGraph(rpn, roi, metas, *dev0, rcnn_test_cfg=model['test_cfg']['rcnn'])
    assert g.launches_per_replay >= 8
    eager = [o.clone() for o in g.eager()]
    replay = [o.clone() for o in g.replay()]
    for a, b in zip(eager, replay):
        assert torch.equal(a, b)
    # refill the static inputs with the second sample
    for dst, src in zip(dev0, host[1]):
        for d, h in zip(dst, src):
            d.copy_(h)
    second = [o.clone() for o in g.replay()]
    assert not torch.equal(second[0], replay[0])
    assert all(torch.equal(a, b) for a, b in zip(second, g.eager()))
    pipe = HostPipeline(rpn, roi, metas, host[0], rcnn_test_cfg=model['test_cfg']['rcnn'])
    t0 = pipe.submit(*host[0])
    t1 = pipe.submit(*host[1])
    r0 = [o.clone() for o in t0.result()]
    r1 = [o.clone() for o in t1.result()]
    t2 = pipe.submit(*host[0])
    r2 = t2.result()
    pipe.drain()
    for a, b, c in zip(r0, replay, r2):
        assert torch.equal(a, b.cpu()) and torch.equal(c, b.cpu())
    for a, b in zip(r1, second):
        assert torch.equal(a, b.cpu())


@pytest.mark.parametrize('case', synth.RCNN_TRAIN_CASES)
def test_fused_assign_sample_targets_equals_executed_reference(cuda, case):
    """brcnn_rcnn_assign + brcnn_rcnn_sample_targets (2 launches + the reference's CPU
    randperm) against (1) the golden produced by EXECUTING the reference's MaxIoUAssigner /
    RandomSampler / BBoxHead.get_targets / ProbRoIHead.forward_train
    (tests/golden/make_golden_train.py) and (2) the numpy oracle, under the same CPU RNG
    seed: rois, labels, weights and priors bit-exact, bbox targets <= 1e-6 (logf).
    `many_positives` reproduces the reference's misaligned pos_inds[num_gts:] prior slice."""
    import os
    from test_oracle_golden import TRAIN_GOLD, _train_prep_oracle
    gold = np.load(TRAIN_GOLD)
    torch.manual_seed(3)
    _, roi, m = configs.build_hot_path('coco', train=True)
    roi = roi.to(cuda).train()
    gts_h, labels_h, plist_h = synth.rcnn_train_case(case)
    B = len(plist_h)
    gts = [torch.from_numpy(g).to(cuda) for g in gts_h]
    labels = [torch.from_numpy(l).to(cuda) for l in labels_h]
    plist = [torch.from_numpy(p).to(cuda) for p in plist_h]
    a, s, h = roi.bbox_assigner, roi.bbox_sampler, roi.bbox_head
    assert (a.pos_iou_thr, a.neg_iou_thr, a.min_pos_iou, s.num, s.pos_fraction) == \
        (0.6, 0.6, 0.6, 512, 0.25)
    torch.manual_seed(123)
    pp = pad_proposals(plist)
    rois, lab, lw, bt, bw, prior, rows = ops.rcnn_assign_sample(
        pp.boxes, pp.num, gts, labels, h.num_classes, a.pos_iou_thr, a.neg_iou_thr, a.min_pos_iou,
        s.num, s.pos_fraction, s.neg_pos_ub, h.bbox_coder.means, h.bbox_coder.stds,
        roi.train_cfg.pos_weight)
    o = _train_prep_oracle(case)
    for ref, what in ((lambda k: gold[f'{case}/{k}'], 'executed reference'),
                      (lambda k: o[k], 'oracle')):
        assert rows == [int(v) for v in ref('rows')], what
        for k, t in (('rois', rois), ('label_weights', lw), ('bbox_weights', bw), ('prior', prior)):
            np.testing.assert_array_equal(t.cpu().numpy().view(np.uint32),
                                          np.asarray(ref(k)).view(np.uint32), f'{k} vs {what}')
        np.testing.assert_array_equal(lab.cpu().numpy(), ref('labels'))
        np.testing.assert_allclose(bt.cpu().numpy(), ref('bbox_targets'), rtol=1e-6, atol=1e-6)
    if case == 'many_positives':
        assert all(int(n) == 128 for n in gold[f'{case}/num_pos'])   # randperm on the positives
    # and through the module: the Python fallback path and the fused path give the same losses
    sizes = synth.featmap_sizes(256, 320)
    feats = [torch.from_numpy(f).to(cuda) for f in synth.fpn_feats(B, 256, sizes, seed=4)]
    out = []
    for force in (True, False):
        roi.force_python_train_prep = force
        torch.manual_seed(77)
        out.append(roi.forward_train(feats, _metas(B), plist, gts, labels))
    for k in ('loss_cls', 'loss_bbox', 'acc'):
        assert torch.allclose(out[0][k], out[1][k], rtol=1e-6, atol=1e-7), k


def test_train_graph_and_assign_async_match_eager(cuda):
    """roi_head.train_graph (graph.py::RcnnTrainGraph: the R-CNN half of the training step as
    a forward and a backward CUDA graph) and ``assign_async`` + ``forward_train(assigned=)``
    give the losses and every gradient of the plain eager ``forward_train`` (bit for bit
    without the graph, <= 1e-5 of the largest element with it), step after step with different proposals / permutations through the SAME captured graphs."""
    torch.manual_seed(5)
    _, roi, m = configs.build_hot_path('coco', train=True)
    roi = roi.to(cuda).train()
    sizes = synth.featmap_sizes(256, 320)
    params = [p for p in roi.parameters() if p.requires_grad]

    def run(case, seed, mode):
        gts_h, labels_h, plist_h = synth.rcnn_train_case(case)
        B = len(plist_h)
        gts = [torch.from_numpy(g).to(cuda) for g in gts_h]
        labels = [torch.from_numpy(l).to(cuda) for l in labels_h]
        plist = [torch.from_numpy(p).to(cuda) for p in plist_h]
        feats = [torch.from_numpy(f).to(cuda).requires_grad_(True)
                 for f in synth.fpn_feats(B, 256, sizes, seed=seed)]
        for p in params:
            p.grad = None
        roi.train_graph = mode == 'graph'
        torch.manual_seed(seed)
        if mode == 'eager':
            out = roi.forward_train(feats, _metas(B), plist, gts, labels)
        else:
            pending = roi.assign_async(plist, gts, labels)
            assert pending is not None
            torch.zeros(1 << 20, device=cuda).sum()           # unrelated work in between
            out = roi.forward_train(feats, _metas(B), plist, gts, labels, assigned=pending)
        (out['loss_cls'] + 3 * out['loss_bbox']).backward()
        torch.cuda.synchronize()
        return ([out[k].detach().clone() for k in ('loss_cls', 'loss_bbox', 'acc')]
                + [p.grad.clone() for p in params] + [f.grad.clone() for f in feats])

    cases = [c for c in synth.RCNN_TRAIN_CASES]
    seen = 0
    for step, case in enumerate(cases + cases[:1]):
        ref = run(case, 100 + step, 'eager')
        for mode in ('async', 'graph'):
            got = run(case, 100 + step, mode)
            assert len(ref) == len(got)
            for i, (a, b) in enumerate(zip(ref, got)):
                if mode == 'async':      # same launches in the same order
                    assert torch.equal(a, b), (case, mode, i)
                else:                    # cuBLAS may pick another algorithm under capture
                    err = (a.float() - b.float()).abs().max().item()
                    assert err <= 1e-5 * a.float().abs().max().item() + 1e-12, (case, mode, i, err)
        seen += 1
    graphs = [g for g in roi._train_graphs.values() if g]
    assert graphs and all(roi._train_graphs.values()), 'capture failed: eager fallback ran'
    # the repeated case replayed an existing graph instead of capturing another one
    assert len(roi._train_graphs) <= len(cases)
    assert graphs[0].launches_per_step >= 5
    roi.train_graph = False


def test_dual_stream_runner_matches_single_graph(cuda):
    """graph.py::DualStreamRunner: two graphs replayed alternately on two streams give the
    same detections as one graph, for any interleaving."""
    from boosting_rcnn_b200.graph import DualStreamRunner, HotPathGraph
    torch.manual_seed(0)
    rpn, roi, model = configs.build_hot_path('utdac')
    rpn, roi = rpn.to(cuda).eval(), roi.to(cuda).eval()
    B, sizes = 2, synth.featmap_sizes(256, 320)
    cls, box, iou = synth.rpn_outputs(B, sizes, rpn.num_anchors, seed=5)
    t = lambda arrs: [torch.from_numpy(a).to(cuda) for a in arrs]
    ins = (t(synth.fpn_feats(B, 256, sizes, seed=6)), t(cls), t(box), t(iou))
    g1 = HotPathGraph(rpn, roi, _metas(B), *ins, rcnn_test_cfg=model['test_cfg']['rcnn'])
    g2 = HotPathGraph(rpn, roi, _metas(B), *ins, rcnn_test_cfg=model['test_cfg']['rcnn'])
    ref = [o.clone() for o in g1.replay()]
    torch.cuda.synchronize()
    dual = DualStreamRunner([g1, g2])
    outs = []
    for _ in range(7):
        o = dual.step()
        outs.append(o)
    dual.drain()
    torch.cuda.synchronize()
    for g in (g1, g2):
        for a, b in zip(g.outputs, ref):
            assert torch.equal(a, b)


def test_golden_rpn_train_size_split_path_cuda(cuda):
    """The CUDA path against the train-size golden produced by executing the reference's
    _get_bboxes_single (11 780 candidates, mmcv split path): same 2000 proposals up to swaps
    of rows whose scores agree to ~1e-7 (torch.sigmoid vs pinned exp, unstable torch.sort)."""
    from test_oracle_golden import _match_rows_allowing_near_tie_swaps, _rpn_train_golden_inputs
    g, sizes, cls, box, iou = _rpn_train_golden_inputs()
    gen = AnchorGenerator(strides=[8, 16, 32, 64, 128], ratios=[0.5, 1.0, 2.0],
                          octave_base_scale=4, scales_per_octave=3)
    p = ops.make_rpn_params(1, sizes, synth.STRIDES, 9, 4000, 2000, 0.7, 0.0)
    t = lambda arrs: [torch.from_numpy(a).to(cuda) for a in arrs]
    hw = torch.tensor([list(g['img_shape'])], dtype=torch.float32, device=cuda)
    props, num = ops.rpn_get_bboxes(p, t(cls), t(box), t(iou), gen.base_anchor_table().to(cuda), hw)
    assert int(num[0]) == 2000
    assert _match_rows_allowing_near_tie_swaps(props[0].cpu().numpy(), g['proposals']) <= 20


def test_golden_multiclass_nms_coco_scale_cuda(cuda):
    """The batched_nms operator on the COCO-scale golden of the executed reference
    multiclass_nms (80 classes, > 10 000 candidates)."""
    from test_oracle_golden import _multiclass_coco_candidates
    g, boxes, scores, labels = _multiclass_coco_candidates()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    dets, keep = ops.batched_nms(t(boxes), t(scores), t(labels), dict(type='nms', iou_threshold=0.5))
    np.testing.assert_array_equal(labels[keep.cpu().numpy()][:100], g['labels'])
    np.testing.assert_array_equal(dets[:100].cpu().numpy().view(np.uint32), g['dets'].view(np.uint32))


def test_simple_test_bboxes_without_nms_cfg_goes_through_the_kernel(cuda):
    """rcnn_test_cfg=None (convfc_bbox_head.py:323-324): per-image decoded boxes (n,4C) and
    fused scores (n,C+1) come from the fusion + decode kernel (no eager torch path) and equal
    the oracle's fuse_scores / delta2bbox on the head outputs (1e-5: the 2-fc head runs on a
    differently padded batch here, cuBLAS may round its GEMMs differently)."""
    from boosting_rcnn_b200 import _lib
    torch.manual_seed(0)
    _, roi, _ = configs.build_hot_path('utdac')
    roi = roi.to(cuda).eval()
    B, sizes = 2, synth.featmap_sizes(256, 320)
    feats = [torch.from_numpy(f).to(cuda) for f in synth.fpn_feats(B, 256, sizes, seed=8)]
    rng = np.random.RandomState(9)
    plist = []
    for b, n in enumerate((37, 12)):
        bx = synth.random_boxes(n, 250, 317, seed=20 + b)
        sc = np.sort(rng.rand(n).astype(np.float32))[::-1].copy()
        plist.append(torch.from_numpy(np.concatenate([bx, sc[:, None]], 1).astype(np.float32)).to(cuda))
    metas = _metas(B)
    l0 = _lib.load().brcnn_launch_count()
    with torch.no_grad():
        bbs, scs = roi.simple_test_bboxes(feats, metas, plist, None, rescale=True)
    assert _lib.load().brcnn_launch_count() - l0 >= 4
    with torch.no_grad():
        rois = bbox2roi(plist)
        res = roi._bbox_forward(feats, rois)
    cs, bp = res['cls_score'].cpu().numpy(), res['bbox_pred'].cpu().numpy()
    off = 0
    for b, p in enumerate(plist):
        n = p.size(0)
        assert bbs[b].shape == (n, 16) and scs[b].shape == (n, 5)
        fused = oracle.fuse_scores(cs[off:off + n], p[:, 4].cpu().numpy())
        np.testing.assert_allclose(scs[b].cpu().numpy(), fused, rtol=1e-5, atol=1e-6)
        dec = oracle.delta2bbox(p[:, :4].cpu().numpy(), bp[off:off + n], (0., 0., 0., 0.),
                                (.1, .1, .2, .2), max_shape=metas[b]['img_shape'])
        dec = (dec.reshape(n, -1, 4) / np.asarray(metas[b]['scale_factor'], np.float32)).reshape(n, -1)
        np.testing.assert_allclose(bbs[b].cpu().numpy(), dec, rtol=1e-5, atol=1e-3)
        off += n
