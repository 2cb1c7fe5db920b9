"""CPU: host-side mirror of the reference interfaces (registry, configs,
padding helpers, assign/sample glue)."""
import os

import numpy as np
import pytest
import torch

from boosting_rcnn_b200 import configs, ops, roi_head, sampling
from boosting_rcnn_b200.registry import HEADS, ConfigDict, build_from_cfg, build_head
from boosting_rcnn_b200.rpn_head import PaddedProposals

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_golden.npz'))


def test_registry_names_and_config_building():
    for cfg in ('utdac', 'coco', 'voc'):
        rpn, roi, m = configs.build_hot_path(cfg, train=True)
        assert type(rpn).__name__ == 'ATSSRPNHead' and type(roi).__name__ == 'ProbRoIHead'
        assert type(roi.bbox_roi_extractor).__name__ == 'SingleRoIExtractor'
        assert type(roi.bbox_head).__name__ == 'ProbConvFCBBoxHead'
        assert roi.bbox_roi_extractor.num_inputs == 5
        assert roi.bbox_roi_extractor.roi_layers[0].output_size == (7, 7)
    for name in ('ATSSRPNHead', 'SingleRoIExtractor', 'ProbRoIHead', 'ProbConvFCBBoxHead'):
        assert HEADS.get(name) is not None
    with pytest.raises(KeyError):
        build_head(dict(type='NoSuchHead'))
    with pytest.raises(KeyError):
        build_from_cfg(dict(), HEADS)


def test_state_dict_keys_match_reference_checkpoints():
    rpn, roi, _ = configs.build_hot_path('utdac')
    keys = set(rpn.state_dict())
    assert {'rpn_convs.0.conv.weight', 'rpn_convs.3.gn.weight', 'rpn_cls.weight', 'rpn_reg.bias',
            'rpn_iou.weight', 'scales.4.scale'} <= keys
    assert 'rpn_convs.0.conv.bias' not in keys  # ConvModule drops the bias under GN
    assert set(roi.state_dict()) == {
        f'bbox_head.{n}.{p}' for n in ('shared_fcs.0', 'shared_fcs.1', 'fc_cls', 'fc_reg')
        for p in ('weight', 'bias')}
    assert abs(rpn.rpn_cls.bias[0].item() + 4.59512) < 1e-4  # bias_prob = 0.01
    _, roi_voc, _ = configs.build_hot_path('voc')
    assert any(k.startswith('bbox_head.reg_convs.3.gn') for k in roi_voc.state_dict())
    assert roi_voc.bbox_head.fc_cls.in_features == 1024 and roi_voc.bbox_head.fc_reg.in_features == 256 * 49


def test_configdict_attribute_access_and_nesting():
    c = ConfigDict(nms=dict(type='nms', iou_threshold=0.7), nms_pre=1000)
    assert c.nms.iou_threshold == 0.7 and c['nms_pre'] == 1000
    assert c.get('missing', 3) == 3
    with pytest.raises(AttributeError):
        c.nope
    import copy
    d = copy.deepcopy(c)
    d.nms.iou_threshold = 0.5
    assert c.nms.iou_threshold == 0.7


def test_pad_proposals_and_padded_rois():
    plist = [torch.rand(3, 5), torch.zeros(0, 5), torch.rand(5, 5)]
    pp = roi_head.pad_proposals(plist)
    assert isinstance(pp, PaddedProposals) and pp.boxes.shape == (3, 5, 5)
    assert pp.num.tolist() == [3, 0, 5]
    rois = roi_head.padded_rois(pp)
    assert rois.shape == (15, 5)
    assert rois[:, 0].tolist() == [0, 0, 0, -1, -1] + [-1] * 5 + [2] * 5
    assert torch.equal(rois[10:, 1:], plist[2][:, :4])
    ref = roi_head.bbox2roi(plist)
    assert torch.equal(rois[rois[:, 0] >= 0], ref)


def test_bbox2result_splits_by_label():
    det = np.arange(20, dtype=np.float32).reshape(4, 5)
    lab = np.array([1, 0, 1, 3])
    res = roi_head.bbox2result(det, lab, 4)
    assert [r.shape[0] for r in res] == [1, 2, 0, 1]
    assert all(r.shape == (0, 5) for r in roi_head.bbox2result(np.zeros((0, 5)), lab[:0], 4))


def test_max_iou_assigner_reference_kat():
    # tests/test_utils/test_assigner.py:16-37
    a = sampling.MaxIoUAssigner(pos_iou_thr=0.5, neg_iou_thr=0.5)
    bboxes = torch.FloatTensor([[0, 0, 10, 10], [10, 10, 20, 20], [5, 5, 15, 15], [32, 32, 38, 42]])
    gt = torch.FloatTensor([[0, 0, 10, 9], [0, 10, 10, 19]])
    res = a.assign(bboxes, gt, gt_labels=torch.LongTensor([2, 3]))
    assert res.gt_inds.tolist() == [1, 0, 2, 0]
    assert len(res.labels) == 4
    res = a.assign(bboxes, torch.empty(0, 4), gt_labels=torch.empty(0, dtype=torch.long))
    assert res.gt_inds.tolist() == [0, 0, 0, 0]


def test_random_sampler_row_order_and_prior_indexing():
    torch.manual_seed(0)
    props = torch.cat([torch.rand(200, 2) * 300, torch.rand(200, 2) * 300 + 300, torch.rand(200, 1)], 1)
    gt = torch.tensor([[10., 10., 320., 330.], [100., 120., 500., 480.]])
    # few positives (< num*pos_fraction) so every GT row survives the sampling,
    # the regime prob_roi_head.py:52-57 silently assumes
    a = sampling.MaxIoUAssigner(pos_iou_thr=0.8, neg_iou_thr=0.8, min_pos_iou=0.8, match_low_quality=False)
    s = sampling.RandomSampler(num=64, pos_fraction=0.25, add_gt_as_proposals=True)
    ar = a.assign(props, gt, None, torch.tensor([1, 2]))
    res = s.sample(ar, props, gt, torch.tensor([1, 2]))
    assert res.bboxes.shape[0] == 64 and res.pos_inds.numel() <= 16
    assert res.pos_inds[:2].tolist() == [0, 1]  # GTs first (prob_roi_head.py:52-57)
    assert torch.equal(res.pos_inds, res.pos_inds.sort().values)
    assert torch.equal(res.neg_inds, res.neg_inds.sort().values)
    assert res.pos_is_gt[:2].tolist() == [1, 1]


def test_coder_encode_vs_executed_reference():
    from boosting_rcnn_b200.coder import DeltaXYWHBBoxCoder
    c = DeltaXYWHBBoxCoder(target_stds=(.1, .1, .2, .2))
    rois = torch.from_numpy(G['d2b_rois'])
    out = c.encode(rois, rois.flip(0)).numpy()
    assert np.allclose(out, G['b2d_out'], rtol=1e-6, atol=1e-6)


def test_ops_reject_cpu_tensors():
    with pytest.raises(RuntimeError, match='no CPU path'):
        ops.map_roi_levels(torch.zeros(3, 5), 5)
    with pytest.raises(RuntimeError):
        ops.delta2bbox(torch.zeros(3, 4), torch.zeros(3, 4))
    with pytest.raises(RuntimeError):
        ops.roi_extract([torch.zeros(1, 4, 8, 8)], torch.zeros(2, 5), [0.125])


def test_sample_plan_matches_random_sampler_counts_and_rng_stream():
    """ops.sample_plan (the host half of the fused training front-end) draws exactly the
    permutations RandomSampler.sample draws, in the same order, from the same CPU RNG
    state, and selects the same index sets."""
    rng = np.random.RandomState(0)
    cases = []
    for n_pos, n_neg in ((3, 2000), (300, 1500), (0, 40), (200, 100), (128, 384), (129, 0)):
        gi = torch.cat([torch.ones(n_pos, dtype=torch.long), torch.zeros(n_neg, dtype=torch.long),
                        -torch.ones(7, dtype=torch.long)])
        gi = gi[torch.from_numpy(rng.permutation(len(gi)))]
        cases.append(gi)
    for ub in (-1, 3):
        s = sampling.RandomSampler(num=512, pos_fraction=0.25, neg_pos_ub=ub,
                                   add_gt_as_proposals=False)
        torch.manual_seed(5)
        ref = []
        for gi in cases:
            ar = sampling.AssignResult(0, gi.clone(), gi.new_zeros(len(gi), dtype=torch.float))
            res = s.sample(ar, torch.zeros(len(gi), 4), torch.zeros(0, 4))
            ref.append((res.pos_inds, res.neg_inds))
        torch.manual_seed(5)
        counts = [(int((gi > 0).sum()), int((gi == 0).sum())) for gi in cases]
        plan, pp, pn, rows = ops.sample_plan(counts, 512, 0.25, ub)
        base = 0
        for b, gi in enumerate(cases):
            n_pos, n_neg, first, use_p, use_n = plan[b].tolist()
            assert first == base and rows[b] == n_pos + n_neg
            base += rows[b]
            pos_list = torch.nonzero(gi > 0).squeeze(1)
            neg_list = torch.nonzero(gi == 0).squeeze(1)
            pos = pos_list[pp[b, :n_pos].long()].sort().values if use_p else pos_list
            neg = neg_list[pn[b, :n_neg].long()].sort().values if use_n else neg_list
            assert torch.equal(pos, ref[b][0]) and torch.equal(neg, ref[b][1]), (ub, b)


def test_register_into_mmdet_against_a_stub_registry(monkeypatch):
    """registry.register_into_mmdet / mmdet_plugin (INTEGRATION.md level 1) against a stub of
    mmdet.models.builder exposing mmcv's Registry.register_module(name=, force=, module=)."""
    import importlib
    import sys
    import types

    class StubRegistry:
        def __init__(self):
            self.module_dict = {}

        def register_module(self, name=None, force=False, module=None):
            if name in self.module_dict and not force:
                raise KeyError(name)
            self.module_dict[name] = module
            return module

    heads, extractors = StubRegistry(), StubRegistry()
    heads.module_dict['ATSSRPNHead'] = object          # the reference's own class is replaced
    builder = types.ModuleType('mmdet.models.builder')
    builder.HEADS, builder.ROI_EXTRACTORS = heads, extractors
    models = types.ModuleType('mmdet.models')
    models.builder = builder
    mmdet = types.ModuleType('mmdet')
    mmdet.models = models
    for n, m in (('mmdet', mmdet), ('mmdet.models', models), ('mmdet.models.builder', builder)):
        monkeypatch.setitem(sys.modules, n, m)
    from boosting_rcnn_b200 import bbox_head, registry, roi_extractor, roi_head, rpn_head
    names = registry.register_into_mmdet(force=True)
    assert set(names) == set(registry.HOT_PATH_CLASSES)
    assert heads.module_dict['ATSSRPNHead'] is rpn_head.ATSSRPNHead
    assert heads.module_dict['ProbRoIHead'] is roi_head.ProbRoIHead
    assert heads.module_dict['ProbConvFCBBoxHead'] is bbox_head.ProbConvFCBBoxHead
    assert extractors.module_dict['SingleRoIExtractor'] is roi_extractor.SingleRoIExtractor
    # subset selection through the plugin module's environment variable
    heads.module_dict.clear()
    monkeypatch.setenv('BRCNN_PLUGIN_CLASSES', 'ProbRoIHead,SingleRoIExtractor')
    sys.modules.pop('boosting_rcnn_b200.mmdet_plugin', None)
    plugin = importlib.import_module('boosting_rcnn_b200.mmdet_plugin')
    assert plugin.REGISTERED == ('ProbRoIHead', 'SingleRoIExtractor')
    assert set(heads.module_dict) == {'ProbRoIHead'}
    sys.modules.pop('boosting_rcnn_b200.mmdet_plugin', None)


def test_two_phase_topk_selection_is_a_superset():
    """The RPN top-k pre-selects on an APPROXIMATE score (|error| <= eps << 1/2048) and ranks the
    survivors on the pinned one (csrc/rpn.cuh: rpn_score_kernel / rpn_collect_kernel).  Model of
    that rule in numpy: with d = the 1/2048-wide value bin holding the k-th best approximate
    score, the elements whose approximate score lies in a bin >= d - 1 must contain the exact
    top-k (ties included) for ANY perturbation within eps — random, adversarial around the
    threshold, and with the scores piled onto bin edges."""
    rng = np.random.RandomState(5)
    eps = 2e-5
    for trial in range(60):
        n = int(rng.choice([500, 3000, 40000]))
        k = int(rng.choice([1, 100, 300, n - 1]))
        mode = trial % 3
        if mode == 0:
            s = rng.rand(n)
        elif mode == 1:        # piled onto bin edges
            s = rng.randint(0, 2049, n) / 2048.0 + rng.randint(-3, 4, n) * 6e-8
        else:                  # narrow spread around one edge
            s = 0.5 + rng.randn(n) * 3e-5
        s = np.clip(s, 0.0, 1.0).astype(np.float64)
        if trial % 2:          # adversarial: push the true top-k down, everything else up
            thr = np.sort(s)[::-1][k - 1]
            noise = np.where(s >= thr, -eps, eps)
        else:
            noise = rng.uniform(-eps, eps, n)
        sa = np.clip(s + noise, 0.0, None)
        bins = np.minimum((sa * 2048).astype(np.int64), 2047)
        hist = np.bincount(bins, minlength=2048)
        ge = np.cumsum(hist[::-1])[::-1]                   # count(bin >= d)
        d = int(np.max(np.nonzero(ge >= k)[0]))            # #(bins > d) < k <= #(bins >= d)
        assert (ge[d + 1] if d + 1 < 2048 else 0) < k <= ge[d]
        picked = bins >= max(d - 1, 0)
        kth = np.sort(s)[::-1][k - 1]
        assert picked[s >= kth].all(), (trial, n, k, d)


def test_device_guard_builds_no_reference_cycles(monkeypatch):
    """ops._device_guard must not leave garbage behind: a recursive local closure over the
    first tensor argument once kept every eager step's autograd graph alive until the cyclic
    GC ran, which pinned the parameters' AccumulateGrad nodes to the default stream and broke
    the capture of the training-step graphs."""
    import gc
    import weakref

    # make the guard take its CUDA branch on this CPU-only machine
    def first_tensor(args):
        for a in args:
            if isinstance(a, torch.Tensor):
                return a
            if isinstance(a, (list, tuple)):
                t = first_tensor(a)
                if t is not None:
                    return t
        return None
    monkeypatch.setattr(ops, '_first_cuda_tensor', first_tensor)
    monkeypatch.setattr(torch.cuda, 'current_device', lambda: None)   # == cpu device index

    @ops._device_guard
    def op(x, ys, k=None):
        return x.sum() + ys[0].sum()

    gc.collect()
    gc.disable()
    try:
        w = torch.ones(4, requires_grad=True)
        x = w * 2
        out = op(x, [torch.ones(2)], k=(torch.zeros(1),))
        ref = weakref.ref(x)
        del x, out
        assert ref() is None, 'the first tensor argument outlived its last user reference'
        assert gc.collect() == 0
    finally:
        gc.enable()


def test_host_to_device_without_cuda_is_a_plain_copy():
    t = ops.host_to_device([3, 4], torch.int32, 'cpu')
    assert t.dtype == torch.int32 and t.tolist() == [3, 4]
    t = ops.host_to_device(torch.tensor([1.5]), torch.float32, torch.device('cpu'))
    assert t.tolist() == [1.5]
