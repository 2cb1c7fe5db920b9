"""ProbRoIHead — drop-in for mmdet/models/roi_heads/prob_roi_head.py:10-283.

B200-native pieces:
  * ``simple_test_bboxes`` (:206-283): bbox2roi, RoI extraction, the
    prior x class-score fusion (:232-240), per-class decode and class-wise NMS
    run batch-wide on the padded proposal layout with no host sync until the
    final ``bbox2result`` copy;
  * ``_bbox_forward_train_boost`` (:107-149) + ``norm_loss`` (:151-154): one
    fused loss kernel (value, accuracy and gradients);
  * the training front-end (:33-64: MaxIoUAssigner, RandomSampler, get_targets, prior
    vector): two launches for the whole batch (``ops.rcnn_assign_sample``) + the reference's
    own CPU ``torch.randperm``; the torch restatement in sampling.py is the fallback for
    settings / sizes the fused kernels do not cover.
Only ``boost=True, quality=False, ams=False`` (every configs/boosting_rcnn/*
file) is supported; mask branches are out of scope.
"""
import numpy as np
import torch
import torch.nn as nn

from . import ops
from .registry import (HEADS, ConfigDict, build_assigner, build_head, build_roi_extractor,
                       build_sampler)
from .rpn_head import PaddedProposals
from . import sampling  # noqa: F401  (registers MaxIoUAssigner / RandomSampler)


def bbox2roi(bbox_list):
    """mmdet/core/bbox/transforms.py:59-78."""
    rois_list = []
    for img_id, bboxes in enumerate(bbox_list):
        if bboxes.size(0) > 0:
            img_inds = bboxes.new_full((bboxes.size(0), 1), img_id)
            rois_list.append(torch.cat([img_inds, bboxes[:, :4]], dim=-1))
        else:
            rois_list.append(bboxes.new_zeros((0, 5)))
    return torch.cat(rois_list, 0)


def bbox2result(bboxes, labels, num_classes):
    """mmdet/core/bbox/transforms.py:100-117 (numpy in, list of arrays out)."""
    if bboxes.shape[0] == 0:
        return [np.zeros((0, 5), dtype=np.float32) for _ in range(num_classes)]
    if isinstance(bboxes, torch.Tensor):
        bboxes, labels = bboxes.detach().cpu().numpy(), labels.detach().cpu().numpy()
    return [bboxes[labels == i, :] for i in range(num_classes)]


def pad_proposals(proposal_list):
    """list of (n_b, >=4) tensors -> PaddedProposals (cap = max n_b, >= 1)."""
    B = len(proposal_list)
    dev = proposal_list[0].device
    cap = max(1, max(p.size(0) for p in proposal_list))
    boxes = proposal_list[0].new_zeros((B, cap, 5))
    for b, p in enumerate(proposal_list):
        n = p.size(0)
        if n:
            boxes[b, :n, :4] = p[:, :4]
            # the reference takes boxes[:, -1] as the prior, whatever it is
            boxes[b, :n, 4] = p[:, -1]
    num = ops.host_to_device([p.size(0) for p in proposal_list], torch.int32, dev)
    return PaddedProposals(boxes, num)


def padded_rois(padded):
    """(B,cap,5) proposals -> (B*cap,5) RoIs [b,x1,y1,x2,y2], padding rows b=-1."""
    boxes, num = padded.boxes, padded.num
    B, cap = boxes.shape[:2]
    idx = torch.arange(B, device=boxes.device, dtype=boxes.dtype)[:, None].expand(B, cap)
    live = torch.arange(cap, device=boxes.device)[None, :] < num[:, None]
    col = torch.where(live, idx, idx.new_full((), -1.0))
    return torch.cat([col[..., None], boxes[..., :4]], dim=-1).reshape(B * cap, 5)


@HEADS.register_module()
class ProbRoIHead(nn.Module):

    def __init__(self, alpha=0, gamma=0.1, boost=False, prob=True, ams=False, quality=False,
                 iou_gamma=0, reg_norm='bbox_num', bbox_roi_extractor=None, bbox_head=None,
                 mask_roi_extractor=None, mask_head=None, shared_head=None, train_cfg=None,
                 test_cfg=None, pretrained=None, init_cfg=None):
        super().__init__()
        if mask_head is not None or shared_head is not None:
            raise NotImplementedError('mask / shared heads are outside the hot path')
        if ams or quality:
            raise NotImplementedError('ams / quality are not used by the named configs')
        self.alpha, self.gamma, self.boost, self.prob = alpha, gamma, boost, prob
        self.ams, self.quality, self.iou_gamma, self.reg_norm = ams, quality, iou_gamma, reg_norm
        self.train_cfg = ConfigDict(train_cfg) if train_cfg is not None else None
        self.test_cfg = ConfigDict(test_cfg) if test_cfg is not None else None
        self.bbox_roi_extractor = build_roi_extractor(bbox_roi_extractor)
        self.bbox_head = build_head(bbox_head)
        self.bbox_assigner = self.bbox_sampler = None
        if self.train_cfg:
            self.bbox_assigner = build_assigner(self.train_cfg.assigner)
            self.bbox_sampler = build_sampler(self.train_cfg.sampler, context=self)
        self._hw_cache = {}
        #: opt-in: replay the sync-free tail of the training step (sample + targets, RoIAlign,
        #: head, boost loss and their backward) as CUDA graphs (graph.RcnnTrainGraph)
        self.train_graph = False
        self._train_graphs = {}
        self._train_buffers = {}

    with_bbox, with_mask, with_shared_head = True, False, False

    # ------------------------------------------------------------- forward
    def _bbox_forward(self, x, rois):
        bbox_feats = self.bbox_roi_extractor(x[:self.bbox_roi_extractor.num_inputs], rois)
        cls_score, bbox_pred = self.bbox_head(bbox_feats)
        return dict(cls_score=cls_score, bbox_pred=bbox_pred, bbox_feats=bbox_feats)

    # --------------------------------------------------------------- train
    def forward_train(self, x, img_metas, proposal_list, gt_bboxes, gt_labels,
                      gt_bboxes_ignore=None, gt_masks=None, assigned=None):
        """prob_roi_head.py:23-105.  ``assigned`` (optional, not in the reference): the handle
        of an earlier ``assign_async`` call on the same proposals / GTs."""
        num_imgs = len(img_metas)
        if assigned is not None:
            return self._forward_train_fused(x, proposal_list, gt_bboxes, gt_labels, assigned)
        if gt_bboxes_ignore is None:
            gt_bboxes_ignore = [None for _ in range(num_imgs)]
        if self._fused_train_prep_ok(gt_bboxes_ignore):
            cap = proposal_list.boxes.size(1) if isinstance(proposal_list, PaddedProposals) \
                else max([int(p.size(0)) for p in proposal_list] + [1])
            if self._fused_train_prep_fits(cap, gt_bboxes, self.bbox_sampler.num):
                return self._forward_train_fused(x, proposal_list, gt_bboxes, gt_labels)
        if isinstance(proposal_list, PaddedProposals):
            n = proposal_list.num.tolist()
            proposal_list = [proposal_list.boxes[b, :k] for b, k in enumerate(n)]
        sampling_results, priors = [], []
        for i in range(num_imgs):
            assign_result = self.bbox_assigner.assign(proposal_list[i], gt_bboxes[i],
                                                      gt_bboxes_ignore[i], gt_labels[i])
            res = self.bbox_sampler.sample(assign_result, proposal_list[i], gt_bboxes[i],
                                           gt_labels[i])
            sampling_results.append(res)
            # prior extraction, prob_roi_head.py:51-64
            num_gts = assign_result.num_gts
            pos_inds = res.pos_inds[num_gts:] - num_gts
            neg_inds = res.neg_inds - num_gts
            pos_prior = proposal_list[i][pos_inds, -1]
            neg_prior = 1 - proposal_list[i][neg_inds, -1]
            priors.append(torch.cat([pos_prior.new_zeros(num_gts), pos_prior, neg_prior]).detach())
        priors = torch.cat(priors, dim=0)
        if not self.boost:
            raise NotImplementedError('boost=False is not used by the named configs')
        bbox_results = self._bbox_forward_train_boost(x, sampling_results, gt_bboxes, gt_labels,
                                                      img_metas, priors)
        return dict(bbox_results['loss_bbox'])

    def _fused_train_prep_ok(self, gt_bboxes_ignore):
        """The two-launch assign+sample+targets path (ops.rcnn_assign_sample) covers the
        train_cfg of every boosting_rcnn config: MaxIoUAssigner(match_low_quality=False,
        scalar neg_iou_thr) + RandomSampler(add_gt_as_proposals=True), no ignore regions."""
        from .sampling import MaxIoUAssigner, RandomSampler
        a, s = self.bbox_assigner, self.bbox_sampler
        return (self.boost and type(a) is MaxIoUAssigner and type(s) is RandomSampler
                and not a.match_low_quality and isinstance(a.neg_iou_thr, float)
                and s.add_gt_as_proposals and all(g is None for g in gt_bboxes_ignore)
                and not getattr(self, 'force_python_train_prep', False))

    def _fused_train_prep_fits(self, num_props_cap, gt_bboxes, num):
        """Limits of the fused kernels: <= 2048 GTs per image (brcnn_rcnn_assign) and a
        candidate list + selection buffer within 200 KB of shared memory
        (brcnn_rcnn_sample_targets); crowded images take the torch fallback instead of
        raising in the middle of training."""
        gmax = max([int(g.size(0)) for g in gt_bboxes] + [1])
        if self.train_graph:
            gmax = (gmax + 31) // 32 * 32      # _gt_capacity
        sel = 1
        while sel < max(1, num):
            sel <<= 1
        return gmax <= 2048 and sel * 8 + (gmax + num_props_cap) * 4 <= 200 * 1024

    def assign_async(self, proposal_list, gt_bboxes, gt_labels, gt_bboxes_ignore=None):
        """Optional first half of ``forward_train``: launch the batch-wide MaxIoU assignment
        now and return a handle for ``forward_train(..., assigned=handle)``.  GPU work queued
        between the two calls (the RPN loss) overlaps the reference's CPU ``randperm``.
        Returns None when the fused path does not cover the settings (``forward_train`` then
        does everything itself)."""
        ignore = gt_bboxes_ignore if gt_bboxes_ignore is not None else [None] * len(gt_bboxes)
        if not self._fused_train_prep_ok(ignore):
            return None
        props = proposal_list if isinstance(proposal_list, PaddedProposals) \
            else pad_proposals(proposal_list)
        if not self._fused_train_prep_fits(props.boxes.size(1), gt_bboxes, self.bbox_sampler.num):
            return None
        return self._assign(props, gt_bboxes, gt_labels)

    def _assign(self, props, gt_bboxes, gt_labels):
        """Batch-wide assignment.  Captured training step: the GT capacity is the next multiple
        of 32 (the graph's shapes repeat from step to step) and the kernels work in a buffer
        set at fixed addresses (nothing is copied between the assignment and the replay)."""
        a = self.bbox_assigner
        cap, static = None, None
        if self.train_graph:
            g = max([int(t.size(0)) for t in gt_bboxes] + [1])
            cap = (g + 31) // 32 * 32
            key = (tuple(props.boxes.shape[:2]), cap, str(props.boxes.device))
            static = self._train_buffers.get(key)
            if static is None:
                static = self._train_buffers[key] = ops.rcnn_assign_buffers(
                    props.boxes.size(0), props.boxes.size(1), cap, props.boxes.device)
        return ops.rcnn_assign(props.boxes, props.num, gt_bboxes, gt_labels, a.pos_iou_thr,
                               a.neg_iou_thr, a.min_pos_iou, max_gts=cap, static=static)

    def _forward_train_fused(self, x, proposal_list, gt_bboxes, gt_labels, assigned=None):
        a, s, h = self.bbox_assigner, self.bbox_sampler, self.bbox_head
        if assigned is None:
            props = proposal_list if isinstance(proposal_list, PaddedProposals) \
                else pad_proposals(proposal_list)
            assigned = self._assign(props, gt_bboxes, gt_labels)
        plan, perm_pos, perm_neg, rows = ops.sample_plan(assigned.counts(), s.num,
                                                         s.pos_fraction, s.neg_pos_ub)
        if self.train_graph:
            from .graph import RcnnTrainGraph
            out = RcnnTrainGraph.run(self, x, assigned, plan, perm_pos, perm_neg, sum(rows))
            if out is not None:
                return out
        dev = assigned.proposals.device
        to = lambda t: ops.host_to_device(t, torch.int32, dev)
        rois, labels, label_weights, bbox_targets, bbox_weights, prior = ops.rcnn_sample_targets(
            assigned.proposals, assigned.num_props, assigned.gtb, assigned.gtl, assigned.num_gt,
            assigned.gt_inds, to(plan), to(perm_pos), to(perm_neg), sum(rows), h.num_classes,
            h.bbox_coder.means, h.bbox_coder.stds, self.train_cfg.pos_weight)
        bbox_results = self._bbox_forward(x, rois)
        return h.boost_loss(bbox_results['cls_score'], bbox_results['bbox_pred'], labels,
                            label_weights, bbox_targets, bbox_weights, prior, self.gamma,
                            self.alpha, self.reg_norm)

    def _bbox_forward_train_boost(self, x, sampling_results, gt_bboxes, gt_labels, img_metas,
                                  priors, ious=None):
        rois = bbox2roi([res.bboxes for res in sampling_results])
        bbox_results = self._bbox_forward(x, rois)
        labels, label_weights, bbox_targets, bbox_weights = self.bbox_head.get_targets(
            sampling_results, gt_bboxes, gt_labels, self.train_cfg)
        loss_bbox = self.bbox_head.boost_loss(bbox_results['cls_score'], bbox_results['bbox_pred'],
                                              labels, label_weights, bbox_targets, bbox_weights,
                                              priors, self.gamma, self.alpha, self.reg_norm)
        bbox_results.update(loss_bbox=loss_bbox)
        return bbox_results

    def norm_loss(self, loss, weights, avg_factor):
        """prob_roi_head.py:151-154 (kept for API parity; the fused kernel
        computes the same quantity)."""
        new_weights = weights * (loss.sum() / (weights * loss).sum())
        return (loss * new_weights.detach()).sum() / avg_factor

    # ---------------------------------------------------------------- test
    def _img_consts(self, img_metas, device):
        key = (str(device), tuple(tuple(m['img_shape'][:2]) for m in img_metas),
               tuple(tuple(np.asarray(m['scale_factor'], dtype=np.float32).reshape(-1).tolist())
                     for m in img_metas))
        c = self._hw_cache.get(key)
        if c is None:
            hw = torch.tensor([list(s) for s in key[1]], dtype=torch.float32).to(device)
            sf = torch.tensor([list(s) for s in key[2]], dtype=torch.float32).to(device)
            c = self._hw_cache[key] = (hw, sf)
            if len(self._hw_cache) > 64:
                self._hw_cache.pop(next(iter(self._hw_cache)))
        return c

    def simple_test_bboxes_padded(self, x, img_metas, proposals, rcnn_test_cfg, rescale=False):
        """Batch-wide test path on device-resident padded proposals.  Returns
        (det_bboxes (B,M,5), det_labels (B,M), num_dets (B,)) on the device."""
        if not isinstance(proposals, PaddedProposals):
            proposals = pad_proposals(proposals)
        B, cap = proposals.boxes.shape[:2]
        rois, prior = ops.bbox2roi_padded(proposals.boxes, proposals.num)
        bbox_results = self._bbox_forward(x, rois)
        hw, sf = self._img_consts(img_metas, rois.device)
        p = self.bbox_head.rcnn_params(B, cap, ConfigDict(rcnn_test_cfg), self.prob, rescale)
        return ops.rcnn_get_bboxes(p, rois, prior, proposals.num, bbox_results['cls_score'],
                                   bbox_results['bbox_pred'], hw, sf if rescale else None)

    def simple_test_bboxes(self, x, img_metas, proposals, rcnn_test_cfg, rescale=False):
        """Reference signature (:206-283): per-image lists of (k,5) / (k,)."""
        if not isinstance(proposals, PaddedProposals):
            if sum(p.size(0) for p in proposals) == 0:
                # no proposal in the whole batch (:216-225)
                det_bbox = proposals[0].new_zeros(0, 5)
                det_label = proposals[0].new_zeros((0,), dtype=torch.long)
                if rcnn_test_cfg is None:
                    det_bbox = det_bbox[:, :4]
                    det_label = proposals[0].new_zeros((0, self.bbox_head.fc_cls.out_features))
                return [det_bbox] * len(proposals), [det_label] * len(proposals)
        if rcnn_test_cfg is None:
            return self._simple_test_bboxes_raw(x, img_metas, proposals, rescale)
        det, lab, num = self.simple_test_bboxes_padded(x, img_metas, proposals, rcnn_test_cfg,
                                                       rescale)
        n = num.tolist()
        return [det[b, :k] for b, k in enumerate(n)], [lab[b, :k] for b, k in enumerate(n)]

    def _simple_test_bboxes_raw(self, x, img_metas, proposals, rescale):
        """rcnn_test_cfg is None: decoded boxes (n,4C) and fused scores (n,C+1) per image
        without NMS (convfc_bbox_head.py:323-324).  Same fusion + decode kernel as the normal
        test path: its intermediates are read back from the workspace of
        ``brcnn_rcnn_get_bboxes`` (the class NMS behind them is run with a 1-detection budget
        and ignored)."""
        if not isinstance(proposals, PaddedProposals):
            proposals = pad_proposals(proposals)
        B, cap = proposals.boxes.shape[:2]
        rois, prior = ops.bbox2roi_padded(proposals.boxes, proposals.num)
        res = self._bbox_forward(x, rois)
        hw, sf = self._img_consts(img_metas, rois.device)
        h = self.bbox_head
        p = ops.make_rcnn_params(B, cap, h.num_classes, 0.05, 0.5, 1, h.bbox_coder.means,
                                 h.bbox_coder.stds, h.reg_class_agnostic, self.prob, rescale)
        _, _, _, ws = ops.rcnn_get_bboxes(p, rois, prior, proposals.num, res['cls_score'],
                                          res['bbox_pred'], hw, sf if rescale else None,
                                          return_workspace=True)
        lay = ops.rcnn_workspace_layout(p)
        C = h.num_classes
        nbox = 1 if h.reg_class_agnostic else C
        scores = ws[lay.scores:lay.scores + B * cap * (C + 1) * 4].view(torch.float32).view(B, cap, C + 1)
        bboxes = ws[lay.bboxes:lay.bboxes + B * cap * nbox * 16].view(torch.float32).view(B, cap, nbox * 4)
        n = proposals.num.tolist()
        return ([bboxes[b, :k].clone() for b, k in enumerate(n)],
                [scores[b, :k].clone() for b, k in enumerate(n)])

    def simple_test(self, x, proposal_list, img_metas, proposals=None, rescale=False):
        det, lab, num = self.simple_test_bboxes_padded(x, img_metas, proposal_list, self.test_cfg,
                                                       rescale)
        # one D2H copy for the whole batch, then the per-class numpy split
        det, lab, num = det.cpu().numpy(), lab.cpu().numpy(), num.cpu().numpy()
        return [bbox2result(det[b, :num[b]], lab[b, :num[b]], self.bbox_head.num_classes)
                for b in range(det.shape[0])]
