"""R-CNN assign + sample in torch ops — the FALLBACK of the training front-end.

The hot path runs ``ops.rcnn_assign_sample`` (brcnn_rcnn_assign + brcnn_rcnn_sample_targets,
csrc/rcnn_train_prep.cuh); these classes are used by
``ProbRoIHead.forward_train`` only for settings the fused kernels do not cover
(match_low_quality, ignore regions, > 2048 GTs per image, other samplers) and as the
reference-shaped API (``build_assigner`` / ``build_sampler``).  Both paths are pinned to the
same executed-reference goldens (tests/golden/make_golden_train.py).

Semantics of mmdet/core/bbox/assigners/max_iou_assigner.py:61-212,
samplers/base_sampler.py:35-102, random_sampler.py:32-82 and
sampling_result.py:26-55, which fix the row order (positives then negatives,
each index-sorted) the prior vector is built against.
"""
import torch

from .registry import BBOX_ASSIGNERS, BBOX_SAMPLERS


def bbox_overlaps(b1, b2, eps=1e-6):
    """IoU matrix (len(b1), len(b2)) — iou2d_calculator.py:75-260, mode='iou'."""
    a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    lt = torch.max(b1[:, None, :2], b2[None, :, :2])
    rb = torch.min(b1[:, None, 2:4], b2[None, :, 2:4])
    wh = (rb - lt).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    union = (a1[:, None] + a2[None, :] - inter).clamp(min=eps)
    return inter / union


class AssignResult:

    def __init__(self, num_gts, gt_inds, max_overlaps, labels=None):
        self.num_gts, self.gt_inds = num_gts, gt_inds
        self.max_overlaps, self.labels = max_overlaps, labels

    def add_gt_(self, gt_labels):
        n = len(gt_labels)
        self.gt_inds = torch.cat([torch.arange(1, n + 1, dtype=torch.long,
                                               device=gt_labels.device), self.gt_inds])
        self.max_overlaps = torch.cat([self.max_overlaps.new_ones(n), self.max_overlaps])
        if self.labels is not None:
            self.labels = torch.cat([gt_labels, self.labels])


@BBOX_ASSIGNERS.register_module()
class MaxIoUAssigner:

    def __init__(self, pos_iou_thr, neg_iou_thr, min_pos_iou=.0, gt_max_assign_all=True,
                 ignore_iof_thr=-1, ignore_wrt_candidates=True, match_low_quality=True,
                 gpu_assign_thr=-1, iou_calculator=None):
        self.pos_iou_thr, self.neg_iou_thr, self.min_pos_iou = pos_iou_thr, neg_iou_thr, min_pos_iou
        self.gt_max_assign_all, self.match_low_quality = gt_max_assign_all, match_low_quality
        self.ignore_iof_thr = ignore_iof_thr

    def assign(self, bboxes, gt_bboxes, gt_bboxes_ignore=None, gt_labels=None):
        if gt_bboxes_ignore is not None and self.ignore_iof_thr > 0:
            raise NotImplementedError('gt_bboxes_ignore is not used by the named configs')
        overlaps = bbox_overlaps(gt_bboxes, bboxes[:, :4])
        k, n = overlaps.shape
        gt_inds = overlaps.new_full((n,), -1, dtype=torch.long)
        if k == 0 or n == 0:
            if k == 0:
                gt_inds[:] = 0
            labels = None if gt_labels is None else overlaps.new_full((n,), -1, dtype=torch.long)
            return AssignResult(k, gt_inds, overlaps.new_zeros((n,)), labels)
        max_ov, argmax_ov = overlaps.max(dim=0)
        gt_max_ov, gt_argmax_ov = overlaps.max(dim=1)
        if isinstance(self.neg_iou_thr, float):
            gt_inds[(max_ov >= 0) & (max_ov < self.neg_iou_thr)] = 0
        else:
            gt_inds[(max_ov >= self.neg_iou_thr[0]) & (max_ov < self.neg_iou_thr[1])] = 0
        pos = max_ov >= self.pos_iou_thr
        gt_inds[pos] = argmax_ov[pos] + 1
        if self.match_low_quality:
            for i in range(k):
                if gt_max_ov[i] >= self.min_pos_iou:
                    if self.gt_max_assign_all:
                        gt_inds[overlaps[i, :] == gt_max_ov[i]] = i + 1
                    else:
                        gt_inds[gt_argmax_ov[i]] = i + 1
        labels = None
        if gt_labels is not None:
            labels = gt_inds.new_full((n,), -1)
            p = torch.nonzero(gt_inds > 0, as_tuple=False).squeeze(1)
            if p.numel() > 0:
                labels[p] = gt_labels[gt_inds[p] - 1]
        return AssignResult(k, gt_inds, max_ov, labels)


class SamplingResult:

    def __init__(self, pos_inds, neg_inds, bboxes, gt_bboxes, assign_result, gt_flags):
        self.pos_inds, self.neg_inds = pos_inds, neg_inds
        self.pos_bboxes, self.neg_bboxes = bboxes[pos_inds], bboxes[neg_inds]
        self.pos_is_gt = gt_flags[pos_inds]
        self.num_gts = gt_bboxes.shape[0]
        self.pos_assigned_gt_inds = assign_result.gt_inds[pos_inds] - 1
        if gt_bboxes.numel() == 0:
            self.pos_gt_bboxes = torch.empty_like(gt_bboxes).view(-1, 4)
        else:
            self.pos_gt_bboxes = gt_bboxes.view(-1, 4)[self.pos_assigned_gt_inds, :]
        self.pos_gt_labels = (assign_result.labels[pos_inds]
                              if assign_result.labels is not None else None)

    @property
    def bboxes(self):
        return torch.cat([self.pos_bboxes, self.neg_bboxes])


@BBOX_SAMPLERS.register_module()
class RandomSampler:

    def __init__(self, num, pos_fraction, neg_pos_ub=-1, add_gt_as_proposals=True, **kwargs):
        self.num, self.pos_fraction = num, pos_fraction
        self.neg_pos_ub, self.add_gt_as_proposals = neg_pos_ub, add_gt_as_proposals

    @staticmethod
    def random_choice(gallery, num):
        # CPU randperm then move, as random_sampler.py:58 does
        perm = torch.randperm(gallery.numel())[:num].to(device=gallery.device)
        return gallery[perm]

    def _pick(self, mask, num_expected):
        inds = torch.nonzero(mask, as_tuple=False).squeeze(1)
        return inds if inds.numel() <= num_expected else self.random_choice(inds, num_expected)

    def sample(self, assign_result, bboxes, gt_bboxes, gt_labels=None, **kwargs):
        bboxes = bboxes[:, :4]
        gt_flags = bboxes.new_zeros((bboxes.shape[0],), dtype=torch.uint8)
        if self.add_gt_as_proposals and len(gt_bboxes) > 0:
            if gt_labels is None:
                raise ValueError('gt_labels must be given when add_gt_as_proposals is True')
            bboxes = torch.cat([gt_bboxes, bboxes], dim=0)
            assign_result.add_gt_(gt_labels)
            gt_flags = torch.cat([bboxes.new_ones(gt_bboxes.shape[0], dtype=torch.uint8), gt_flags])
        pos_inds = self._pick(assign_result.gt_inds > 0, int(self.num * self.pos_fraction)).unique()
        num_neg = self.num - pos_inds.numel()
        if self.neg_pos_ub >= 0:
            num_neg = min(num_neg, int(self.neg_pos_ub * max(1, pos_inds.numel())))
        neg_inds = self._pick(assign_result.gt_inds == 0, num_neg).unique()
        return SamplingResult(pos_inds, neg_inds, bboxes, gt_bboxes, assign_result, gt_flags)
