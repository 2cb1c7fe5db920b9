"""DeltaXYWHBBoxCoder and the loss-config holders the reference configs name.

Host-side mirror of mmdet/core/bbox/coder/delta_xywh_bbox_coder.py:10-95:
``decode`` runs the library's delta2bbox kernel (the fused RPN / R-CNN kernels
read ``means`` / ``stds`` straight from this object); ``encode`` (bbox2delta,
:98-141) is target-side host glue outside the hot path and stays in torch.
"""
import numpy as np
import torch
import torch.nn as nn

from . import ops
from .anchors import AnchorGenerator
from .registry import ANCHOR_GENERATORS, BBOX_CODERS, LOSSES

ANCHOR_GENERATORS.register_module(module=AnchorGenerator)


@BBOX_CODERS.register_module()
class DeltaXYWHBBoxCoder:

    def __init__(self, target_means=(0., 0., 0., 0.), target_stds=(1., 1., 1., 1.),
                 clip_border=True, add_ctr_clamp=False, ctr_clamp=32):
        if add_ctr_clamp:
            raise NotImplementedError('add_ctr_clamp (YOLOF) is outside the hot path')
        self.means = tuple(float(v) for v in target_means)
        self.stds = tuple(float(v) for v in target_stds)
        self.clip_border = clip_border
        self.add_ctr_clamp = add_ctr_clamp
        self.ctr_clamp = ctr_clamp

    def decode(self, bboxes, pred_bboxes, max_shape=None, wh_ratio_clip=16 / 1000):
        assert pred_bboxes.size(0) == bboxes.size(0)
        if pred_bboxes.dim() != 2:
            raise NotImplementedError('batched (B,N,4) decode is outside the hot path')
        return ops.delta2bbox(bboxes, pred_bboxes, self.means, self.stds, max_shape,
                              wh_ratio_clip, self.clip_border)

    def encode(self, bboxes, gt_bboxes):
        """bbox2delta (delta_xywh_bbox_coder.py:98-141), torch ops."""
        assert bboxes.size(0) == gt_bboxes.size(0)
        assert bboxes.size(-1) == gt_bboxes.size(-1) == 4
        proposals, gt = bboxes.float(), gt_bboxes.float()
        px = (proposals[..., 0] + proposals[..., 2]) * 0.5
        py = (proposals[..., 1] + proposals[..., 3]) * 0.5
        pw = proposals[..., 2] - proposals[..., 0]
        ph = proposals[..., 3] - proposals[..., 1]
        gx = (gt[..., 0] + gt[..., 2]) * 0.5
        gy = (gt[..., 1] + gt[..., 3]) * 0.5
        gw = gt[..., 2] - gt[..., 0]
        gh = gt[..., 3] - gt[..., 1]
        deltas = torch.stack([(gx - px) / pw, (gy - py) / ph, torch.log(gw / pw),
                              torch.log(gh / ph)], dim=-1)
        means = deltas.new_tensor(self.means).unsqueeze(0)
        stds = deltas.new_tensor(self.stds).unsqueeze(0)
        return deltas.sub_(means).div_(stds)


class _LossSpec(nn.Module):
    """Holds a loss config (type, weights, flags).  The fused kernels read
    ``loss_weight`` / ``use_sigmoid`` from it; losses whose arithmetic is not
    on the hot path (the RPN-side Focal/IoU/MSE losses)
    refuse to be called."""

    def __init__(self, **kwargs):
        super().__init__()
        self.cfg = dict(kwargs)
        self.loss_weight = float(kwargs.get('loss_weight', 1.0))
        self.use_sigmoid = bool(kwargs.get('use_sigmoid', False))
        self.reduction = kwargs.get('reduction', 'mean')

    def forward(self, *args, **kwargs):
        raise NotImplementedError(
            f'{type(self).__name__} is not computed by boosting_rcnn_b200: the '
            'R-CNN losses are fused into brcnn_boost_loss and the RPN losses '
            'are outside the ported hot path (SURVEY.md §8f).')


for _name in ('CrossEntropyLoss', 'L1Loss', 'SmoothL1Loss', 'FocalLoss', 'VarifocalLoss',
              'IoULoss', 'GIoULoss', 'CIoULoss', 'MSELoss', 'QualityFocalLoss'):
    LOSSES.register_module(name=_name, module=type(_name, (_LossSpec,), {}))
