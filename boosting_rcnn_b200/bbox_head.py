"""ProbConvFCBBoxHead — drop-in for
mmdet/models/roi_heads/bbox_heads/convfc_bbox_head.py:283-451 (and the parts
of ConvFCBBoxHead :13-192 / BBoxHead bbox_head.py:17-253 it inherits).

The FC / conv layers stay on torch/cuBLAS/cuDNN (north star) with the
reference's parameter names (``shared_fcs.N``, ``cls_fcs``, ``reg_fcs``,
``fc_cls``, ``fc_reg``).  B200-native: ``get_bboxes`` (decode + rescale +
multiclass_nms, :294-330) and ``loss`` (:332-418) which run through
``brcnn_rcnn_get_bboxes`` / ``brcnn_boost_loss``.
"""
import torch
import torch.nn as nn
from torch.nn.modules.utils import _pair

from . import ops
from .registry import HEADS, ConfigDict, build_bbox_coder, build_loss
from .rpn_head import ConvModule


@HEADS.register_module()
class ProbConvFCBBoxHead(nn.Module):

    def __init__(self, fc_out_channels=1024, focal_reg=False, gamma=1, num_shared_convs=0,
                 num_shared_fcs=0, num_cls_convs=0, num_cls_fcs=0, num_reg_convs=0,
                 num_reg_fcs=0, conv_out_channels=256, conv_cfg=None, norm_cfg=None,
                 with_avg_pool=False, with_cls=True, with_reg=True, roi_feat_size=7,
                 in_channels=256, num_classes=80,
                 bbox_coder=dict(type='DeltaXYWHBBoxCoder', clip_border=True,
                                 target_means=[0., 0., 0., 0.],
                                 target_stds=[0.1, 0.1, 0.2, 0.2]),
                 reg_class_agnostic=False, reg_decoded_bbox=False,
                 reg_predictor_cfg=dict(type='Linear'), cls_predictor_cfg=dict(type='Linear'),
                 loss_cls=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=1.0),
                 loss_bbox=dict(type='SmoothL1Loss', beta=1.0, loss_weight=1.0),
                 init_cfg=None):
        super().__init__()
        assert with_cls and with_reg, 'the boosting configs use both branches'
        assert (num_shared_convs + num_shared_fcs + num_cls_convs + num_cls_fcs +
                num_reg_convs + num_reg_fcs > 0)
        if num_cls_convs > 0 or num_reg_convs > 0:
            assert num_shared_fcs == 0
        if focal_reg or reg_decoded_bbox:
            raise NotImplementedError('focal_reg / reg_decoded_bbox are not used by the named configs')
        self.focal_reg, self.gamma = focal_reg, gamma
        self.with_avg_pool, self.with_cls, self.with_reg = with_avg_pool, with_cls, with_reg
        self.roi_feat_size = _pair(roi_feat_size)
        self.roi_feat_area = self.roi_feat_size[0] * self.roi_feat_size[1]
        self.in_channels, self.num_classes = in_channels, num_classes
        self.reg_class_agnostic, self.reg_decoded_bbox = reg_class_agnostic, reg_decoded_bbox
        self.conv_out_channels, self.fc_out_channels = conv_out_channels, fc_out_channels
        self.conv_cfg, self.norm_cfg = conv_cfg, norm_cfg
        self.num_shared_convs, self.num_shared_fcs = num_shared_convs, num_shared_fcs
        self.num_cls_convs, self.num_cls_fcs = num_cls_convs, num_cls_fcs
        self.num_reg_convs, self.num_reg_fcs = num_reg_convs, num_reg_fcs
        self.bbox_coder = build_bbox_coder(bbox_coder)
        self.loss_cls = build_loss(loss_cls)
        self.loss_bbox = build_loss(loss_bbox)
        if type(self.loss_cls).__name__ != 'CrossEntropyLoss' or self.loss_cls.use_sigmoid:
            raise NotImplementedError('the boost loss is softmax cross-entropy')
        if type(self.loss_bbox).__name__ != 'L1Loss':
            raise NotImplementedError('the fused loss implements L1Loss (named configs)')
        if with_avg_pool:
            self.avg_pool = nn.AvgPool2d(self.roi_feat_size)

        self.shared_convs, self.shared_fcs, last = self._branch(
            num_shared_convs, num_shared_fcs, in_channels, True)
        self.shared_out_channels = last
        self.cls_convs, self.cls_fcs, self.cls_last_dim = self._branch(
            num_cls_convs, num_cls_fcs, last)
        self.reg_convs, self.reg_fcs, self.reg_last_dim = self._branch(
            num_reg_convs, num_reg_fcs, last)
        if num_shared_fcs == 0 and not with_avg_pool:
            if num_cls_fcs == 0:
                self.cls_last_dim *= self.roi_feat_area
            if num_reg_fcs == 0:
                self.reg_last_dim *= self.roi_feat_area
        self.relu = nn.ReLU(inplace=True)
        self.fc_cls = nn.Linear(self.cls_last_dim, num_classes + 1)
        self.fc_reg = nn.Linear(self.reg_last_dim, 4 if reg_class_agnostic else 4 * num_classes)
        self.init_weights()
        # RoI-feature hand-off: when the first layer after the flatten is
        # shared_fcs.0, its weight lives with columns in (ph, pw, c) order so that the
        # channels-last RoI features ((R,7,7,C) storage) feed it as a free view and the
        # gradient comes back bin-major for the RoIAlign backward.  state_dict() / load
        # convert to / from the reference's (c, ph, pw) order (convfc_bbox_head.py:164), so
        # checkpoints are interchangeable.
        self.fc1_hwc = (num_shared_fcs > 0 and not with_avg_pool and self.roi_feat_area > 1)
        if self.fc1_hwc:
            with torch.no_grad():
                self.shared_fcs[0].weight.copy_(self._fc1_to_hwc(self.shared_fcs[0].weight))
            self._register_state_dict_hook(ProbConvFCBBoxHead._state_dict_hook)
            self._register_load_state_dict_pre_hook(self._load_state_dict_pre_hook)

    def _branch(self, n_convs, n_fcs, in_channels, is_shared=False):
        last = in_channels
        convs = nn.ModuleList()
        for i in range(n_convs):
            convs.append(ConvModule(last if i == 0 else self.conv_out_channels,
                                    self.conv_out_channels, 3, padding=1,
                                    conv_cfg=self.conv_cfg, norm_cfg=self.norm_cfg))
        if n_convs > 0:
            last = self.conv_out_channels
        fcs = nn.ModuleList()
        if n_fcs > 0:
            if (is_shared or self.num_shared_fcs == 0) and not self.with_avg_pool:
                last *= self.roi_feat_area
            for i in range(n_fcs):
                fcs.append(nn.Linear(last if i == 0 else self.fc_out_channels,
                                     self.fc_out_channels))
            last = self.fc_out_channels
        return convs, fcs, last

    # ---- fc1 column order: reference (c, ph, pw) <-> stored (ph, pw, c) ----
    def _fc1_channels(self):
        return self.shared_fcs[0].in_features // self.roi_feat_area

    def _fc1_to_hwc(self, w):
        o = w.size(0)
        return w.reshape(o, self._fc1_channels(), self.roi_feat_area).permute(0, 2, 1).reshape(o, -1)

    def _fc1_to_ref(self, w):
        o = w.size(0)
        return w.reshape(o, self.roi_feat_area, self._fc1_channels()).permute(0, 2, 1).reshape(o, -1)

    @staticmethod
    def _state_dict_hook(module, state_dict, prefix, local_metadata):
        key = prefix + 'shared_fcs.0.weight'
        if key in state_dict:
            state_dict[key] = module._fc1_to_ref(state_dict[key]).contiguous()

    def _load_state_dict_pre_hook(self, state_dict, prefix, *args):
        key = prefix + 'shared_fcs.0.weight'
        if key in state_dict and state_dict[key].dim() == 2 and \
                state_dict[key].size(1) == self.shared_fcs[0].in_features:
            state_dict[key] = self._fc1_to_hwc(state_dict[key]).contiguous()

    def _flatten(self, x):
        """``x.flatten(1)`` of the reference (:164) against the stored column order."""
        if self.fc1_hwc and x.dim() == 4:
            return x.permute(0, 2, 3, 1).reshape(x.size(0), -1)   # a view for channels_last x
        return x.flatten(1)

    def init_weights(self):
        # bbox_head.py:89-101 + convfc_bbox_head.py:97-107
        for group in (self.shared_fcs, self.cls_fcs, self.reg_fcs):
            for m in group:
                nn.init.xavier_uniform_(m.weight)
                nn.init.constant_(m.bias, 0)
        nn.init.normal_(self.fc_cls.weight, 0, 0.01)
        nn.init.constant_(self.fc_cls.bias, 0)
        nn.init.normal_(self.fc_reg.weight, 0, 0.001)
        nn.init.constant_(self.fc_reg.bias, 0)

    def forward(self, x):
        for conv in self.shared_convs:
            x = conv(x)
        if self.num_shared_fcs > 0:
            if self.with_avg_pool:
                x = self.avg_pool(x)
            x = self._flatten(x)
            for fc in self.shared_fcs:
                x = self.relu(fc(x))
        x_cls, x_reg = x, x
        for conv in self.cls_convs:
            x_cls = conv(x_cls)
        if x_cls.dim() > 2:
            if self.with_avg_pool:
                x_cls = self.avg_pool(x_cls)
            x_cls = x_cls.flatten(1)
        for fc in self.cls_fcs:
            x_cls = self.relu(fc(x_cls))
        for conv in self.reg_convs:
            x_reg = conv(x_reg)
        if x_reg.dim() > 2:
            if self.with_avg_pool:
                x_reg = self.avg_pool(x_reg)
            x_reg = x_reg.flatten(1)
        for fc in self.reg_fcs:
            x_reg = self.relu(fc(x_reg))
        return self.fc_cls(x_cls), self.fc_reg(x_reg)

    # ---------------------------------------------------------------- test
    def rcnn_params(self, batch, rois_per_img, cfg, prob, rescale):
        nms = cfg.nms
        if nms.get('type', 'nms') != 'nms':
            raise NotImplementedError(f"nms type {nms.get('type')!r} is outside the hot path")
        return ops.make_rcnn_params(batch, rois_per_img, self.num_classes, cfg.score_thr,
                                    nms.iou_threshold, cfg.max_per_img, self.bbox_coder.means,
                                    self.bbox_coder.stds, self.reg_class_agnostic, prob, rescale)

    def get_bboxes(self, rois, cls_score, bbox_pred, img_shape, scale_factor, rescale=False,
                   cfg=None):
        """Single-image reference signature (:294-330).  ``cls_score`` is used
        as-is (the reference skips the softmax here: fusion already happened)."""
        scores = cls_score
        if cfg is None:
            bboxes = self.bbox_coder.decode(rois[..., 1:], bbox_pred, max_shape=img_shape)
            if rescale and bboxes.size(0) > 0:
                sf = bboxes.new_tensor(scale_factor)
                bboxes = (bboxes.view(bboxes.size(0), -1, 4) / sf).view(bboxes.size(0), -1)
            return bboxes, scores
        cfg = ConfigDict(cfg)
        n = rois.size(0)
        dev = rois.device
        p = self.rcnn_params(1, max(n, 1), cfg, prob=False, rescale=rescale)
        hw = torch.tensor([[img_shape[0], img_shape[1]]], dtype=torch.float32, device=dev)
        sf = torch.as_tensor(scale_factor, dtype=torch.float32).reshape(1, 4).to(dev) if rescale else None
        if n == 0:
            return rois.new_zeros(0, 5), rois.new_zeros((0,), dtype=torch.long)
        rois1 = rois.clone()
        rois1[:, 0] = 0
        det, lab, num = ops.rcnn_get_bboxes(p, rois1, None, torch.tensor([n], dtype=torch.int32, device=dev),
                                            scores, bbox_pred, hw, sf)
        k = int(num.item())
        return det[0, :k], lab[0, :k]

    # --------------------------------------------------------------- train
    def _get_target_single(self, pos_bboxes, neg_bboxes, pos_gt_bboxes, pos_gt_labels, cfg):
        """bbox_head.py:122-186 (host glue, torch)."""
        num_pos, num_neg = pos_bboxes.size(0), neg_bboxes.size(0)
        num_samples = num_pos + num_neg
        labels = pos_bboxes.new_full((num_samples,), self.num_classes, dtype=torch.long)
        label_weights = pos_bboxes.new_zeros(num_samples)
        bbox_targets = pos_bboxes.new_zeros(num_samples, 4)
        bbox_weights = pos_bboxes.new_zeros(num_samples, 4)
        if num_pos > 0:
            labels[:num_pos] = pos_gt_labels
            label_weights[:num_pos] = 1.0 if cfg.pos_weight <= 0 else cfg.pos_weight
            bbox_targets[:num_pos, :] = self.bbox_coder.encode(pos_bboxes, pos_gt_bboxes)
            bbox_weights[:num_pos, :] = 1
        if num_neg > 0:
            label_weights[-num_neg:] = 1.0
        return labels, label_weights, bbox_targets, bbox_weights

    def get_targets(self, sampling_results, gt_bboxes, gt_labels, rcnn_train_cfg, concat=True):
        outs = [self._get_target_single(r.pos_bboxes, r.neg_bboxes, r.pos_gt_bboxes,
                                        r.pos_gt_labels, rcnn_train_cfg) for r in sampling_results]
        labels, label_weights, bbox_targets, bbox_weights = map(list, zip(*outs))
        if concat:
            return (torch.cat(labels, 0), torch.cat(label_weights, 0),
                    torch.cat(bbox_targets, 0), torch.cat(bbox_weights, 0))
        return labels, label_weights, bbox_targets, bbox_weights

    def boost_loss(self, cls_score, bbox_pred, labels, label_weights, bbox_targets,
                   bbox_weights, prior, gamma, alpha=0, reg_norm='bbox_num'):
        """``loss(..., reduction_override='none')`` + ``norm_loss`` + the
        ``loss_bbox`` normalisation of prob_roi_head.py:137-148, fused."""
        loss_cls, loss_bbox, acc, _ = ops.boost_loss(
            cls_score, bbox_pred, labels, label_weights, prior, bbox_targets, bbox_weights,
            self.num_classes, self.reg_class_agnostic, gamma, alpha,
            self.loss_cls.loss_weight, self.loss_bbox.loss_weight, reg_norm == 'mean')
        # acc is a view of the kernel's scalar block: detached, so that it does not inherit
        # requires_grad from its (differentiable) base
        return dict(loss_cls=loss_cls, acc=acc.detach(), loss_bbox=loss_bbox)
