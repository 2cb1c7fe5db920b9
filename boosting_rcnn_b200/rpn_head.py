"""ATSSRPNHead ("RetinaRPN") — drop-in for
mmdet/models/dense_heads/atss_rpn_head.py:109-783.

What is B200-native here: ``get_bboxes`` (:466-503, :688-760).  One
``brcnn_rpn_get_bboxes`` call handles the whole batch and every pyramid level
(score, per-level top-k, anchor + delta decode, size filter, batched NMS,
top ``max_per_img``) — no Python loop over images/levels, no host sync.
The conv tower (4x conv3x3+GN+ReLU, rpn_cls/rpn_reg/rpn_iou, per-level Scale)
stays on torch/cuDNN as the north star says; parameter names match the
reference so its checkpoints load (``rpn_convs.N.conv/gn``, ``rpn_cls``,
``rpn_reg``, ``rpn_iou``, ``scales.N.scale``).

``loss`` (:299-464, with ``get_targets`` :505-603 and the AnchorHead target code it
dispatches to when ``atss=False``) is B200-native as well: anchor targets (MaxIoUAssigner with
match_low_quality + PseudoSampler), sigmoid focal loss, IoU-log + MSE regression losses and
the BCE IoU branch run in three launches (``brcnn_rpn_loss_forward``) with hand-derived
gradients; the two ``reduce_mean(...).item()`` normalisers become one device-side fused
all-reduce.  FocalLoss (UTDAC / COCO configs) and VarifocalLoss (VOC config) are covered;
unsupported variants (atss=True, GHM, reg_decoded_bbox=False) raise NotImplementedError at the
first ``loss`` call.
"""
import math
from collections import namedtuple

import torch
import torch.nn as nn

from . import ops
from .registry import (HEADS, ConfigDict, build_anchor_generator, build_bbox_coder,
                       build_loss)

PaddedProposals = namedtuple('PaddedProposals', ['boxes', 'num'])
PaddedProposals.__doc__ = """Fixed-capacity proposal batch that stays on the
device: ``boxes`` (B, max_per_img, 5) zero padded [x1,y1,x2,y2,prior],
``num`` (B,) int32.  ``to_list()``-style conversion is ``unpad_proposals``."""


def unpad_proposals(padded):
    """-> list of (n_b, 5) tensors like the reference (one host sync)."""
    num = padded.num.tolist()
    return [padded.boxes[b, :n] for b, n in enumerate(num)]


class Scale(nn.Module):
    """mmcv.cnn.Scale: one learnable scalar named ``scale``."""

    def __init__(self, scale=1.0):
        super().__init__()
        self.scale = nn.Parameter(torch.tensor(scale, dtype=torch.float))

    def forward(self, x):
        return x * self.scale


class ConvModule(nn.Module):
    """conv -> norm -> ReLU with mmcv.cnn.ConvModule's sub-module names
    (``conv``, ``gn``/``bn``, ``activate``)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0,
                 conv_cfg=None, norm_cfg=None, act=True):
        super().__init__()
        if conv_cfg not in (None, dict(type='Conv2d')):
            raise NotImplementedError(f'conv_cfg {conv_cfg} is outside the hot path')
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride,
                              padding=padding, bias=norm_cfg is None)
        self.norm_name = None
        if norm_cfg is not None:
            kind = norm_cfg.get('type')
            if kind == 'GN':
                self.norm_name = 'gn'
                self.add_module('gn', nn.GroupNorm(norm_cfg.get('num_groups', 32), out_channels))
            elif kind == 'BN':
                self.norm_name = 'bn'
                self.add_module('bn', nn.BatchNorm2d(out_channels))
            else:
                raise NotImplementedError(f'norm {kind}')
            for prm in getattr(self, self.norm_name).parameters():
                prm.requires_grad = norm_cfg.get('requires_grad', True)
        self.activate = nn.ReLU(inplace=True) if act else None

    def forward(self, x):
        x = self.conv(x)
        if self.norm_name is not None:
            x = getattr(self, self.norm_name)(x)
        if self.activate is not None:
            x = self.activate(x)
        return x


@HEADS.register_module()
class ATSSRPNHead(nn.Module):

    def __init__(self, in_channels, num_classes=1, feat_channels=256, stacked_convs=4,
                 conv_cfg=None, gamma=1, atss=False, bridge=False, last_conv='norm',
                 aug_reg_loss=None,
                 norm_cfg=dict(type='GN', num_groups=32, requires_grad=True),
                 anchor_generator=dict(type='AnchorGenerator', scales=[8, 16, 32],
                                       ratios=[0.5, 1.0, 2.0], strides=[4, 8, 16, 32, 64]),
                 bbox_coder=dict(type='DeltaXYWHBBoxCoder', clip_border=True,
                                 target_means=(.0, .0, .0, .0),
                                 target_stds=(1.0, 1.0, 1.0, 1.0)),
                 reg_decoded_bbox=False,
                 loss_cls=dict(type='CrossEntropyLoss', use_sigmoid=True, loss_weight=1.0),
                 loss_bbox=dict(type='SmoothL1Loss', beta=1.0 / 9.0, loss_weight=1.0),
                 loss_centerness=dict(type='CrossEntropyLoss', use_sigmoid=True,
                                      loss_weight=0.5),
                 train_cfg=None, test_cfg=None, init_cfg=None, num_convs=1):
        super().__init__()
        if last_conv != 'norm':
            raise NotImplementedError("last_conv='dcn'/'aspp' is not used by the named configs")
        if bridge:
            raise NotImplementedError('bridge=True is not used by the named configs')
        self.in_channels, self.num_classes, self.feat_channels = in_channels, num_classes, feat_channels
        self.stacked_convs, self.conv_cfg, self.norm_cfg = stacked_convs, conv_cfg, norm_cfg
        self.gamma, self.atss, self.bridge, self.last_conv = gamma, atss, bridge, last_conv
        self.use_sigmoid_cls = loss_cls.get('use_sigmoid', False)
        if not self.use_sigmoid_cls:
            raise NotImplementedError('softmax RPN scores are not used by the named configs')
        self.cls_out_channels = num_classes
        self.reg_decoded_bbox = reg_decoded_bbox
        self.bbox_coder = build_bbox_coder(bbox_coder)
        self.loss_cls = build_loss(loss_cls)
        self.loss_bbox = build_loss(loss_bbox)
        self.loss_centerness = build_loss(loss_centerness)
        self.with_aug_loss = aug_reg_loss is not None
        if self.with_aug_loss:
            self.aug_loss = build_loss(aug_reg_loss)
        self.train_cfg = ConfigDict(train_cfg) if train_cfg is not None else None
        self.test_cfg = ConfigDict(test_cfg) if test_cfg is not None else None
        self.anchor_generator = build_anchor_generator(anchor_generator)
        self.num_anchors = self.anchor_generator.num_base_anchors[0]
        self._init_layers()
        self.init_weights()
        self._const_cache = {}

    # ---------------------------------------------------------------- layers
    def _init_layers(self):
        self.rpn_convs = nn.ModuleList()
        for i in range(self.stacked_convs):
            chn = self.in_channels if i == 0 else self.feat_channels
            self.rpn_convs.append(ConvModule(chn, self.feat_channels, 3, stride=1, padding=1,
                                             conv_cfg=self.conv_cfg, norm_cfg=self.norm_cfg))
        self.rpn_cls = nn.Conv2d(self.feat_channels, self.num_anchors * self.cls_out_channels,
                                 3, padding=1)
        self.rpn_reg = nn.Conv2d(self.feat_channels, self.num_anchors * 4, 3, padding=1)
        self.rpn_iou = nn.Conv2d(self.feat_channels, self.num_anchors * 1, 3, padding=1)
        self.scales = nn.ModuleList([Scale(1.0) for _ in self.anchor_generator.strides])

    def init_weights(self):
        """init_cfg of atss_rpn_head.py:123-131: Normal(std=.01) on every
        Conv2d, rpn_cls bias from bias_prob=0.01."""
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.normal_(m.weight, 0, 0.01)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
        nn.init.constant_(self.rpn_cls.bias, float(-math.log((1 - 0.01) / 0.01)))

    def forward_single(self, x, scale):
        for conv in self.rpn_convs:
            x = conv(x)
        rpn_cls_score = self.rpn_cls(x)
        rpn_bbox_pred = scale(self.rpn_reg(x)).float()
        rpn_iou_pred = self.rpn_iou(x)
        return rpn_cls_score, rpn_bbox_pred, rpn_iou_pred

    def forward(self, feats):
        outs = [self.forward_single(x, s) for x, s in zip(feats, self.scales)]
        return tuple(map(list, zip(*outs)))

    # ------------------------------------------------------------- proposals
    def _constants(self, device, img_shapes):
        key = (str(device), tuple(img_shapes))
        c = self._const_cache.get(key)
        if c is None:
            base = self.anchor_generator.base_anchor_table().to(device)
            hw = torch.tensor([[s[0], s[1]] for s in img_shapes], dtype=torch.float32).to(device)
            c = self._const_cache[key] = (base, hw)
            if len(self._const_cache) > 64:
                self._const_cache.pop(next(iter(self._const_cache)))
        return c

    def get_bboxes_padded(self, cls_scores, bbox_preds, iou_preds, img_metas, cfg=None):
        """Whole-batch proposal generation; results stay on the device."""
        assert len(cls_scores) == len(bbox_preds) == len(iou_preds)
        cfg = self.test_cfg if cfg is None else ConfigDict(cfg)
        nms = cfg.nms
        if nms.get('type', 'nms') != 'nms':
            raise NotImplementedError(f"nms type {nms.get('type')!r} is outside the hot path")
        B = cls_scores[0].size(0)
        assert len(img_metas) == B
        sizes = [tuple(t.shape[-2:]) for t in cls_scores]
        img_shapes = [tuple(m['img_shape'][:2]) for m in img_metas]
        base, hw = self._constants(cls_scores[0].device, img_shapes)
        p = ops.make_rpn_params(B, sizes, self.anchor_generator.strides, self.num_anchors,
                                cfg.nms_pre, cfg.max_per_img, nms.iou_threshold,
                                cfg.min_bbox_size, self.bbox_coder.means, self.bbox_coder.stds)
        boxes, num = ops.rpn_get_bboxes(p, [t.detach() for t in cls_scores],
                                        [t.detach() for t in bbox_preds],
                                        [t.detach() for t in iou_preds], base, hw)
        return PaddedProposals(boxes, num)

    def get_bboxes(self, cls_scores, bbox_preds, iou_preds, img_metas, cfg=None,
                   rescale=False, with_nms=True):
        """Reference signature (:466-503): list of (n,5) proposals per image.
        ``rescale`` is accepted and ignored exactly like the reference."""
        assert with_nms, '``with_nms`` in RPNHead should always True'
        return unpad_proposals(self.get_bboxes_padded(cls_scores, bbox_preds, iou_preds,
                                                      img_metas, cfg))

    def simple_test_rpn(self, x, img_metas, padded=False):
        cls_scores, bbox_preds, iou_preds = self(x)
        if padded:
            return self.get_bboxes_padded(cls_scores, bbox_preds, iou_preds, img_metas)
        return self.get_bboxes(cls_scores, bbox_preds, iou_preds, img_metas)

    # ------------------------------------------------------------------ loss
    def _loss_supported(self):
        tc = self.train_cfg
        a = tc.get('assigner', {}) if tc is not None else {}
        problems = []
        if self.atss:
            problems.append('atss=True (ATSSAssigner)')
        if not self.reg_decoded_bbox:
            problems.append('reg_decoded_bbox=False')
        if self.num_classes != 1:
            problems.append('num_classes != 1')
        for name, want in (('loss_cls', ('FocalLoss', 'VarifocalLoss')), ('loss_bbox', ('IoULoss',)),
                           ('loss_centerness', ('CrossEntropyLoss',))):
            if type(getattr(self, name)).__name__ not in want:
                problems.append(f'{name}={type(getattr(self, name)).__name__}')
        if type(self.loss_cls).__name__ == 'VarifocalLoss' and \
                not self.loss_cls.cfg.get('iou_weighted', True):
            problems.append('VarifocalLoss(iou_weighted=False)')
        if self.with_aug_loss and type(self.aug_loss).__name__ != 'MSELoss':
            problems.append(f'aug_reg_loss={type(self.aug_loss).__name__}')
        if self.loss_bbox.cfg.get('mode', 'log') != 'log' or self.loss_bbox.cfg.get('linear', False):
            problems.append('IoULoss mode != log')
        if tc is None:
            problems.append('train_cfg is None')
        else:
            if a.get('type') != 'MaxIoUAssigner' or not a.get('match_low_quality', True) \
                    or not a.get('gt_max_assign_all', True) or a.get('ignore_iof_thr', -1) > 0 \
                    or not isinstance(a.get('neg_iou_thr'), (int, float)):
                problems.append(f'assigner {dict(a)}')
            if tc.get('sampler', {}).get('type') != 'PseudoSampler':
                problems.append(f"sampler {tc.get('sampler')}")
            if tc.get('allowed_border', -1) >= 0 or tc.get('pos_weight', -1) > 0:
                problems.append('allowed_border >= 0 / pos_weight > 0')
        if tuple(self.bbox_coder.means) != (0., 0., 0., 0.) or \
                tuple(self.bbox_coder.stds) != (1., 1., 1., 1.):
            problems.append('bbox_coder means/stds other than 0/1')
        return problems

    def loss(self, cls_scores, bbox_preds, iou_preds, gt_bboxes, img_metas, gt_bboxes_ignore=None):
        """Reference signature (:405-464).  Returns dict(loss_rpn_cls, loss_rpn_bbox,
        loss_rpn_iou), each a list with one 0-dim tensor per pyramid level."""
        problems = self._loss_supported()
        if gt_bboxes_ignore is not None and any(g is not None and len(g) for g in gt_bboxes_ignore):
            problems.append('gt_bboxes_ignore')
        if problems:
            raise NotImplementedError('the fused RPN loss covers the ATSSRPNHead settings of '
                                      'configs/boosting_rcnn/*; unsupported: '
                                      + '; '.join(problems))
        B = cls_scores[0].size(0)
        assert len(img_metas) == B and len(gt_bboxes) == B
        dev = cls_scores[0].device
        sizes = [tuple(t.shape[-2:]) for t in cls_scores]
        assert len(sizes) == len(self.anchor_generator.strides)
        img_shapes = [tuple(m['img_shape'][:2]) for m in img_metas]
        base, _ = self._constants(dev, img_shapes)
        Gs = [int(g.size(0)) for g in gt_bboxes]
        Gmax = max(1, max(Gs))
        gtb = torch.zeros((B, Gmax, 4), dtype=torch.float32, device=dev)
        for b in range(B):
            if Gs[b]:
                gtb[b, :Gs[b]] = gt_bboxes[b][:, :4].float()
        num_gt = ops.host_to_device(Gs, torch.int32, dev)
        pad_hw = ops.host_to_device([[m['pad_shape'][0], m['pad_shape'][1]] for m in img_metas],
                                    torch.float32, dev)
        a = self.train_cfg.assigner
        varifocal = type(self.loss_cls).__name__ == 'VarifocalLoss'
        p = ops.make_rpn_loss_params(
            B, sizes, self.anchor_generator.strides, self.num_anchors, Gmax, a.pos_iou_thr,
            a.neg_iou_thr, a.get('min_pos_iou', 0.0), self.gamma,
            self.loss_cls.cfg.get('gamma', 2.0),
            self.loss_cls.cfg.get('alpha', 0.75 if varifocal else 0.25),
            self.loss_cls.loss_weight, self.loss_bbox.loss_weight, self.loss_centerness.loss_weight,
            self.aug_loss.loss_weight if self.with_aug_loss else 0.0,
            cls_loss='varifocal' if varifocal else 'focal')
        l_cls, l_bbox, l_iou, _ = ops.rpn_loss(p, cls_scores, bbox_preds, iou_preds, base, gtb,
                                               num_gt, pad_hw)
        if not self.with_aug_loss:
            # without aug_reg_loss the reference does not halve loss_bbox (:346-348)
            l_bbox = [v * 2 for v in l_bbox]
        return dict(loss_rpn_cls=l_cls, loss_rpn_bbox=l_bbox, loss_rpn_iou=l_iou)

    def forward_train(self, x, img_metas, gt_bboxes, gt_labels=None, gt_bboxes_ignore=None,
                      proposal_cfg=None, **kwargs):
        """atss_rpn_head.py:270-294: losses, and proposals when ``proposal_cfg`` is given.
        With ``self.padded_proposals = True`` (opt-in, default off) the proposals are returned
        as device-resident ``PaddedProposals`` — what ``ProbRoIHead.forward_train`` of this
        package consumes directly — instead of the reference's per-image list, whose lengths
        cost a host read per step."""
        outs = self(x)
        losses = self.loss(*outs, gt_bboxes, img_metas, gt_bboxes_ignore=gt_bboxes_ignore)
        if proposal_cfg is None:
            return losses
        if getattr(self, 'padded_proposals', False):
            return losses, self.get_bboxes_padded(*outs, img_metas, cfg=proposal_cfg)
        return losses, self.get_bboxes(*outs, img_metas, cfg=proposal_cfg)
