"""Base-anchor table of mmdet's AnchorGenerator.

Follows mmdet/core/anchor/anchor_generator.py:61-113 (constructor: scales from
``octave_base_scale * 2**(i/scales_per_octave)`` in float64 numpy, cast by
``torch.Tensor``) and :151-194 (``gen_single_level_base_anchors``), computed on
the CPU in fp32 torch exactly like the reference does at construction time.
The per-location grid (:338-381, ``base_anchors + shifts``) is NOT
materialised: the RPN kernel adds ``(x*stride_w, y*stride_h)`` on the fly.
"""
import numpy as np
import torch
from torch.nn.modules.utils import _pair


class AnchorGenerator:

    def __init__(self, strides, ratios, scales=None, base_sizes=None,
                 scale_major=True, octave_base_scale=None, scales_per_octave=None,
                 centers=None, center_offset=0.):
        if center_offset != 0:
            assert centers is None
        if not (0 <= center_offset <= 1):
            raise ValueError('center_offset should be in range [0, 1], '
                             f'{center_offset} is given.')
        if centers is not None:
            assert len(centers) == len(strides)
        self.strides = [_pair(stride) for stride in strides]
        self.base_sizes = [min(stride) for stride in self.strides
                           ] if base_sizes is None else base_sizes
        assert len(self.base_sizes) == len(self.strides)
        assert ((octave_base_scale is not None and scales_per_octave is not None)
                ^ (scales is not None))
        if scales is not None:
            self.scales = torch.Tensor(scales)
        else:
            octave_scales = np.array(
                [2**(i / scales_per_octave) for i in range(scales_per_octave)])
            self.scales = torch.Tensor(octave_scales * octave_base_scale)
        self.octave_base_scale = octave_base_scale
        self.scales_per_octave = scales_per_octave
        self.ratios = torch.Tensor(ratios)
        self.scale_major = scale_major
        self.centers = centers
        self.center_offset = center_offset
        self.base_anchors = self.gen_base_anchors()

    @property
    def num_base_anchors(self):
        return [base_anchors.size(0) for base_anchors in self.base_anchors]

    @property
    def num_levels(self):
        return len(self.strides)

    def gen_base_anchors(self):
        multi_level_base_anchors = []
        for i, base_size in enumerate(self.base_sizes):
            center = self.centers[i] if self.centers is not None else None
            multi_level_base_anchors.append(
                self.gen_single_level_base_anchors(
                    base_size, scales=self.scales, ratios=self.ratios, center=center))
        return multi_level_base_anchors

    def gen_single_level_base_anchors(self, base_size, scales, ratios, center=None):
        w = h = base_size
        if center is None:
            x_center = self.center_offset * w
            y_center = self.center_offset * h
        else:
            x_center, y_center = center
        h_ratios = torch.sqrt(ratios)
        w_ratios = 1 / h_ratios
        if self.scale_major:
            ws = (w * w_ratios[:, None] * scales[None, :]).view(-1)
            hs = (h * h_ratios[:, None] * scales[None, :]).view(-1)
        else:
            ws = (w * scales[:, None] * w_ratios[None, :]).view(-1)
            hs = (h * scales[:, None] * h_ratios[None, :]).view(-1)
        base_anchors = [
            x_center - 0.5 * ws, y_center - 0.5 * hs, x_center + 0.5 * ws,
            y_center + 0.5 * hs
        ]
        return torch.stack(base_anchors, dim=-1)

    def base_anchor_table(self):
        """(L, A, 4) fp32 CPU tensor handed to brcnn_rpn_get_bboxes."""
        a = self.num_base_anchors
        assert all(x == a[0] for x in a), 'levels must share the anchor count'
        return torch.stack(self.base_anchors, dim=0).float().contiguous()

    def grid_anchors_cpu(self, featmap_sizes):
        """Materialised anchors (tests / host-side loss code only): same values
        as anchor_generator.py:338-381."""
        out = []
        for (h, w), base, (sw, sh) in zip(featmap_sizes, self.base_anchors, self.strides):
            sx = torch.arange(0, w) * sw
            sy = torch.arange(0, h) * sh
            xx = sx.repeat(len(sy))
            yy = sy.view(-1, 1).repeat(1, len(sx)).view(-1)
            shifts = torch.stack([xx, yy, xx, yy], dim=-1).type_as(base)
            out.append((base[None, :, :] + shifts[:, None, :]).view(-1, 4))
        return out
