"""SingleRoIExtractor — drop-in for
mmdet/models/roi_heads/roi_extractors/single_level_roi_extractor.py:9-115 and
base_roi_extractor.py:11-88.

The reference maps RoIs to levels with six torch ops, then loops over levels
doing nonzero -> gather -> mmcv RoIAlign -> scatter.  Here ``forward`` is ONE
kernel launch for all levels (``brcnn_roi_extract_forward``: level map fused,
output written once, no zero-init + scatter), autograd-enabled through the
deterministic backward; every level receives a gradient (:105-114).
"""
import torch
import torch.nn as nn

from . import ops
from .registry import ROI_EXTRACTORS


@ROI_EXTRACTORS.register_module()
class SingleRoIExtractor(nn.Module):

    def __init__(self, roi_layer, out_channels, featmap_strides, finest_scale=56,
                 init_cfg=None, channels_last_out=True):
        super().__init__()
        # (R,C,oh,ow) result stored (R,oh,ow,C) (torch.channels_last): the RoI-feature
        # hand-off ProbConvFCBBoxHead reads as a free view; any other consumer sees the
        # reference's logical tensor (``flatten(1)`` then costs one copy)
        self.channels_last_out = channels_last_out
        self.roi_layers = self.build_roi_layers(roi_layer, featmap_strides)
        self.out_channels = out_channels
        self.featmap_strides = featmap_strides
        self.finest_scale = finest_scale
        self.fp16_enabled = False

    @property
    def num_inputs(self):
        return len(self.featmap_strides)

    def build_roi_layers(self, layer_cfg, featmap_strides):
        cfg = dict(layer_cfg)
        layer_type = cfg.pop('type')
        if layer_type != 'RoIAlign':
            raise NotImplementedError(f'roi_layer {layer_type} is outside the hot path')
        return nn.ModuleList([ops.RoIAlign(spatial_scale=1 / s, **cfg) for s in featmap_strides])

    def map_roi_levels(self, rois, num_levels):
        return ops.map_roi_levels(rois, num_levels, self.finest_scale)

    def roi_rescale(self, rois, scale_factor):
        cx = (rois[:, 1] + rois[:, 3]) * 0.5
        cy = (rois[:, 2] + rois[:, 4]) * 0.5
        new_w = (rois[:, 3] - rois[:, 1]) * scale_factor
        new_h = (rois[:, 4] - rois[:, 2]) * scale_factor
        return torch.stack((rois[:, 0], cx - new_w * 0.5, cy - new_h * 0.5,
                            cx + new_w * 0.5, cy + new_h * 0.5), dim=-1)

    def forward(self, feats, rois, roi_scale_factor=None):
        layer = self.roi_layers[0]
        if len(rois) == 0:
            return feats[0].new_zeros(0, self.out_channels, *layer.output_size)
        if roi_scale_factor is not None:
            # the reference maps levels on the ORIGINAL rois, then pools the
            # rescaled ones (:81-84); no named config uses this branch
            raise NotImplementedError('roi_scale_factor is outside the hot path')
        feats = feats[:self.num_inputs]
        return ops.roi_extract(feats, rois, [l.spatial_scale for l in self.roi_layers][:len(feats)],
                               layer.output_size, layer.sampling_ratio, layer.aligned,
                               self.finest_scale, channels_last_out=self.channels_last_out)
