"""developer probe: RoIAlign forward of the configs[1] bench RoIs against a
-DBRCNN_DEBUG_TIMING build (build/libbrcnn_dbg.so): per-role phase cycle sums."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import torch
import bench
from boosting_rcnn_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, 'build', 'libbrcnn_dbg.so')
from boosting_rcnn_b200 import configs, ops

dev = torch.device('cuda', 0)
geom = configs.IMAGE_GEOMETRY['utdac']
rpn_head, roi_head, model = configs.build_hot_path('utdac')
rpn_head = rpn_head.to(dev).eval()
B, C = 16, 256
sizes, h_feats, h_cls, h_box, h_iou = bench.make_inputs(B, geom['pad_shape'][:2], 9, C, seed=1234, pin=False)
metas = bench.img_metas_for(B, geom)
feats = [t.to(dev).contiguous(memory_format=torch.channels_last) for t in h_feats]
cls, box, iou = ([t.to(dev) for t in ts] for ts in (h_cls, h_box, h_iou))
with torch.no_grad():
    props = rpn_head.get_bboxes_padded(cls, box, iou, metas)
    rois, _ = ops.bbox2roi_padded(props.boxes, props.num)
    scales = [1.0 / s for s in bench.STRIDES]
    for _ in range(3):
        ops.roi_extract(feats, rois, scales, 7, channels_last_out=True)
        torch.cuda.synchronize()
