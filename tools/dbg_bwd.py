"""developer probe: run the RoIAlign backward of tools/roi_microbench.py's train case against
a -DBRCNN_DEBUG_TIMING build (build/libbrcnn_dbg.so) and print its per-phase cycle sums."""
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
geom = configs.IMAGE_GEOMETRY['coco']
rpn_head, roi_head, model = configs.build_hot_path('coco', train=True)
rpn_head = rpn_head.to(dev).eval()
B, C = 2, 256
sizes, h_feats, h_cls, h_box, h_iou = bench.make_inputs(B, geom['pad_shape'][:2], 9, C, seed=1234, pin=False)
metas = bench.img_metas_for(B, geom)
cls, box, iou = ([t.to(dev) for t in ts] for ts in (h_cls, h_box, h_iou))
with torch.no_grad():
    props = rpn_head.get_bboxes_padded(cls, box, iou, metas, cfg=model['train_cfg']['rpn_proposal'])
rois, _ = ops.bbox2roi_padded(props.boxes[:, :512].contiguous(), props.num.clamp(max=512))
scales = [1.0 / s for s in bench.STRIDES]
p = ops.make_roi_params(B, C, sizes, scales, 7)
go = torch.randn((rois.size(0), C, 7, 7), device=dev).contiguous(memory_format=torch.channels_last)
for _ in range(4):
    ops.roi_extract_backward(p, go, rois)
    torch.cuda.synchronize()
