"""developer probe: per-phase cycle sums of the RPN NMS cluster kernel (CTA 0) for the test and
the train proposal settings, against a -DBRCNN_DEBUG_TIMING build (build/libbrcnn_dbg.so)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import torch
import bench
from boosting_rcnn_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, 'build', 'libbrcnn_dbg.so')
from boosting_rcnn_b200 import configs

dev = torch.device('cuda', 0)
for cfg_name, B, train in (('utdac', 16, False), ('coco', 2, True)):
    geom = configs.IMAGE_GEOMETRY[cfg_name]
    rpn_head, roi_head, model = configs.build_hot_path(cfg_name, train=train)
    rpn_head = rpn_head.to(dev).eval()
    sizes, h_feats, h_cls, h_box, h_iou = bench.make_inputs(B, geom['pad_shape'][:2], rpn_head.num_anchors, 8,
                                                            seed=1234, pin=False)
    metas = bench.img_metas_for(B, geom)
    cls, box, iou = ([t.to(dev) for t in ts] for ts in (h_cls, h_box, h_iou))
    cfg = model['train_cfg']['rpn_proposal'] if train else None
    print(cfg_name, 'train' if train else 'test', file=sys.stderr, flush=True)
    with torch.no_grad():
        for _ in range(2):
            props = rpn_head.get_bboxes_padded(cls, box, iou, metas, cfg=cfg)
            torch.cuda.synchronize()
    print('  proposals per image', props.num.tolist()[:4], file=sys.stderr, flush=True)
