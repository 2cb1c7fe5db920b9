#!/usr/bin/env python
"""Device time of the RoIAlign forward / backward kernels on bench-shaped inputs, each
captured alone in a CUDA graph, L2 flushed before every repetition.

  python tools/roi_microbench.py --cfg utdac --batch 16            # configs[1] RoIs (256 / img)
  python tools/roi_microbench.py --cfg coco --batch 2 --train      # configs[2] RoIs (512 / img)

Prints one JSON line; `union_bytes` is the number of distinct feature bytes the RoIs touch
(bitmap union of the footprints per (image, level)), `algo_*` = output bytes + union bytes."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402


def graph_time(fn, reps, dev):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    g.replay()
    evs = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in evs)
    return float(np.mean(t)), float(t[len(t) // 2])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--cfg', default='utdac')
    ap.add_argument('--batch', type=int, default=16)
    ap.add_argument('--train', action='store_true')
    ap.add_argument('--rois-per-img', type=int, default=0)
    ap.add_argument('--reps', type=int, default=20)
    ap.add_argument('--dump-rois', default='', help='write the (R,5) RoIs to this .npy file')
    args = ap.parse_args()
    from boosting_rcnn_b200 import configs, ops
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    geom = configs.IMAGE_GEOMETRY[args.cfg]
    rpn_head, roi_head, model = configs.build_hot_path(args.cfg, train=args.train)
    rpn_head = rpn_head.to(dev).eval()
    A, C = rpn_head.num_anchors, 256
    B = args.batch
    sizes, h_feats, h_cls, h_box, h_iou = bench.make_inputs(B, geom['pad_shape'][:2], A, C,
                                                            seed=1234, pin=False)
    metas = bench.img_metas_for(B, geom)
    feats = [t.to(dev).contiguous(memory_format=torch.channels_last) for t in h_feats]
    cls, box, iou = ([t.to(dev) for t in ts] for ts in (h_cls, h_box, h_iou))
    cfg = model['train_cfg']['rpn_proposal'] if args.train else None
    with torch.no_grad():
        props = rpn_head.get_bboxes_padded(cls, box, iou, metas, cfg=cfg)
    n = args.rois_per_img or (512 if args.train else props.boxes.size(1))
    boxes = props.boxes[:, :n].contiguous()
    num = props.num.clamp(max=n)
    rois, _ = ops.bbox2roi_padded(boxes, num)
    scales = [1.0 / s for s in bench.STRIDES]
    R = rois.size(0)
    if args.dump_rois:
        np.save(args.dump_rois, rois.cpu().numpy())
    out = {'cfg': args.cfg, 'batch': B, 'rois': int(R), 'live': int((rois[:, 0] >= 0).sum()),
           }
    with torch.no_grad():
        for name, cl in (('fwd_nchw', False), ('fwd_hwc', True)):
            out[name + '_ms'] = graph_time(
                lambda: ops.roi_extract(feats, rois, scales, 7, channels_last_out=cl),
                args.reps, dev)
    if args.train or os.environ.get('ROI_MB_BWD'):
        sizes_hw = [tuple(f.shape[-2:]) for f in feats]
        for name, cl in (('bwd_nchw', False), ('bwd_hwc', True)):
            p = ops.make_roi_params(B, C, sizes_hw, scales, 7)
            go = torch.randn((R, C, 7, 7), device=dev)
            if cl:
                go = go.contiguous(memory_format=torch.channels_last)
            with torch.no_grad():
                out[name + '_ms'] = graph_time(
                    lambda: ops.roi_extract_backward(p, go, rois), args.reps, dev)
    rois_h = rois.cpu().numpy()
    out['union_bytes'] = bench.roi_union_bytes(rois_h, sizes, C) if hasattr(bench, 'roi_union_bytes') else None
    out['out_bytes'] = int(R) * C * 49 * 4
    out['feat_bytes'] = int(sum(f.numel() * 4 for f in feats))
    peak, _ = bench._peak()
    for k in ('fwd_nchw', 'fwd_hwc'):
        if out['union_bytes'] is not None:
            ms = out[k + '_ms'][0]
            out[k + '_frac'] = (out['out_bytes'] + out['union_bytes']) / (ms * 1e-3) / 1e9 / peak
    for k in ('bwd_nchw', 'bwd_hwc'):
        if k + '_ms' in out:
            ms = out[k + '_ms'][0]
            out[k + '_frac'] = (out['out_bytes'] + out['feat_bytes']) / (ms * 1e-3) / 1e9 / peak
    print(json.dumps(out))


if __name__ == '__main__':
    main()
