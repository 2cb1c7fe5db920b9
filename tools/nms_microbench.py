"""Wall-clock per call of the mmcv-style batched_nms operator (ops.py) at detection sizes.
  gpurun -- 'python tools/nms_microbench.py'"""
import os, sys, time, numpy as np, torch
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [_ROOT, os.path.join(_ROOT, 'tests')]
import synth
from boosting_rcnn_b200 import ops
dev = torch.device('cuda')
def run(K, nid, clustered):
    b = torch.from_numpy(synth.random_boxes(K, 800, 1333, seed=K, clustered=clustered)).to(dev)
    s = torch.rand(K, device=dev)
    ids = torch.randint(0, nid, (K,), device=dev)
    cfg = dict(type='nms', iou_threshold=0.7)
    for _ in range(3): d, k = ops.batched_nms(b, s, ids, cfg)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10): d, k = ops.batched_nms(b, s, ids, cfg)
    torch.cuda.synchronize()
    print(f'batched_nms K={K} ids={nid} clustered={clustered}: {(time.perf_counter()-t0)*100:.3f} ms/call, kept {k.numel()}')
for K, nid in ((1000, 4), (4693, 5), (15150, 5), (20000, 80), (5000, 80)):
    run(K, nid, False); run(K, nid, True)
