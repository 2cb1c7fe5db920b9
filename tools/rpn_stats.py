"""Diagnostic: per-level kept counts / suppression stats of the bench workload."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import bench
from boosting_rcnn_b200 import configs, ops

dev = torch.device('cuda:0')
rpn_head, roi_head, model = configs.build_hot_path('utdac')
B = 16
sizes, f, c, bx, u = bench.make_inputs(B, (800, 1344), 9, 256, 1234, False)
metas = bench.img_metas_for(B, configs.IMAGE_GEOMETRY['utdac'])
t = lambda ts: [x.to(dev) for x in ts]
cfg = model['test_cfg']['rpn']
p = ops.make_rpn_params(B, sizes, bench.STRIDES, 9, cfg['nms_pre'], cfg['max_per_img'], 0.7, 0.0)
lay = ops.rpn_workspace_layout(p)
base = rpn_head.anchor_generator.base_anchor_table().to(dev)
hw = torch.tensor([[800, 1333]] * B, dtype=torch.float32, device=dev)
props, num, ws = ops.rpn_get_bboxes(p, t(c), t(bx), t(u), base, hw, return_workspace=True)
torch.cuda.synchronize()
ws = ws.cpu().numpy()
L = 5
kc = ws[lay.kept_count:lay.kept_count + B * L * 4].view(np.int32).reshape(B, L)
kp = ws[lay.kept_pos:lay.kept_pos + B * L * lay.keep_cap * 4].view(np.int32).reshape(B, L, lay.keep_cap)
cv = ws[lay.cand_valid:lay.cand_valid + B * L * lay.cand_cap].reshape(B, L, lay.cand_cap)
print('kept_count per level (img 0..3):\n', kc[:4])
last = np.array([[kp[b, l, kc[b, l] - 1] if kc[b, l] else -1 for l in range(L)] for b in range(B)])
print('candidate rank of last kept (img 0..3):\n', last[:4])
print('valid frac per level', cv.reshape(B, L, -1).mean(axis=(0, 2)))
print('num proposals', num.cpu().numpy())
pr = props.cpu().numpy()
w = pr[0, :, 2] - pr[0, :, 0]; h = pr[0, :, 3] - pr[0, :, 1]
print('proposal size percentiles (w):', np.percentile(w, [5, 25, 50, 75, 95]), ' (h):', np.percentile(h, [5, 25, 50, 75, 95]))
print('score percentiles:', np.percentile(pr[0, :, 4], [0, 50, 100]))
