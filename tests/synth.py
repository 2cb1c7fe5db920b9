"""Seeded synthetic inputs shared by the parity tests, smoke() and bench.py
(SURVEY.md §8d).  numpy on the host; callers move them to the GPU."""
import math

import numpy as np

STRIDES = (8, 16, 32, 64, 128)


def featmap_sizes(pad_h, pad_w, strides=STRIDES):
    return [(math.ceil(pad_h / s), math.ceil(pad_w / s)) for s in strides]


def rpn_outputs(batch, sizes, num_anchors, seed=0, cls_std=1.5, box_std=0.3,
                duplicate_frac=0.0):
    """cls/iou ~ N(0, cls_std), bbox ~ N(0, box_std); optional exact duplicates
    (quantised logits) to exercise the tie-break."""
    rng = np.random.RandomState(seed)
    cls, box, iou = [], [], []
    for (h, w) in sizes:
        c = rng.normal(0, cls_std, (batch, num_anchors, h, w)).astype(np.float32)
        u = rng.normal(0, cls_std, (batch, num_anchors, h, w)).astype(np.float32)
        if duplicate_frac > 0:
            m = rng.rand(*c.shape) < duplicate_frac
            c[m] = np.round(c[m] * 4) / 4
            u[m] = np.round(u[m] * 4) / 4
        cls.append(c)
        iou.append(u)
        box.append(rng.normal(0, box_std, (batch, 4 * num_anchors, h, w)).astype(np.float32))
    return cls, box, iou


def fpn_feats(batch, channels, sizes, seed=0):
    rng = np.random.RandomState(seed)
    return [rng.normal(0, 1, (batch, channels, h, w)).astype(np.float32) for (h, w) in sizes]


def random_rois(batch, n_per_img, img_h, img_w, seed=0, min_size=2.0, max_size=None,
                clustered=False):
    """(batch*n, 5) [b, x1, y1, x2, y2]; log-uniform sizes spanning all levels."""
    rng = np.random.RandomState(seed)
    max_size = max_size or max(img_h, img_w)
    rois = []
    for b in range(batch):
        if clustered:
            seeds = rng.rand(max(n_per_img // 10, 1), 2) * [img_w, img_h]
            ctr = seeds[rng.randint(0, len(seeds), n_per_img)] + rng.normal(0, 4, (n_per_img, 2))
        else:
            ctr = rng.rand(n_per_img, 2) * [img_w, img_h]
        wh = np.exp(rng.uniform(np.log(min_size), np.log(max_size), (n_per_img, 2)))
        x1 = np.clip(ctr[:, 0] - wh[:, 0] / 2, 0, img_w)
        x2 = np.clip(ctr[:, 0] + wh[:, 0] / 2, 0, img_w)
        y1 = np.clip(ctr[:, 1] - wh[:, 1] / 2, 0, img_h)
        y2 = np.clip(ctr[:, 1] + wh[:, 1] / 2, 0, img_h)
        rois.append(np.stack([np.full(n_per_img, b), x1, y1, x2, y2], 1))
    return np.concatenate(rois, 0).astype(np.float32)


def random_boxes(n, img_h, img_w, seed=0, clustered=False):
    r = random_rois(1, n, img_h, img_w, seed=seed, min_size=8, max_size=400,
                    clustered=clustered)
    return r[:, 1:].copy()


def multiclass_inputs(R, C, seed=0, img=(800, 1333)):
    """(R, 4C) per-class boxes clustered around R seeds and (R, C+1) scores with
    distinct values (no ties): input of multiclass_nms at COCO scale."""
    rng = np.random.RandomState(seed)
    ctr = rng.uniform(40, 760, (R, 1, 2)) * [img[1] / 800.0, 1.0] + rng.normal(0, 8, (R, C, 2))
    wh = rng.uniform(20, 200, (R, C, 2))
    b = np.concatenate([ctr - wh / 2, ctr + wh / 2], -1)
    b[..., 0::2] = b[..., 0::2].clip(0, img[1])
    b[..., 1::2] = b[..., 1::2].clip(0, img[0])
    s = rng.permutation(R * (C + 1)).reshape(R, C + 1).astype(np.float32) / (R * (C + 1))
    return b.reshape(R, C * 4).astype(np.float32), s.astype(np.float32)


RCNN_TRAIN_CASES = ('plenty', 'few_negatives', 'no_gt_image', 'many_positives')


def rcnn_train_case(case, batch=3, num_classes=80):
    """Inputs of the R-CNN training front-end (assign + sample + targets + prior) for one
    named case: per image GT boxes (G,4), GT labels (G,) int64 and proposals (n,5)
    [x1,y1,x2,y2,score] with scores descending (the RPN's output order).
      plenty         : many negatives, a few positives (both lists go through randperm or not)
      few_negatives  : fewer negatives than the quota (no permutation of the negatives)
      no_gt_image    : image 1 has no GT (everything background)
      many_positives : > 128 positive candidates -> randperm on the positives; some GTs are
                       dropped, which misaligns the reference's `pos_inds[num_gts:]` prior
                       slice (prob_roi_head.py:52) — reproduced, not fixed."""
    assert case in RCNN_TRAIN_CASES
    rng = np.random.RandomState(11)
    gts, labels, plist = [], [], []
    for b in range(batch):
        G = 6 if case != 'no_gt_image' or b != 1 else 0
        g = random_boxes(max(G, 1), 250, 317, seed=40 + b)[:G].reshape(-1, 4)
        gts.append(g.astype(np.float32))
        labels.append(rng.randint(0, num_classes, G).astype(np.int64))
        n_jit = {'plenty': 10, 'few_negatives': 60, 'no_gt_image': 10, 'many_positives': 60}[case]
        n_rnd = {'plenty': 900, 'few_negatives': 40, 'no_gt_image': 500, 'many_positives': 700}[case]
        parts = [g + rng.normal(0, 1.5, g.shape) for _ in range(n_jit)] if G else []
        parts.append(random_boxes(n_rnd, 250, 317, seed=60 + b))
        bx = np.concatenate(parts).astype(np.float32)
        bx = bx[rng.permutation(len(bx))]
        sc = np.sort(rng.rand(len(bx)).astype(np.float32))[::-1].copy()
        plist.append(np.concatenate([bx, sc[:, None]], 1).astype(np.float32))
    return gts, labels, plist


RPN_LOSS_CASES = ('basic', 'isolated_gt', 'no_gt_image', 'partial_valid')


def rpn_loss_case(case, batch=2, pad_hw=(128, 160), num_anchors=9):
    """Inputs of the RPN loss path (ATSSRPNHead.loss, atss_rpn_head.py:405-464) for one named
    case: RPN head outputs per level (B,A,H,W)/(B,4A,H,W)/(B,A,H,W), GT boxes per image and
    img_metas (img_shape, pad_shape).
      basic         : 3-5 GTs per image of mixed sizes
      isolated_gt   : one extra GT outside every anchor's reach in image 0 — with
                      min_pos_iou=0 and match_low_quality=True it claims every anchor whose
                      overlap with it equals its (zero) maximum (max_iou_assigner.py:187-202)
      no_gt_image   : image 1 has no GT (all anchors negative)
      partial_valid : image 1's own pad_shape is smaller than the batch pad, so part of each
                      level is flagged invalid (anchor_generator.py:383-434)"""
    assert case in RPN_LOSS_CASES
    seed = 500 + RPN_LOSS_CASES.index(case)
    rng = np.random.RandomState(seed)
    sizes = featmap_sizes(*pad_hw)
    cls, box, iou = rpn_outputs(batch, sizes, num_anchors, seed=seed, cls_std=1.5, box_std=0.3)
    img_hw = (pad_hw[0] - 6, pad_hw[1] - 3)
    gts, metas = [], []
    for b in range(batch):
        n = int(rng.randint(3, 6))
        wh = np.exp(rng.uniform(np.log(12), np.log(100), (n, 2)))
        ctr = rng.uniform(0.15, 0.85, (n, 2)) * [img_hw[1], img_hw[0]]
        g = np.concatenate([ctr - wh / 2, ctr + wh / 2], 1)
        g[:, 0::2] = g[:, 0::2].clip(0, img_hw[1])
        g[:, 1::2] = g[:, 1::2].clip(0, img_hw[0])
        if case == 'isolated_gt' and b == 0:
            g = np.concatenate([g[:2], [[900., 900., 910., 910.]], g[2:]], 0)
        if case == 'no_gt_image' and b == 1:
            g = np.zeros((0, 4))
        gts.append(g.astype(np.float32))
        pad = pad_hw
        if case == 'partial_valid' and b == 1:
            pad = (pad_hw[0] - 32, pad_hw[1] - 48)
        metas.append(dict(img_shape=(min(img_hw[0], pad[0]), min(img_hw[1], pad[1]), 3),
                          pad_shape=(pad[0], pad[1], 3)))
    return dict(cls=cls, box=box, iou=iou, gt_bboxes=gts, img_metas=metas, sizes=sizes,
                num_anchors=num_anchors)
