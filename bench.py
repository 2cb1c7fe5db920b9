#!/usr/bin/env python
"""bench.py — RoI-path images/s of the Boosting R-CNN proposal-to-RoI hot path.

Workload (BASELINE.json configs[1]): boosting_rcnn_r50_pafpn_1x_utdac inference,
16 synthetic 1333x800 images per GPU, random-init head weights.  One "step" =
RPN conv outputs + FPN maps  ->  proposals (score, top-k, decode, NMS)
-> RoI features (level map + RoIAlign) -> 2-fc head (torch/cuBLAS, fp32)
-> score fusion + per-class decode + class NMS -> detections, for the batch.

  python bench.py [--gpus N --steps K --warmup W]      # this repo (sm_100a kernels)
  python bench.py --impl reference [...]               # CPU reference arm (oracle)

Prints ONE JSON line (see README/DESIGN.md for the keys).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, 'tests')):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

STRIDES = (8, 16, 32, 64, 128)
WORKLOAD = 'boosting_rcnn_r50_pafpn_1x_utdac inference, 16 synthetic 1333x800 images per GPU'
SCALE_FACTOR = (1.6662, 1.6667, 1.6662, 1.6667)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=16, help='images per GPU per step')
    ap.add_argument('--cfg', default='utdac', choices=['utdac', 'coco', 'voc'])
    ap.add_argument('--mode', default='infer', choices=['infer', 'train', 'stress'],
                    help='infer: configs[1] (the headline); train: configs[2] training step '
                         '(use --cfg coco --batch 2); stress: configs[4] proposal stress test')
    ap.add_argument('--cpu-images', type=int, default=0,
                    help='images in the CPU-baseline sample (0: one per host thread, <= 16)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--stress-input', default='boxes', choices=['boxes', 'clustered', 'anchors'],
                    help='--mode stress: random boxes (SURVEY §8d cfg 5), the clustered variant, '
                         'or RPN outputs on the COCO pyramid')
    ap.add_argument('--lean', action='store_true',
                    help='secondary workloads with large heads (VOC at 1000 proposals): one graph, '
                         'no channels_last / TF32 / two-stream variants')
    ap.add_argument('--no-train-record', action='store_true',
                    help='skip the configs[2] training-step sub-record of the default line')
    ap.add_argument('--strong-total', type=int, default=16,
                    help='images of the strong-scaling sub-record (sharded over the GPUs)')
    ap.add_argument('--no-graph', action='store_true', help='launch kernels one by one')
    ap.add_argument('--train-eager', action='store_true',
                    help='training record without the captured R-CNN half (RcnnTrainGraph)')
    ap.add_argument('--rpn-max-per-img', type=int, default=0,
                    help='override test_cfg.rpn.max_per_img (BASELINE configs[3]: VOC at 1000)')
    return ap.parse_args()


# ----------------------------------------------------------------------------
# synthetic inputs, generated on the host, seeded per rank
# ----------------------------------------------------------------------------
def make_inputs(batch, pad_hw, num_anchors, channels, seed, pin):
    g = torch.Generator().manual_seed(seed)
    sizes = [(-(-pad_hw[0] // s), -(-pad_hw[1] // s)) for s in STRIDES]

    def rnd(shape, std):
        t = torch.empty(shape, dtype=torch.float32)
        t.normal_(0, std, generator=g)
        return t.pin_memory() if pin else t

    feats = [rnd((batch, channels, h, w), 1.0) for (h, w) in sizes]
    cls = [rnd((batch, num_anchors, h, w), 1.5) for (h, w) in sizes]
    box = [rnd((batch, 4 * num_anchors, h, w), 0.3) for (h, w) in sizes]
    iou = [rnd((batch, num_anchors, h, w), 1.5) for (h, w) in sizes]
    return sizes, feats, cls, box, iou


def img_metas_for(batch, geom):
    return [dict(img_shape=geom['img_shape'], pad_shape=geom['pad_shape'],
                 scale_factor=np.array(SCALE_FACTOR, dtype=np.float32)) for _ in range(batch)]


# ----------------------------------------------------------------------------
# CPU reference arm: the oracle (restated mmdet + mmcv CPU ops) on host cores
# ----------------------------------------------------------------------------
class CpuReference:
    """Same step on the host: oracle C restatement per image, one image per
    thread (ctypes releases the GIL), the 2-fc head on torch CPU."""

    def __init__(self, rpn_head, roi_head, model, geom, threads):
        from oracle import oracle
        self.o = oracle
        oracle.lib()
        self.base = rpn_head.anchor_generator.base_anchor_table().numpy()
        self.test_rpn = model['test_cfg']['rpn']
        self.test_rcnn = model['test_cfg']['rcnn']
        self.geom = geom
        self.nc = roi_head.bbox_head.num_classes
        self.head = roi_head.bbox_head
        self.cpu_head = None
        self.threads = threads

    def _head_cpu(self):
        if self.cpu_head is None:
            import copy
            self.cpu_head = copy.deepcopy(self.head).cpu().eval()
        return self.cpu_head

    def one_image(self, feats, cls, box, iou):
        o = self.o
        hw = self.geom['img_shape'][:2]
        props = o.rpn_get_bboxes_single(cls, box, iou, self.base, STRIDES, hw,
                                        self.test_rpn['nms_pre'], self.test_rpn['max_per_img'],
                                        self.test_rpn['nms']['iou_threshold'],
                                        self.test_rpn['min_bbox_size'])
        n = props.shape[0]
        rois = np.concatenate([np.zeros((n, 1), np.float32), props[:, :4]], 1)
        rf, _ = o.roi_extract_forward([f[None] for f in feats], rois, [1.0 / s for s in STRIDES])
        with torch.no_grad():
            torch.set_num_threads(1)
            cs, bp = self._head_cpu()(torch.from_numpy(rf))
        fused = o.fuse_scores(cs.numpy(), props[:, 4])
        return o.rcnn_get_bboxes_single(rois, fused, bp.numpy(), hw, SCALE_FACTOR, self.nc,
                                        self.test_rcnn['score_thr'],
                                        self.test_rcnn['nms']['iou_threshold'],
                                        self.test_rcnn['max_per_img'], rescale=True)

    def run(self, feats, cls, box, iou, n_images):
        """n_images may exceed the images held in the tensors (one image per host
        thread on boxes with more threads than images per step): indices wrap."""
        from concurrent.futures import ThreadPoolExecutor
        self._head_cpu()
        nb = feats[0].shape[0]
        fn = [[f[b % nb].numpy() for f in feats] for b in range(n_images)]
        cn = [[c[b % nb].numpy() for c in cls] for b in range(n_images)]
        bn = [[c[b % nb].numpy() for c in box] for b in range(n_images)]
        un = [[c[b % nb].numpy() for c in iou] for b in range(n_images)]
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=self.threads) as ex:
            res = list(ex.map(lambda i: self.one_image(fn[i], cn[i], bn[i], un[i]), range(n_images)))
        return time.perf_counter() - t0, res


# ----------------------------------------------------------------------------
# clocks sampling during the timed region (B200_PROFILING.md)
# ----------------------------------------------------------------------------
class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(['nvidia-smi', f'--id={gpu_index}', f'--query-gpu={self.Q}',
                                       '--format=csv,noheader,nounits', '-lms', '50'],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(',')]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        os.unlink(self.f.name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None,
                    sm_max_mhz=max(mx) if mx else None, samples=len(sm),
                    reasons=sorted(reasons))


def roi_footprint_bytes(rois, sizes, channels, finest=56.0):
    """Algorithmic feature bytes of RoIAlign: per RoI the
    (floor..ceil)+tap footprint on its level, x C x 4 B."""
    r = rois[rois[:, 0] >= 0]
    w, h = r[:, 3] - r[:, 1], r[:, 4] - r[:, 2]
    scale = np.sqrt(w * h)
    lvl = np.clip(np.floor(np.log2(scale / finest + 1e-6)), 0, len(sizes) - 1).astype(int)
    tot = 0
    for l, (H, W) in enumerate(sizes):
        m = lvl == l
        s = 1.0 / STRIDES[l]
        fw = np.minimum(np.ceil(r[m, 3] * s) - np.floor(r[m, 1] * s) + 2, W)
        fh = np.minimum(np.ceil(r[m, 4] * s) - np.floor(r[m, 2] * s) + 2, H)
        tot += float((fw * fh).sum()) * channels * 4
    return tot


def roi_union_bytes(rois, sizes, channels, finest=56.0):
    """Distinct feature bytes the RoIs read: per (image, level) the UNION of the footprint
    pixels (rows/columns any bilinear tap of the RoI can touch, the kernels' own
    conservative box) as a bitmap, x C x 4 B.  This is what has to cross HBM at least
    once; `roi_footprint_bytes` (SURVEY.md §8d's sum of footprints) counts a pixel once
    per RoI that covers it and is only an upper bound on it."""
    r = rois[rois[:, 0] >= 0]
    if len(r) == 0:
        return 0
    nb = int(r[:, 0].max()) + 1
    w, h = r[:, 3] - r[:, 1], r[:, 4] - r[:, 2]
    scale = np.sqrt(w * h)
    lvl = np.clip(np.floor(np.log2(scale / finest + 1e-6)), 0, len(sizes) - 1).astype(int)
    tot = 0
    for l, (H, W) in enumerate(sizes):
        s = np.float32(1.0 / STRIDES[l])
        bm = np.zeros((nb, H, W), dtype=bool)
        for q in r[lvl == l]:
            x0, y0 = q[1] * s - 0.5, q[2] * s - 0.5
            x1, y1 = q[3] * s - 0.5, q[4] * s - 0.5
            if x1 < -1 or y1 < -1 or x0 > W or y0 > H:
                continue
            xa, xb = max(0, int(np.floor(max(x0, 0)))), min(W - 1, int(np.floor(min(max(x1, 0), W))) + 1)
            ya, yb = max(0, int(np.floor(max(y0, 0)))), min(H - 1, int(np.floor(min(max(y1, 0), H))) + 1)
            bm[int(q[0]), ya:yb + 1, xa:xb + 1] = True
        tot += int(bm.sum()) * channels * 4
    return tot


# ----------------------------------------------------------------------------
# secondary workloads (not the headline line; same JSON contract)
# ----------------------------------------------------------------------------
def _dev_time(fn, reps, dev, flush_mb=256):
    """mean device ms of fn() over reps, L2 flushed before each rep"""
    flush = torch.empty(flush_mb << 20, dtype=torch.uint8, device=dev)
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    return float(np.mean([a.elapsed_time(b) for a, b in evs]))


def _graph_time(fn, reps, dev, flush_mb=256):
    """mean device ms of fn() captured alone in a CUDA graph (no host launch gaps in the
    number), L2 flushed before every repetition"""
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    flush = torch.empty(flush_mb << 20, dtype=torch.uint8, device=dev)
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
    return float(np.mean([a.elapsed_time(b) for a, b in evs]))


def _peak():
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_path):
        return json.load(open(peaks_path))['hbm_gbs'], 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def _dist_setup(world, dev):
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
        return dist
    return None


def _timed_steps(fn, K, W, world, dist, dev):
    for _ in range(W):
        fn()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        fn()
    e1.record()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def train_record(args, rank, world, local_rank, dist, cfg_name='coco', B=2, K=20, W=3,
                 with_stages=True):
    """configs[2]: R-CNN training step on B200-generated proposals: train-cfg proposal
    generation (nms_pre 4000 / 2000 per image), fused assign + sample + targets + prior
    (2 launches + the reference's CPU randperm), RoIAlign forward ((R,7,7,C) hand-off), 2-fc
    head, boost loss, backward (head GEMMs on cuBLAS, cp.async-staged RoIAlign backward gather),
    then the reference's collectives: NCCL all-reduce of the head gradients (one flat bucket,
    mmdet/apis/train.py:75-83) + ONE fused scalar all-reduce for the logged losses
    (detectors/base.py:201-207).  Returns the record (rank 0) or None."""
    from boosting_rcnn_b200 import _lib, configs, ops
    from boosting_rcnn_b200 import dist as bdist
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device('cuda', local_rank)
    lib = _lib.load()
    geom = configs.IMAGE_GEOMETRY[cfg_name]
    torch.manual_seed(0)
    rpn_head, roi_head, model = configs.build_hot_path(cfg_name, train=True)
    rpn_head, roi_head = rpn_head.to(dev).eval(), roi_head.to(dev).train()
    A, C = rpn_head.num_anchors, roi_head.bbox_roi_extractor.out_channels
    NC = roi_head.bbox_head.num_classes
    sizes, h_feats, h_cls, h_box, h_iou = make_inputs(B, geom['pad_shape'][:2], A, C,
                                                      seed=4321 + rank, pin=False)
    metas = img_metas_for(B, geom)
    feats = [t.to(dev).requires_grad_(True) for t in h_feats]
    # the RPN head outputs are leaves here (the conv tower is torch/cuDNN, outside the path):
    # the RPN loss sends its gradients to them
    cls, box, iou = ([t.to(dev).requires_grad_(True) for t in ts] for ts in (h_cls, h_box, h_iou))
    rng = np.random.RandomState(77 + rank)
    gts, labels = [], []
    for b in range(B):
        n = int(rng.randint(1, 21))
        wh = rng.uniform(16, 400, (n, 2))
        ctr = rng.uniform(0, 1, (n, 2)) * [geom['img_shape'][1], geom['img_shape'][0]]
        g = np.concatenate([ctr - wh / 2, ctr + wh / 2], 1)
        g[:, 0::2] = g[:, 0::2].clip(0, geom['img_shape'][1])
        g[:, 1::2] = g[:, 1::2].clip(0, geom['img_shape'][0])
        gts.append(torch.from_numpy(g.astype(np.float32)).to(dev))
        labels.append(torch.from_numpy(rng.randint(0, NC, n)).to(dev))
    prop_cfg = model['train_cfg']['rpn_proposal']
    params = [p for p in roi_head.parameters() if p.requires_grad]
    n_grad = sum(p.numel() for p in params)
    scal = {}
    comm_events = []

    def step(comm=True):
        for p in params:
            p.grad = None
        for f in feats + cls + box + iou:
            f.grad = None
        # RPN loss on the head outputs (anchor targets, focal / IoU / MSE / BCE, gradients): its
        # two reduce_mean normalisers are one fused device-side all-reduce inside ops.rpn_loss
        # Order: proposals and the R-CNN assignment are launched first, the RPN loss is queued
        # behind them, and only then does the host wait (an event) for the two candidate
        # counts the reference's CPU randperm needs: the GPU works on the RPN loss meanwhile.
        with torch.no_grad():   # padded proposals stay on the device (no sync)
            plist = rpn_head.get_bboxes_padded(cls, box, iou, metas, cfg=prop_cfg)
        pending = roi_head.assign_async(plist, gts, labels)
        rpn_losses = rpn_head.loss(cls, box, iou, gts, metas)
        losses = roi_head.forward_train(feats, metas, plist, gts, labels, assigned=pending)
        rpn_sums = {k: torch.stack(v).sum() for k, v in rpn_losses.items()}
        total = losses['loss_cls'] + losses['loss_bbox']
        for v in rpn_sums.values():
            total = total + v
        total.backward()
        losses = dict(losses, **rpn_sums)
        if world > 1 and comm:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            # one flat bucket (the head has ~14M parameters): all-reduce, average, scatter back
            flat = torch.cat([p.grad.reshape(-1) for p in params])
            dist.all_reduce(flat)
            flat /= world
            off = 0
            for p in params:
                n = p.numel()
                p.grad.copy_(flat[off:off + n].view_as(p.grad))
                off += n
            scal.update(bdist.fused_scalar_allreduce({k: v.detach() for k, v in losses.items()}))
            e1.record()
            comm_events.append((e0, e1))
        else:
            scal.update({k: v.detach() for k, v in losses.items()})

    # the captured R-CNN half must reproduce the eager one bit for bit (same kernels, same
    # CPU RNG stream): one step each way from the same seed, losses and every gradient compared
    graph_check = None
    roi_head.train_graph = False
    if not getattr(args, 'train_eager', False):
        def snapshot():
            torch.cuda.synchronize()
            return ([p.grad.clone() for p in params] + [f.grad.clone() for f in feats + cls + box + iou]
                    + [scal[k].clone() for k in sorted(scal)])
        torch.manual_seed(1234)
        step(False)
        ref = snapshot()
        roi_head.train_graph = True
        torch.manual_seed(1234)
        step(False)
        got = snapshot()
        captured = any(bool(g) for g in roi_head._train_graphs.values())
        graph_check = {'captured': captured,
                       'bit_identical_to_eager': all(torch.equal(a, b) for a, b in zip(ref, got))}
        if not captured:
            roi_head.train_graph = False
        del ref, got
    graphs = [g for g in roi_head._train_graphs.values() if g]
    l0 = lib.brcnn_launch_count()
    ms = _timed_steps(step, K, W, world, dist, dev)
    launches = (lib.brcnn_launch_count() - l0) * K // (K + W)
    if graphs:   # kernels inside the replayed graphs are not seen by the launch counter
        launches += graphs[0].launches_per_step * K
    ms_comm = 0.0
    if comm_events:
        torch.cuda.synchronize()
        ms_comm = float(np.mean([a.elapsed_time(b) for a, b in comm_events[-K:]]))
        if dist:
            t = torch.tensor([ms_comm], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_comm = float(t.item())
    ms_nocomm = _timed_steps(lambda: step(False), K, 1, world, dist, dev) if world > 1 else ms
    # the collective alone: the same flat bucket all-reduced back to back after a barrier (the
    # in-step figure above also contains the wait for the slowest rank of an eager step)
    ms_coll = 0.0
    if world > 1:
        flat = torch.zeros(n_grad, device=dev)
        for _ in range(3):
            dist.all_reduce(flat)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            dist.all_reduce(flat)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 10], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_coll = float(t.item())
        del flat
    if os.environ.get('BENCH_TRAIN_TRACE') and world == 1:
        # developer aid: chrome trace of three steps (kernels + host ranges) for gap analysis
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                step(False)
            torch.cuda.synchronize()
        prof.export_chrome_trace(os.environ['BENCH_TRAIN_TRACE'])
    n_graphs = len(graphs)
    launches_in_graphs = graphs[0].launches_per_step if graphs else 0
    del graphs
    roi_head._train_graphs.clear()     # graph <-> head reference cycle: release the pools now
    if rank != 0:
        return None
    rec = {
        'workload': f'boosting_rcnn {cfg_name} hot-path training step (RPN loss + proposals + '
                    f'R-CNN assign/sample + RoIAlign fwd/bwd + boost loss + 2-fc head), {B} '
                    f'synthetic 1333x800 images per GPU, 512 sampled RoIs/img (BASELINE configs[2])',
        'value': B * world * K / (ms * 1e-3), 'unit': 'images/s', 'n_gpus': world,
        'images_per_gpu': B, 'steps': K, 'warmup': W, 'ms_per_step': ms / K,
        'ms_allreduce': ms_comm, 'ms_per_step_no_comm': ms_nocomm / K,
        'ms_allreduce_collective_only': ms_coll,
        'allreduce_note': 'ms_allreduce = device time from the end of backward to the end of the '
                          'gradient exchange inside the step (bucket copies + NCCL all-reduce + the '
                          'wait for the slowest rank of an eager, launch-bound step); '
                          'ms_allreduce_collective_only = the same bucket all-reduced back to back',
        'allreduce_bytes': n_grad * 4,
        'collectives': ('none (1 GPU)' if world == 1 else
                        'NCCL all-reduce of the head gradients (one flat fp32 bucket) + one fused '
                        'scalar all-reduce of the logged losses'),
        'gpu_launches_per_step': int(launches // max(K, 1)),
        'launch_mode': ('RPN loss / proposals / R-CNN assignment eager, then one event wait for '
                        'the reference\'s CPU randperm, then the R-CNN half (sample + targets, '
                        'RoIAlign, 2-fc head, boost loss, backward) as two CUDA-graph replays'
                        if n_graphs else
                        'eager (one event wait per step: the reference\'s CPU randperm)'),
        'gpu_launches_per_step_inside_graphs': launches_in_graphs,
        'train_graph_check': graph_check,
        'losses': {k: float(v) for k, v in scal.items()},
    }
    if with_stages:
        peak, peak_src = _peak()
        with torch.no_grad():
            plist = rpn_head.get_bboxes(cls, box, iou, metas, cfg=prop_cfg)
        rois = torch.cat([torch.cat([p.new_full((min(512, p.size(0)), 1), float(b)),
                                     p[:512, :4]], 1) for b, p in enumerate(plist)])
        R = rois.size(0)
        scales = [1.0 / s_ for s_ in STRIDES]
        nhwc = [f.detach().contiguous(memory_format=torch.channels_last) for f in feats]
        sizes_hw = [tuple(f.shape[-2:]) for f in nhwc]
        rp = ops.make_roi_params(B, C, sizes_hw, scales, 7)
        go = torch.randn(R, C, 7, 7, device=dev).contiguous(memory_format=torch.channels_last)
        cs = torch.randn(R, NC + 1, device=dev)
        bp = torch.randn(R, 4 * NC, device=dev)
        lab = torch.randint(0, NC + 1, (R,), device=dev)
        lw = torch.ones(R, device=dev)
        pr = torch.rand(R, device=dev)
        bt, bw = torch.randn(R, 4, device=dev), (lab < NC).float()[:, None].expand(R, 4).contiguous()
        with torch.no_grad():
            st = {
                'rpn_get_bboxes_train_cfg': _graph_time(
                    lambda: rpn_head.get_bboxes_padded(cls, box, iou, metas, cfg=prop_cfg), 10, dev),
                'roi_align_fwd': _graph_time(
                    lambda: ops.roi_extract(nhwc, rois, scales, 7, channels_last_out=True), 10, dev),
                'roi_align_bwd': _graph_time(lambda: ops.roi_extract_backward(rp, go, rois), 10, dev),
                'boost_loss_fwd_bwd': _graph_time(
                    lambda: ops.boost_loss(cs, bp, lab, lw, pr, bt, bw, NC, False, 0.5, 0.0, 2.0,
                                           2.0, False), 10, dev),
            }
        dcls, dbox, diou = ([t.detach() for t in ts] for ts in (cls, box, iou))
        with torch.no_grad():
            # eager (the GT / pad tensors are built from host lists inside loss()).  This stage
            # timer runs on rank 0 only: the loss must not issue its reduce_mean all-reduce here
            ops.RPN_LOSS_REDUCE_MEAN = False
            try:
                st['rpn_loss_fwd_eager (targets + losses + raw grads)'] = _dev_time(
                    lambda: rpn_head.loss(dcls, dbox, diou, gts, metas), 10, dev)
            finally:
                ops.RPN_LOSS_REDUCE_MEAN = True
        feat_bytes = sum(f.numel() * 4 for f in feats)
        bwd_bytes = R * C * 49 * 4 + feat_bytes
        ach = bwd_bytes / (st['roi_align_bwd'] * 1e-3) / 1e9
        rec['stages_ms'] = st
        rec['roofline'] = dict(kernel='roi_bwd_prep_kernel + roi_bwd_gather5_kernel', bound='hbm',
                               achieved=ach, peak=peak, unit='GB/s', frac=ach / peak, traffic=None,
                               peak_source=peak_src, algorithmic_bytes_per_launch=bwd_bytes,
                               ms_per_launch=st['roi_align_bwd'],
                               note='bytes = grad_out read once + every grad map written once')
    return rec


def bench_train(args, rank, world, local_rank):
    """`--mode train`: the configs[2] record as the top-level JSON line."""
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    dist = _dist_setup(world, dev)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    rec = train_record(args, rank, world, local_rank, dist, cfg_name=args.cfg, B=args.batch,
                       K=args.steps, W=max(args.warmup, 3))
    clocks = sampler.stop() if sampler else None
    if rank == 0:
        out = {
            'metric': 'RoI-path training images/s @1333x800', 'value': rec['value'],
            'unit': 'images/s', 'n_gpus': world, 'steps': rec['steps'], 'warmup': rec['warmup'],
            'ms_per_step': rec['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': dict(workload=rec['workload'], images_per_gpu=rec['images_per_gpu'],
                           parallelism=f'image-sharded x{world}; ' + rec['collectives']),
            'gpu_launches': rec['gpu_launches_per_step'] * rec['steps'], 'clocks': clocks,
            'roofline': rec.get('roofline'), 'stages_ms': rec.get('stages_ms'),
            'ms_allreduce': rec['ms_allreduce'], 'ms_per_step_no_comm': rec['ms_per_step_no_comm'],
            'ms_allreduce_collective_only': rec.get('ms_allreduce_collective_only'),
            'losses': rec['losses'], 'launch_mode': rec['launch_mode'],
            'train_graph_check': rec['train_graph_check'],
        }
        print(json.dumps(out))
    if dist:
        dist.destroy_process_group()
    return 0


def stress_boxes(batch, img_hw, seed, clustered=False, per_level=4000, levels=5):
    """SURVEY §8d cfg 5 ("bypass anchors"): per image `levels` x `per_level` boxes, centres
    U(image), sides log-U(8*2^l, 64*2^l), unique scores in (0,1); `clustered`: centres are 200
    seeds + N(0, 4 px) jitter (long suppression chains, clustered RoIs)."""
    rng = np.random.RandomState(seed)
    H, W = img_hw
    out_b, out_s, out_l = [], [], []
    for _ in range(batch):
        bs, ls = [], []
        for l in range(levels):
            n = per_level
            if clustered:
                seeds = rng.rand(200, 2) * [W, H]
                ctr = seeds[rng.randint(0, 200, n)] + rng.normal(0, 4, (n, 2))
            else:
                ctr = rng.rand(n, 2) * [W, H]
            wh = np.exp(rng.uniform(np.log(8 * 2 ** l), np.log(64 * 2 ** l), (n, 2)))
            b = np.concatenate([ctr - wh / 2, ctr + wh / 2], 1)
            b[:, 0::2] = b[:, 0::2].clip(0, W)
            b[:, 1::2] = b[:, 1::2].clip(0, H)
            bs.append(b)
            ls.append(np.full(n, l))
        n_all = levels * per_level
        out_b.append(np.concatenate(bs).astype(np.float32))
        out_l.append(np.concatenate(ls).astype(np.int64))
        out_s.append(((rng.permutation(n_all) + 0.5) / n_all).astype(np.float32))
    return out_b, out_s, out_l


def bench_stress(args, rank, world, local_rank):
    """configs[4]: proposal stress test — 5 FPN levels x 4000 pre-NMS proposals, 2000 post-NMS
    RoIs per image, all RoIs through RoIAlign.
      --stress-input boxes | clustered  (SURVEY §8d as written): random boxes bypass the anchor
          decode; per image batched_nms (ids = level, K = 20 000) -> first 2000 -> RoIAlign;
      --stress-input anchors: the same sizes through brcnn_rpn_get_bboxes on the COCO pyramid
          (levels 3/4 only hold 2 457 / 693 anchors, K = 15 150).
    One CUDA graph per step."""
    from boosting_rcnn_b200 import _lib, configs, ops
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    dist = _dist_setup(world, dev)
    lib = _lib.load()
    geom = configs.IMAGE_GEOMETRY['coco']
    B = args.batch
    rpn_head, roi_head, model = configs.build_hot_path('coco')
    rpn_head = rpn_head.to(dev).eval()
    A, C = rpn_head.num_anchors, 256
    sizes, h_feats, h_cls, h_box, h_iou = make_inputs(B, geom['pad_shape'][:2], A, C,
                                                      seed=99 + rank, pin=False)
    metas = img_metas_for(B, geom)
    feats = [t.to(dev).contiguous(memory_format=torch.channels_last) for t in h_feats]
    cfg = dict(nms_pre=4000, max_per_img=2000, nms=dict(type='nms', iou_threshold=0.7),
               min_bbox_size=0)
    scales = [1.0 / s for s in STRIDES]
    mode = args.stress_input
    M = cfg['max_per_img']
    if mode == 'anchors':
        cls, box, iou = ([t.to(dev) for t in ts] for ts in (h_cls, h_box, h_iou))

        @torch.no_grad()
        def proposals():
            props = rpn_head.get_bboxes_padded(cls, box, iou, metas, cfg=cfg)
            return props.boxes, props.num
    else:
        bx, sc, lv = stress_boxes(B, geom['img_shape'][:2], 4242 + rank, clustered=(mode == 'clustered'))
        d_bx = [torch.from_numpy(b).to(dev) for b in bx]
        d_sc = [torch.from_numpy(x).to(dev) for x in sc]
        d_lv = [torch.from_numpy(x).to(dev) for x in lv]

        # mmcv batched_nms per image, ids = pyramid level (no host sync).  The images are
        # independent: one stream each, so the per-image cluster kernels (8 SMs each) run side by
        # side instead of one after the other (forked from / joined to the calling stream, which
        # is also how the calls are captured into the step's CUDA graph)
        side = [torch.cuda.Stream(device=dev) for _ in range(B)]

        @torch.no_grad()
        def proposals():
            cur = torch.cuda.current_stream()
            boxes = torch.zeros((B, M, 5), dtype=torch.float32, device=dev)
            nums = [None] * B
            for b in range(B):
                side[b].wait_stream(cur)
                with torch.cuda.stream(side[b]):
                    dets, _, num = ops._nms_raw(d_bx[b], d_sc[b], d_lv[b], 0.7, 0, num_ids=5,
                                                max_num=M)
                    boxes[b] = dets[:M]
                    nums[b] = num
            for st in side:
                cur.wait_stream(st)
            return boxes, torch.cat(nums).clamp_(max=M)

    @torch.no_grad()
    def fn():
        boxes, num = proposals()
        rois, _ = ops.bbox2roi_padded(boxes, num)
        return boxes, num, rois, ops.roi_extract(feats, rois, scales, 7, channels_last_out=True)

    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    l0 = lib.brcnn_launch_count()
    with torch.cuda.graph(g):
        boxes, num, rois, rf = fn()
    per_step = int(lib.brcnn_launch_count() - l0)
    K, W = args.steps, max(args.warmup, 3)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms = _timed_steps(g.replay, K, W, world, dist, dev)
    clocks = sampler.stop() if sampler else None
    if rank == 0:
        peak, peak_src = _peak()
        with torch.no_grad():
            t_prop = _graph_time(proposals, 10, dev)
            t_roi = _graph_time(lambda: ops.roi_extract(feats, rois, scales, 7,
                                                        channels_last_out=True), 10, dev)
        rois_h = rois.cpu().numpy()
        feat_bytes = sum(f.numel() * 4 for f in feats)
        out_bytes = rois_h.shape[0] * C * 49 * 4
        un = roi_union_bytes(rois_h, sizes, C)
        nbytes = out_bytes + un
        ach = nbytes / (t_roi * 1e-3) / 1e9
        survey = out_bytes + min(roi_footprint_bytes(rois_h, sizes, C), feat_bytes)
        print(json.dumps({
            'metric': 'proposal stress images/s', 'value': B * world * K / (ms * 1e-3),
            'unit': 'images/s', 'n_gpus': world, 'steps': K, 'warmup': W, 'ms_per_step': ms / K,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic',
            'config': dict(workload=f'proposal stress ({mode}): 5 levels x 4000 pre-NMS proposals, '
                                    f'2000 post-NMS RoIs/img through RoIAlign, {B} images per GPU '
                                    f'(BASELINE configs[4])',
                           images_per_gpu=B, rpn=cfg, input=mode, rois=int(rois_h.shape[0]),
                           live_rois=int((rois_h[:, 0] >= 0).sum()),
                           kept_per_image=[int(v) for v in num.cpu().tolist()]),
            'gpu_launches': per_step * K, 'clocks': clocks,
            'roofline': dict(kernel='roi_align_fwd3_kernel', bound='hbm', achieved=ach, peak=peak,
                             unit='GB/s', frac=ach / peak, traffic=None, peak_source=peak_src,
                             algorithmic_bytes_per_launch=nbytes, ms_per_launch=t_roi,
                             frac_survey_formula=survey / (t_roi * 1e-3) / 1e9 / peak,
                             bytes_definition='output bytes + union of footprint pixels per '
                                              '(image, level)'),
            'stages_ms': {'proposals_nms_20000_to_2000': t_prop, 'roi_align_fwd': t_roi}}))
    if dist:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    # a hung rank (GPU or collective) must not sit in NCCL's 10-minute timeout: dump every
    # thread's stack and leave after BENCH_WATCHDOG_S seconds (a normal run takes 1-3 minutes)
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get('BENCH_WATCHDOG_S', '420')), exit=True)
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if world > 1 and args.impl == 'reference' and rank != 0:
        return 0  # the CPU arm runs on rank 0 only
    if args.impl == 'b200' and args.mode != 'infer':
        assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
        return (bench_train if args.mode == 'train' else bench_stress)(args, rank, world, local_rank)

    from boosting_rcnn_b200 import configs
    geom = configs.IMAGE_GEOMETRY[args.cfg]
    pad_hw = geom['pad_shape'][:2]
    B = args.batch
    torch.manual_seed(0)
    rpn_head, roi_head, model = configs.build_hot_path(args.cfg)
    if args.rpn_max_per_img > 0:
        model['test_cfg']['rpn']['max_per_img'] = args.rpn_max_per_img
        rpn_head.test_cfg.max_per_img = args.rpn_max_per_img
    A, C = rpn_head.num_anchors, roi_head.bbox_roi_extractor.out_channels
    threads = os.cpu_count() or 1
    base_cfg = dict(workload=WORKLOAD if args.cfg == 'utdac' else f'{args.cfg} inference',
                    images_per_gpu=B, rpn=model['test_cfg']['rpn'], rcnn=model['test_cfg']['rcnn'],
                    head='2-fc shared head on torch/cuBLAS fp32 (TF32 off)',
                    l2_policy='per-step inputs (>=440 MB/GPU) exceed the 126 MB L2',
                    parallelism=f'image-sharded x{world}, no data-path collective',
                    concurrency='one CUDA graph per step; two consecutive batches in flight on two '
                                'streams (the memory-bound phase of batch i+1 runs beside the cuBLAS '
                                'head of batch i); value_single_stream = one batch at a time')

    # ------------------------------------------------------------------ CPU arm
    if args.impl == 'reference':
        n_img = args.cpu_images or max(threads, B)     # every host thread gets an image
        sizes, feats, cls, box, iou = make_inputs(min(n_img, B), pad_hw, A, C, seed=1234, pin=False)
        ref = CpuReference(rpn_head, roi_head, model, geom, threads)
        for _ in range(max(args.warmup, 0)):
            ref.run(feats, cls, box, iou, min(n_img, threads))
        t = 0.0
        for _ in range(args.steps):
            dt, _ = ref.run(feats, cls, box, iou, n_img)
            t += dt
        val = n_img * args.steps / t
        sample = f'{n_img} images/step x {args.steps} steps, one image per thread'
        print(json.dumps({
            'impl': 'reference', 'metric': 'RoI-path images/s @1333x800', 'value': val,
            'unit': 'images/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': 1e3 * t / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': base_cfg,
            'cpu_baseline': {'value': val, 'unit': 'images/s', 'cores': threads, 'kind': 'port',
                             'sample': sample},
            'e2e': {'value': val, 'unit': 'images/s', 'h2d_bytes_per_step': 0,
                    'd2h_bytes_per_step': 0}}))
        return 0

    # ------------------------------------------------------------------ GPU arm
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
        # NUMA-local pinned buffers for the end-to-end leg (one process per GPU)
        from boosting_rcnn_b200.dist import bind_to_gpu_cpus
        orig_affinity = os.sched_getaffinity(0)
        cpus = bind_to_gpu_cpus(local_rank)
        if cpus:
            torch.set_num_threads(max(1, min(len(cpus), 8)))
    from boosting_rcnn_b200 import _lib, ops
    from boosting_rcnn_b200.registry import ConfigDict
    lib = _lib.load()
    sampler = ClockSampler(local_rank) if rank == 0 else None   # covers every timed loop
    # ---- configs[2] training step at this N (gradient all-reduce over NVLink when N > 1).
    # Runs first: it is an eager step that allocates through the default caching allocator,
    # which slows down 3x once the inference graphs below hold most of the HBM in private pools
    train = None
    if not args.no_train_record:
        import torch.distributed as tdist
        try:
            train = train_record(args, rank, world, local_rank, tdist if world > 1 else None,
                                 cfg_name='coco', B=2, K=min(max(args.steps, 10), 20), W=3,
                                 with_stages=(rank == 0))
        except Exception as e:  # noqa: BLE001 - the sub-record must not take the headline down
            import traceback
            traceback.print_exc()
            train = {'error': f'{type(e).__name__}: {e}'} if rank == 0 else None
        torch.cuda.empty_cache()
    rpn_head, roi_head = rpn_head.to(dev).eval(), roi_head.to(dev).eval()
    sizes, h_feats, h_cls, h_box, h_iou = make_inputs(B, pad_hw, A, C, seed=1234 + rank, pin=True)
    metas = img_metas_for(B, geom)
    to_dev = lambda ts: [t.to(dev, non_blocking=True) for t in ts]
    d_feats, d_cls, d_box, d_iou = to_dev(h_feats), to_dev(h_cls), to_dev(h_box), to_dev(h_iou)
    d_feats_cl = [f.contiguous(memory_format=torch.channels_last) for f in d_feats]
    torch.cuda.synchronize()
    test_rcnn = model['test_cfg']['rcnn']
    from boosting_rcnn_b200.graph import HostPipeline, HotPathGraph

    # the step: one CUDA graph over the whole batch; --no-graph
    # launches the same kernels one by one from Python
    if args.no_graph:
        @torch.no_grad()
        def _eager(feats):
            props = rpn_head.get_bboxes_padded(d_cls, d_box, d_iou, metas)
            return roi_head.simple_test_bboxes_padded(feats, metas, props, test_rcnn, rescale=True)
        step, step_cl = (lambda: _eager(d_feats)), (lambda: _eager(d_feats_cl))
        dual = g_tf32 = None
        l0 = lib.brcnn_launch_count()
        step()
        launches_per_step = int(lib.brcnn_launch_count() - l0)
    else:
        g_nchw = HotPathGraph(rpn_head, roi_head, metas, d_feats, d_cls, d_box, d_iou,
                              rcnn_test_cfg=test_rcnn, rescale=True)
        step = g_nchw.replay
        launches_per_step = g_nchw.launches_per_replay
        step_cl = dual = g_tf32 = None
        if not args.lean:
            g_cl = HotPathGraph(rpn_head, roi_head, metas, d_feats_cl, d_cls, d_box, d_iou,
                                rcnn_test_cfg=test_rcnn, rescale=True)
            step_cl = g_cl.replay
            # two batches in flight on two streams (graph.py::DualStreamRunner)
            from boosting_rcnn_b200.graph import DualStreamRunner
            g_nchw2 = HotPathGraph(rpn_head, roi_head, metas, d_feats, d_cls, d_box, d_iou,
                                   rcnn_test_cfg=test_rcnn, rescale=True)
            dual = DualStreamRunner([g_nchw, g_nchw2])

            # informational only: the reference pins PyTorch 1.7, whose default lets cuBLAS use
            # TF32 for the 2-fc head on Ampere+; the headline keeps IEEE fp32 GEMMs
            torch.backends.cuda.matmul.allow_tf32 = True
            g_tf32 = HotPathGraph(rpn_head, roi_head, metas, d_feats, d_cls, d_box, d_iou,
                                  rcnn_test_cfg=test_rcnn, rescale=True)
            torch.backends.cuda.matmul.allow_tf32 = False

    # end to end: pinned host buffers -> H2D -> graph -> D2H, double buffered
    pipe = HostPipeline(rpn_head, roi_head, metas, (h_feats, h_cls, h_box, h_iou),
                        rcnn_test_cfg=test_rcnn, rescale=True, slots=1 if args.lean else 2)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, drain=None):
        for _ in range(warmup):
            fn()
        if drain:
            drain()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if drain:
            drain()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    class E2E:
        """K steps through HostPipeline; every step copies its inputs from
        pinned host memory and reads its detections back to the host."""
        prev = None

        def step(self):
            t = pipe.submit(h_feats, h_cls, h_box, h_iou)
            if self.prev is not None:
                self.prev.result()
            self.prev = t

        def drain(self):
            if self.prev is not None:
                self.prev.result()
                self.prev = None
            # e1 is recorded on the default stream: make it wait for the pipeline
            torch.cuda.current_stream().wait_stream(pipe.compute_stream)

    W, K = max(args.warmup, 3), args.steps
    ms_single = timed(step, K, W)
    launches = launches_per_step * K
    ms_cl = timed(step_cl, K, W) if step_cl is not None else None
    ms_tf32 = timed(g_tf32.replay, K, W) if (not args.no_graph and g_tf32 is not None) else None
    ms_dual = timed(dual.step, K, W, drain=dual.drain) if (not args.no_graph and dual is not None) else None
    # headline: throughput with two batches in flight (one captured graph per stream);
    # --no-graph: the eager single-stream number
    ms = ms_dual if ms_dual is not None else ms_single
    # the same K-step block four more times: median of 5 (variance of a 45 ms timed region)
    head_fn, head_drain = (dual.step, dual.drain) if ms_dual is not None else (step, None)
    blocks = sorted([ms] + [timed(head_fn, K, 0, drain=head_drain) for _ in range(4)])
    ms_median = blocks[2]

    # ---- strong scaling (BASELINE configs[1] as written: ONE batch of 16 images sharded over
    # the GPUs, SURVEY §8e 16/8/4/2 per GPU): rank r owns images shard_range(16, r, world)
    strong = None
    if not args.no_graph:
        from boosting_rcnn_b200.dist import shard_range
        tot = args.strong_total
        lo, hi = shard_range(tot, rank, world)
        lo, hi = min(lo, B), min(hi, B)
        nb = hi - lo
        g_strong = None
        if nb > 0:
            sl = lambda ts: [t[lo:hi].contiguous() for t in ts]
            g_strong = HotPathGraph(rpn_head, roi_head, metas[lo:hi], sl(d_feats), sl(d_cls),
                                    sl(d_box), sl(d_iou), rcnn_test_cfg=test_rcnn, rescale=True)
        ms_strong = timed(g_strong.replay if g_strong is not None else (lambda: None), K, W)
        strong = dict(images_total=tot, images_per_gpu=nb, value=tot * K / (ms_strong * 1e-3),
                      unit='images/s', ms_per_step=ms_strong / K, scaling='strong',
                      note='one CUDA graph per step on one stream, inputs resident; no data-path '
                           'collective (detections stay on their rank)')
    verified = None
    if ms_dual is not None:
        # the two in-flight graphs must reproduce the single-stream detections bit for bit
        ref_out = [o.clone() for o in g_nchw.replay()]
        torch.cuda.synchronize()
        for _ in range(4):
            dual.step()
        dual.drain()
        torch.cuda.synchronize()
        verified = all(torch.equal(a, b) for g in (g_nchw, g_nchw2)
                       for a, b in zip(g.outputs, ref_out))
        assert verified, 'two-stream replay changed the detections'
    e2e = E2E()
    k_e2e = max(K // 2, 4)
    ms_e2e = timed(e2e.step, k_e2e, 3, drain=e2e.drain)
    # ceiling of the end-to-end leg: plain pinned-host -> device cudaMemcpyAsync of the largest
    # input map, all ranks at once (the host's aggregate H2D bandwidth is shared by the GPUs)
    probe_src, probe_dst = h_feats[0], d_feats[0]
    ms_probe = timed(lambda: probe_dst.copy_(probe_src, non_blocking=True), 8, 2)
    h2d_probe_gbs = probe_src.numel() * 4 * 8 / (ms_probe * 1e-3) / 1e9

    # -------------------------------------------------- per-stage device times
    stage_ms, roof = {}, None
    if rank == 0:
        with torch.no_grad():
            props = rpn_head.get_bboxes_padded(d_cls, d_box, d_iou, metas)
            rois, prior = ops.bbox2roi_padded(props.boxes, props.num)
            scales = [1.0 / s for s in STRIDES]
            nhwc = ops.pyramid_to_nhwc(d_feats)
            feats_cl = [t.permute(0, 3, 1, 2) for t in nhwc]
            rf = ops.roi_extract(feats_cl, rois, scales, 7, channels_last_out=True)
            cs, bp = roi_head.bbox_head(rf)
            hw, sf = roi_head._img_consts(metas, dev)
            rp = roi_head.bbox_head.rcnn_params(B, props.boxes.size(1), ConfigDict(test_rcnn), True, True)
            stages = {
                'rpn_get_bboxes': lambda: rpn_head.get_bboxes_padded(d_cls, d_box, d_iou, metas),
                'nchw_to_nhwc_x5': lambda: ops.pyramid_to_nhwc(d_feats),
                'roi_align_fwd': lambda: ops.roi_extract(feats_cl, rois, scales, 7,
                                                         channels_last_out=True),
                'fc_head_cublas': lambda: roi_head.bbox_head(rf),
                'rcnn_get_bboxes': lambda: ops.rcnn_get_bboxes(rp, rois, prior, props.num, cs, bp,
                                                               hw, sf),
            }
            # device time of each stage: captured alone in a CUDA graph (no
            # host launch overhead in the number), L2 flushed before every rep
            flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
            for name, fn in stages.items():
                for _ in range(2):
                    fn()
                torch.cuda.synchronize()
                if args.no_graph:
                    run = fn
                else:
                    sg = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(sg):
                        fn()
                    run = sg.replay
                run()
                evs = []
                for _ in range(10):
                    flush.zero_()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    run()
                    b.record()
                    evs.append((a, b))
                torch.cuda.synchronize()
                stage_ms[name] = float(np.mean([a.elapsed_time(b) for a, b in evs]))
            del flush
        peak, peak_src = _peak()
        rois_h = rois.cpu().numpy()
        n_live = int((rois_h[:, 0] >= 0).sum())
        feat_bytes = sum(f.numel() * 4 for f in d_feats)
        out_bytes = rois_h.shape[0] * C * 49 * 4
        fp = roi_footprint_bytes(rois_h, sizes, C)
        un = roi_union_bytes(rois_h, sizes, C)
        # ncu dram__bytes_read+write per launch of the same command (profiles/)
        traffic = {}
        tpath = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(f'{args.cfg}_b{B}', {})
        # algorithmic bytes per launch: output bytes + the UNION of the footprint pixels per
        # (image, level) (what must cross HBM at least once).  SURVEY §8d's formula
        # (sum of footprints capped at the map bytes) is reported beside it.
        kern = {
            'roi_align_fwd3_kernel': dict(bytes=out_bytes + un, ms=stage_ms['roi_align_fwd'],
                                          live_rois=n_live, rois=int(rois_h.shape[0]),
                                          union_feature_bytes=un,
                                          bytes_survey_formula=out_bytes + min(fp, feat_bytes)),
            'transpose_multi_kernel': dict(bytes=2 * feat_bytes, ms=stage_ms['nchw_to_nhwc_x5']),
        }
        for k, v in kern.items():
            v['achieved'] = v['bytes'] / (v['ms'] * 1e-3) / 1e9
            v['frac'] = v['achieved'] / peak
            v['traffic'] = traffic.get(k)
            if 'bytes_survey_formula' in v:
                v['frac_survey_formula'] = v['bytes_survey_formula'] / (v['ms'] * 1e-3) / 1e9 / peak
        # headline roofline kernel = the RoIAlign forward, the largest of the path's own compute
        # kernels (the NCHW->NHWC transpose beside it is a layout adapter that disappears when
        # the neck runs channels_last; it is reported under `kernels`)
        dom = 'roi_align_fwd3_kernel'
        d = kern[dom]
        roof = dict(kernel=dom, bound='hbm', achieved=d['achieved'], peak=peak, unit='GB/s',
                    frac=d['frac'], traffic=d['traffic'], peak_source=peak_src,
                    algorithmic_bytes_per_launch=d['bytes'], ms_per_launch=d['ms'],
                    frac_survey_formula=d.get('frac_survey_formula'),
                    bytes_definition='output bytes + union of footprint pixels per (image, level)',
                    kernels=kern)

    # ---- own-kernel stage times at the strong-scaling per-GPU batch (rank 0's shard) ----
    if rank == 0 and strong is not None and strong['images_per_gpu'] not in (0, B):
        nb = strong['images_per_gpu']
        with torch.no_grad():
            sl = lambda ts: [t[:nb].contiguous() for t in ts]
            s_feats, s_cls, s_box, s_iou = sl(d_feats), sl(d_cls), sl(d_box), sl(d_iou)
            s_metas = metas[:nb]
            sprops = rpn_head.get_bboxes_padded(s_cls, s_box, s_iou, s_metas)
            srois, sprior = ops.bbox2roi_padded(sprops.boxes, sprops.num)
            s_cl = [t.permute(0, 3, 1, 2) for t in ops.pyramid_to_nhwc(s_feats)]
            srf = ops.roi_extract(s_cl, srois, scales, 7, channels_last_out=True)
            scs, sbp = roi_head.bbox_head(srf)
            shw, ssf = roi_head._img_consts(s_metas, dev)
            srp = roi_head.bbox_head.rcnn_params(nb, sprops.boxes.size(1), ConfigDict(test_rcnn),
                                                 True, True)
            strong['stages_ms'] = {
                'rpn_get_bboxes': _graph_time(
                    lambda: rpn_head.get_bboxes_padded(s_cls, s_box, s_iou, s_metas), 10, dev),
                'nchw_to_nhwc_x5': _graph_time(lambda: ops.pyramid_to_nhwc(s_feats), 10, dev),
                'roi_align_fwd': _graph_time(
                    lambda: ops.roi_extract(s_cl, srois, scales, 7, channels_last_out=True), 10, dev),
                'fc_head_cublas': _graph_time(lambda: roi_head.bbox_head(srf), 10, dev),
                'rcnn_get_bboxes': _graph_time(
                    lambda: ops.rcnn_get_bboxes(srp, srois, sprior, sprops.num, scs, sbp, shw, ssf),
                    10, dev),
            }
            st = strong['stages_ms']
            strong['own_kernels_ms'] = sum(v for k, v in st.items() if k != 'fc_head_cublas')
            strong['limiter'] = max(st, key=st.get)

    clocks = sampler.stop() if sampler else None

    # ------------------------------------------------------------ CPU baseline
    cpu = None
    if world > 1:
        os.sched_setaffinity(0, orig_affinity)   # the CPU baseline may use every host core
        torch.set_num_threads(threads)
    if rank == 0 and not args.no_cpu_baseline:
        n_img = args.cpu_images or max(threads, B)     # every host thread gets an image
        ref = CpuReference(rpn_head, roi_head, model, geom, threads)
        ref.run(h_feats, h_cls, h_box, h_iou, min(2, n_img))  # warm-up
        reps, t = 0, 0.0
        while t < 8.0 and reps < 20:
            dt, _ = ref.run(h_feats, h_cls, h_box, h_iou, n_img)
            t += dt
            reps += 1
        cpu = {'value': n_img * reps / t, 'unit': 'images/s', 'cores': threads, 'kind': 'port',
               'sample': f'{n_img} images x {reps} passes of the same workload, one image per '
                         f'host thread (oracle C restatement + torch CPU 2-fc head)'}

    if rank == 0:
        h2d = sum(t.numel() * 4 for ts in (h_feats, h_cls, h_box, h_iou) for t in ts)
        mp = test_rcnn['max_per_img']
        d2h = B * mp * 5 * 4 + B * mp * 8 + B * 4
        out = {
            'metric': 'RoI-path images/s @1333x800', 'value': B * world * K / (ms * 1e-3),
            'unit': 'images/s', 'n_gpus': world, 'steps': K, 'warmup': W,
            'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': dict(base_cfg, feat_layout='NCHW-contiguous FPN maps in (reference neck '
                           'layout); NCHW->NHWC conversion kernels are inside the timed step'),
            'value_single_stream_channels_last_feats': (B * world * K / (ms_cl * 1e-3)) if ms_cl else None,
            'value_single_stream': B * world * K / (ms_single * 1e-3),
            'two_stream_outputs_bit_identical': verified,
            'ms_per_step_single_stream': ms_single / K,
            'value_median_of_5_blocks': B * world * K / (ms_median * 1e-3),
            'strong_scaling': strong, 'train': train,
            'value_single_stream_tf32_head_informational': (B * world * K / (ms_tf32 * 1e-3)) if ms_tf32 else None,
            'e2e': {'value': B * world * k_e2e / (ms_e2e * 1e-3), 'unit': 'images/s',
                    'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'ms_per_step': ms_e2e / k_e2e,
                    'h2d_gbs_per_gpu_achieved': h2d / (ms_e2e / k_e2e * 1e-3) / 1e9,
                    'h2d_gbs_per_gpu_memcpy_probe': h2d_probe_gbs,
                    'note': 'bound by the host->device copy of the step inputs; the probe is a '
                            'plain pinned cudaMemcpyAsync run by all ranks at once'},
            'gpu_launches': int(launches), 'clocks': clocks, 'roofline': roof,
            'stages_ms': stage_ms, 'cpu_baseline': cpu,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
