"""CUDA-graph execution of the inference hot path for fixed shapes.

The proposal-to-detection step is ~20 short kernels; launched one by one from
Python it is host-bound (the kernels take ~1 ms, the launches ~1.5 ms).  The
whole step is sync-free by construction (fixed-capacity buffers + device
counters), so it captures into one CUDA graph:

    g = HotPathGraph(rpn_head, roi_head, img_metas, feats, cls, box, iou, rescale=True)
    det, lab, num = g.replay()          # reads the tensors given at capture time

``HostPipeline`` puts two such graphs behind a double-buffered host->device
copy stream so that the H2D of step i+1 overlaps the compute of step i; this
is the end-to-end entry point for callers whose inputs live in host memory.
"""
import warnings

import torch
import torch.nn as nn

from . import _lib, ops
from .registry import ConfigDict


class HotPathGraph:
    """One captured pass RPN outputs + FPN maps -> detections over a batch.

    The input tensors passed to the constructor are the graph's static inputs:
    refill them in place (``copy_``) between replays."""

    def __init__(self, rpn_head, roi_head, img_metas, feats, cls_scores, bbox_preds, iou_preds,
                 rcnn_test_cfg=None, rescale=True, warmup=2, stream=None):
        assert torch.cuda.is_available(), 'HotPathGraph needs a CUDA device'
        lib = _lib.load()
        self.inputs = (list(feats), list(cls_scores), list(bbox_preds), list(iou_preds))
        cfg = ConfigDict(rcnn_test_cfg if rcnn_test_cfg is not None else roi_head.test_cfg)

        @torch.no_grad()
        def fn():
            f, c, b, i = self.inputs
            props = rpn_head.get_bboxes_padded(c, b, i, img_metas)
            return roi_head.simple_test_bboxes_padded(f, img_metas, props, cfg, rescale=rescale)

        self._fn = fn
        cur = torch.cuda.current_stream()
        side = stream if stream is not None else torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):   # allocator / cuBLAS / attribute warm-up
                fn()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        l0 = lib.brcnn_launch_count()
        with torch.cuda.graph(self.graph, stream=side):
            self.outputs = fn()
        #: libbrcnn kernels inside one replay (the rest are cuBLAS + bias adds)
        self.launches_per_replay = int(lib.brcnn_launch_count() - l0)

    def replay(self):
        self.graph.replay()
        return self.outputs

    def eager(self):
        return self._fn()


class DualStreamRunner:
    """Two (or more) captured steps in flight, one stream each.

    One step is a short bandwidth/latency-bound phase (proposals, layout hand-off,
    RoIAlign, post-processing: our kernels) around a long compute-bound phase (the 2-fc
    head on cuBLAS SGEMM).  Consecutive batches are independent, so replaying them
    alternately on two streams lets the memory-bound phase of batch i+1 run beside the
    GEMMs of batch i.  Each stream owns its own graph (own workspaces and outputs); the
    input tensors may be shared (they are only read)."""

    def __init__(self, graphs):
        assert len(graphs) >= 2
        self.graphs = graphs
        self.streams = [torch.cuda.Stream() for _ in graphs]
        self._i = 0
        cur = torch.cuda.current_stream()
        for s in self.streams:
            s.wait_stream(cur)

    def step(self):
        k = self._i % len(self.graphs)
        self._i += 1
        with torch.cuda.stream(self.streams[k]):
            self.graphs[k].replay()
        return self.graphs[k].outputs

    def drain(self):
        """Make the current stream wait for both branches (call before recording the
        closing event / reading outputs)."""
        cur = torch.cuda.current_stream()
        for s in self.streams:
            cur.wait_stream(s)


class HostPipeline:
    """Double-buffered end-to-end pipeline for host-resident inputs.

    ``submit(feats, cls, box, iou)`` takes pinned host tensors, enqueues their
    H2D copies on a copy stream, the graph replay on a compute stream and the
    D2H of the detections into pinned host buffers; it returns a ticket whose
    ``result()`` blocks until that step's detections are on the host.  With two
    slots the copy of step i+1 overlaps the kernels of step i."""

    def __init__(self, rpn_head, roi_head, img_metas, example, rcnn_test_cfg=None, rescale=True,
                 slots=2):
        dev = next(roi_head.parameters()).device
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.compute_stream = torch.cuda.Stream(device=dev)
        self.slots = []
        for _ in range(slots):
            bufs = tuple([torch.empty_like(t, device=dev) for t in ts] for ts in example)
            g = HotPathGraph(rpn_head, roi_head, img_metas, *bufs, rcnn_test_cfg=rcnn_test_cfg,
                             rescale=rescale, stream=self.compute_stream)
            host_out = tuple(torch.empty_like(o, device='cpu').pin_memory() for o in g.outputs)
            self.slots.append(dict(bufs=bufs, graph=g, host_out=host_out,
                                   copied=torch.cuda.Event(), done=torch.cuda.Event(),
                                   busy=False))
        self.h2d_bytes = sum(t.numel() * t.element_size() for ts in example for t in ts)
        self.d2h_bytes = sum(o.numel() * o.element_size() for o in self.slots[0]['host_out'])
        self.launches_per_step = self.slots[0]['graph'].launches_per_replay
        self._next = 0

    class Ticket:
        def __init__(self, slot):
            self._slot = slot

        def result(self):
            self._slot['done'].synchronize()
            self._slot['busy'] = False
            return self._slot['host_out']

    def submit(self, feats, cls_scores, bbox_preds, iou_preds):
        s = self.slots[self._next]
        self._next = (self._next + 1) % len(self.slots)
        if s['busy']:
            # the slot's previous step must have left the device before its
            # input buffers are overwritten
            s['done'].synchronize()
        with torch.cuda.stream(self.copy_stream):
            for dst, src in zip(s['bufs'], (feats, cls_scores, bbox_preds, iou_preds)):
                for d, h in zip(dst, src):
                    d.copy_(h, non_blocking=True)
            s['copied'].record(self.copy_stream)
        with torch.cuda.stream(self.compute_stream):
            self.compute_stream.wait_event(s['copied'])
            s['graph'].replay()
            for h, o in zip(s['host_out'], s['graph'].outputs):
                h.copy_(o, non_blocking=True)
            s['done'].record(self.compute_stream)
        s['busy'] = True
        return HostPipeline.Ticket(s)

    def drain(self):
        for s in self.slots:
            if s['busy']:
                s['done'].synchronize()
                s['busy'] = False


class _RcnnTrainBody(nn.Module):
    """The sync-free tail of ``ProbRoIHead.forward_train`` on static inputs: sample + targets
    + prior (one launch) -> RoI extraction -> 2-fc head -> boost loss."""

    def __init__(self, roi_head, static, num_rows):
        super().__init__()
        self.roi_head = roi_head          # its parameters are this module's parameters
        self.static, self.num_rows = static, num_rows

    def forward(self, *feats):
        st, rh = self.static, self.roi_head
        h = rh.bbox_head
        rois, labels, label_weights, bbox_targets, bbox_weights, prior = ops.rcnn_sample_targets(
            st['proposals'], st['num_props'], st['gtb'], st['gtl'], st['num_gt'], st['gt_inds'],
            st['plan'], st['perm_pos'], st['perm_neg'], self.num_rows, h.num_classes,
            h.bbox_coder.means, h.bbox_coder.stds, rh.train_cfg.pos_weight)
        res = rh._bbox_forward(feats, rois)
        out = h.boost_loss(res['cls_score'], res['bbox_pred'], labels, label_weights,
                           bbox_targets, bbox_weights, prior, rh.gamma, rh.alpha, rh.reg_norm)
        return out['loss_cls'], out['loss_bbox'], out['acc']


class RcnnTrainGraph:
    """CUDA-graph replay of the R-CNN half of the training step (prob_roi_head.py:66-149 after
    the sampler's CPU ``randperm``): forward and backward are one graph each
    (``torch.cuda.make_graphed_callables``), so the ~45 eager launches of sample/targets,
    layout hand-off, RoIAlign, the 2-fc head, the boost loss and their backward become two
    replays.  The result is an ordinary autograd node: ``loss.backward()`` of the caller
    replays the backward graph, accumulates the head's parameter gradients and hands the
    gradients of the FPN maps on to the neck.

    Static shapes: a graph is keyed by (rows N, proposal capacity, GT capacity, feature-map
    shapes); ``ProbRoIHead`` pads the GT tensors to a multiple of 32 in this mode.  Any other
    combination builds another graph (or, if capture fails, the eager path runs).  Enabled
    with ``roi_head.train_graph = True``."""

    _NAMES = ('proposals', 'num_props', 'gtb', 'gtl', 'num_gt', 'gt_inds')

    def __init__(self, roi_head, feats, assigned, plan, perm_pos, perm_neg, num_rows):
        dev = assigned.proposals.device
        # the head's reusable assignment buffers are the graph's inputs as they are; anything
        # else is cloned and refilled by copies before every replay
        self.static = {k: (getattr(assigned, k) if assigned.is_static
                           else getattr(assigned, k).clone()) for k in self._NAMES}
        B, cap = perm_pos.shape
        # plan + both permutations cross the bus as ONE pinned buffer
        self._host = torch.empty((B * 5 + 2 * B * cap,), dtype=torch.int32).pin_memory()
        self._dev = torch.empty_like(self._host, device=dev)
        self._cuts = (B * 5, B * 5 + B * cap)
        a, b = self._cuts
        self.static.update(plan=self._dev[:a].view(B, 5), perm_pos=self._dev[a:b].view(B, cap),
                           perm_neg=self._dev[b:].view(B, cap))
        self._fill(assigned, plan, perm_pos, perm_neg, first=True)
        self.body = _RcnnTrainBody(roi_head, self.static, int(num_rows))
        # An autograd graph of an earlier eager step that is still alive (kept by a reference
        # cycle somewhere in the caller) pins the parameters' AccumulateGrad nodes to the
        # stream that step ran on; autograd would then synchronise the capturing stream with
        # it — illegal for the default stream, the capture is invalidated.  Collect first.
        import gc
        gc.collect()
        lib = _lib.load()
        l0 = lib.brcnn_launch_count()
        torch.cuda.make_graphed_callables(self.body, tuple(feats), num_warmup_iters=3,
                                          allow_unused_input=True)
        # kernels of the library inside one forward + backward replay (3 warm-up passes + 1
        # capture were recorded)
        self.launches_per_step = int(lib.brcnn_launch_count() - l0) // 4

    def _fill(self, assigned, plan, perm_pos, perm_neg, first=False):
        a, b = self._cuts
        h = self._host
        h[:a] = plan.reshape(-1)
        h[a:b] = perm_pos.reshape(-1)
        h[b:] = perm_neg.reshape(-1)
        self._dev.copy_(h, non_blocking=True)
        if not first:
            todo = [k for k in self._NAMES
                    if getattr(assigned, k).data_ptr() != self.static[k].data_ptr()]
            if todo:
                torch._foreach_copy_([self.static[k] for k in todo],
                                     [getattr(assigned, k) for k in todo])

    @staticmethod
    def _key(feats, assigned, perm_pos, num_rows):
        return (int(num_rows), tuple(assigned.proposals.shape), int(assigned.gtb.size(1)),
                assigned.proposals.data_ptr() if assigned.is_static else 0,
                tuple(perm_pos.shape), str(assigned.proposals.device),
                tuple((tuple(f.shape), tuple(f.stride()), bool(f.requires_grad)) for f in feats))

    @classmethod
    def run(cls, roi_head, feats, assigned, plan, perm_pos, perm_neg, num_rows):
        """Losses of this step through the graph matching its shapes (built on first use);
        None when there is nothing to replay (no rows, capture failed): the caller then runs
        the eager path."""
        if num_rows <= 0 or not torch.is_grad_enabled():
            return None
        feats = tuple(feats[:roi_head.bbox_roi_extractor.num_inputs])
        key = cls._key(feats, assigned, perm_pos, num_rows)
        g = roi_head._train_graphs.get(key)
        if g is False:
            return None
        if g is None:
            try:
                g = cls(roi_head, feats, assigned, plan, perm_pos, perm_neg, num_rows)
            except Exception as e:  # noqa: BLE001 - capture is an optimisation, never fatal
                # (a failure in the warm-up passes is harmless; one inside the stream capture
                # itself can leave torch's CUDA generator registered with the dead capture, and
                # later random ops then raise — treat this warning as a bug report)
                warnings.warn(f'RcnnTrainGraph: capture failed ({type(e).__name__}: {e}); '
                              f'the eager training path runs instead')
                roi_head._train_graphs[key] = False
                return None
            roi_head._train_graphs[key] = g
        else:
            # the previous step's copies out of the pinned buffer must have left the host
            g._copied.synchronize()
            g._fill(assigned, plan, perm_pos, perm_neg)
        g._copied = torch.cuda.Event()
        g._copied.record()
        loss_cls, loss_bbox, acc = g.body(*feats)
        return dict(loss_cls=loss_cls, acc=acc, loss_bbox=loss_bbox)
