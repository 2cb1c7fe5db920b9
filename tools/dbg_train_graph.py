"""Probe for graph.RcnnTrainGraph capture: runs a few eager R-CNN training steps, then the
captured one, in a subprocess per variant (a failed capture poisons the process).
    python tools/dbg_train_graph.py            # all variants
    python tools/dbg_train_graph.py side gc    # one variant in-process
Variants: 'default' / 'side' (stream the steps run on), 'gc' (collect before capture),
'noeager' (no eager step before the capture)."""
import os
import subprocess
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]


def one(flags):
    import gc
    import warnings

    import torch

    import synth
    from boosting_rcnn_b200 import configs, graph
    cuda = torch.device('cuda', 0)
    torch.manual_seed(5)
    _, roi, _ = configs.build_hot_path('coco', train=True)
    roi = roi.to(cuda).train()
    sizes = synth.featmap_sizes(256, 320)
    gts_h, labels_h, plist_h = synth.rcnn_train_case('plenty')
    B = len(plist_h)
    gts = [torch.from_numpy(g).to(cuda) for g in gts_h]
    labels = [torch.from_numpy(l).to(cuda) for l in labels_h]
    plist = [torch.from_numpy(p).to(cuda) for p in plist_h]
    metas = [dict(img_shape=(250, 317, 3), pad_shape=(256, 320, 3), scale_factor=[1, 1, 1, 1])] * B
    params = [p for p in roi.parameters() if p.requires_grad]
    feats = [torch.from_numpy(f).to(cuda).requires_grad_(True)
             for f in synth.fpn_feats(B, 256, sizes, seed=3)]

    # say which capture is running
    orig_enter, orig_exit = torch.cuda.graph.__enter__, torch.cuda.graph.__exit__
    n = [0]

    def enter(self):
        n[0] += 1
        print(f'  capture {n[0]} begins', flush=True)
        return orig_enter(self)

    def exit_(self, *a):
        print(f'  capture {n[0]} ends (exc={a[0]})', flush=True)
        return orig_exit(self, *a)
    torch.cuda.graph.__enter__, torch.cuda.graph.__exit__ = enter, exit_
    orig_mgc = torch.cuda.make_graphed_callables

    def mgc(*a, **k):
        if 'gc' in flags:
            print('  gc.collect ->', gc.collect(), flush=True)
        return orig_mgc(*a, **k)
    torch.cuda.make_graphed_callables = mgc
    warnings.simplefilter('always')

    def step(use_graph):
        for p in params + feats:
            p.grad = None
        roi.train_graph = use_graph
        torch.manual_seed(9)
        out = roi.forward_train(feats, metas, plist, gts, labels)
        (out['loss_cls'] + out['loss_bbox']).backward()
        torch.cuda.synchronize()
        return [out['loss_cls'].item(), out['loss_bbox'].item(),
                float(sum(p.grad.double().abs().sum() for p in params)),
                float(sum(f.grad.double().abs().sum() for f in feats))]

    def run():
        ref = None
        if 'noeager' not in flags:
            ref = step(False)
            print('  eager  ', ref, flush=True)
        try:
            for i in range(3):
                got = step(True)
                print('  graphed', got, 'graphs:', [bool(g) for g in roi._train_graphs.values()],
                      flush=True)
        except Exception:   # noqa: BLE001
            traceback.print_exc()
    if 'side' in flags:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            run()
    else:
        run()


if __name__ == '__main__':
    if len(sys.argv) > 1:
        one(sys.argv[1:])
    else:
        for v in (['default'], ['default', 'gc'], ['default', 'noeager'], ['side'], ['side', 'gc']):
            print('variant', v, flush=True)
            r = subprocess.run([sys.executable, os.path.abspath(__file__)] + v, timeout=120,
                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            print(r.stdout[-2500:], flush=True)
