"""GPU parity: fused level-map + multi-level RoIAlign forward/backward vs the
oracle's restated mmcv roi_align + SingleRoIExtractor.  Bar: level ids
bit-exact; features and gradients <= 1e-5 relative (fp32)."""
import numpy as np
import pytest
import torch

import synth
from boosting_rcnn_b200 import ops
from oracle import oracle

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def _close(a, b, what):
    scale = max(float(np.abs(b).max()), 1e-6)
    err = float(np.abs(a - b).max())
    assert err <= RTOL * scale, f'{what}: max abs err {err:.3e} vs scale {scale:.3e}'


def _case(dev, batch, pad_hw, C, n_per_img, seed, clustered=False, channels_last=False,
          extra_rois=None, channels_last_out=False):
    sizes = synth.featmap_sizes(*pad_hw)
    scales = [1.0 / s for s in synth.STRIDES]
    feats = synth.fpn_feats(batch, C, sizes, seed=seed)
    rois = synth.random_rois(batch, n_per_img, pad_hw[0], pad_hw[1], seed=seed + 1,
                             clustered=clustered)
    if extra_rois is not None:
        rois = np.concatenate([rois, extra_rois], 0).astype(np.float32)
    tf = [torch.from_numpy(f).to(dev) for f in feats]
    if channels_last:
        tf = [f.contiguous(memory_format=torch.channels_last) for f in tf]
    tf = [f.requires_grad_(True) for f in tf]
    out, lv = ops.roi_extract(tf, torch.from_numpy(rois).to(dev), scales, 7, 0, True, 56,
                              return_levels=True, channels_last_out=channels_last_out)
    if channels_last_out:   # (R,7,7,C) storage behind the reference's logical shape
        assert out.shape[1] == C and out.permute(0, 2, 3, 1).is_contiguous()
    ref, rlv = oracle.roi_extract_forward(feats, rois, scales)
    np.testing.assert_array_equal(lv.cpu().numpy().astype(np.int64), rlv)
    _close(out.detach().cpu().numpy(), ref, 'roi features')
    return tf, feats, rois, scales, out


def test_roi_forward_small(cuda):
    _case(cuda, 2, (256, 320), 64, 50, seed=0)


def test_roi_forward_channels_last_input(cuda):
    _case(cuda, 2, (256, 320), 256, 40, seed=1, channels_last=True)


def test_roi_forward_edge_rois(cuda):
    # degenerate, out-of-image, huge and level-threshold-straddling RoIs
    ex = []
    for s in (112.0, 224.0, 448.0, 896.0):
        for d in (-1e-3, 0.0, 1e-3):
            ex.append([0, 10, 10, 10 + s + d, 10 + s + d])
    ex += [[1, 5, 5, 5, 5], [1, 0, 0, 320, 256], [0, 300, 250, 330, 270], [1, 100, 50, 90, 40],
           [0, 2, 2, 318, 6], [1, -20, -20, 10, 10]]
    _case(cuda, 2, (256, 320), 32, 10, seed=2, extra_rois=np.array(ex, dtype=np.float32))


def test_roi_forward_full_size_utdac(cuda):
    # configs[1] geometry: 1344x800 maps, 256 channels, 256 RoIs / image
    _case(cuda, 2, (800, 1344), 256, 256, seed=3, clustered=True)


def test_roi_levels_operator(cuda):
    rois = synth.random_rois(1, 2000, 800, 1333, seed=9)
    lv = ops.map_roi_levels(torch.from_numpy(rois).to(cuda), 5, 56)
    np.testing.assert_array_equal(lv.cpu().numpy(), oracle.map_roi_levels(rois, 5, 56))
    assert lv.dtype == torch.int64


def test_roi_padding_rows_are_zero(cuda):
    sizes = synth.featmap_sizes(128, 128)
    feats = [torch.from_numpy(f).to(cuda) for f in synth.fpn_feats(1, 16, sizes, 0)]
    rois = torch.tensor([[0, 4, 4, 60, 60], [-1, 0, 0, 0, 0]], dtype=torch.float32, device=cuda)
    out, lv = ops.roi_extract(feats, rois, [1 / s for s in synth.STRIDES], return_levels=True)
    assert out[1].abs().max().item() == 0 and lv[1].item() == -1 and out[0].abs().max().item() > 0


@pytest.mark.parametrize('channels_last', [False, True])
def test_roi_backward(cuda, channels_last):
    tf, feats, rois, scales, out = _case(cuda, 2, (256, 320), 64, 120, seed=4, clustered=True,
                                         channels_last=channels_last)
    g = np.random.RandomState(5).normal(0, 1, out.shape).astype(np.float32)
    out.backward(torch.from_numpy(g).to(cuda))
    ref = oracle.roi_extract_backward(g, rois, [f.shape for f in feats], scales)
    for l, (t, r) in enumerate(zip(tf, ref)):
        assert t.grad is not None and t.grad.shape == r.shape  # every level gets a grad
        _close(t.grad.cpu().numpy(), r, f'grad level {l}')


def test_roi_backward_full_size_train_cfg(cuda):
    # configs[2] geometry: 2 images, 512 RoIs each, 256 channels, clustered
    tf, feats, rois, scales, out = _case(cuda, 2, (800, 1344), 256, 512, seed=6, clustered=True)
    g = np.random.RandomState(7).normal(0, 1, out.shape).astype(np.float32)
    out.backward(torch.from_numpy(g).to(cuda))
    ref = oracle.roi_extract_backward(g, rois, [f.shape for f in feats], scales)
    for l, (t, r) in enumerate(zip(tf, ref)):
        _close(t.grad.cpu().numpy(), r, f'grad level {l}')


def test_roi_backward_deterministic(cuda):
    res = []
    for _ in range(2):
        tf, feats, rois, scales, out = _case(cuda, 1, (256, 320), 64, 400, seed=8, clustered=True)
        g = np.random.RandomState(9).normal(0, 1, out.shape).astype(np.float32)
        out.backward(torch.from_numpy(g).to(cuda))
        res.append([t.grad.cpu().numpy() for t in tf])
    for a, b in zip(*res):
        np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32))


def test_roi_align_module_single_level(cuda):
    feat = synth.fpn_feats(2, 32, [(50, 84)], seed=10)[0]
    rois = synth.random_rois(2, 30, 800, 1344, seed=11)
    layer = ops.RoIAlign(7, spatial_scale=1 / 16, sampling_ratio=0)
    assert layer.output_size == (7, 7)
    out = layer(torch.from_numpy(feat).to(cuda), torch.from_numpy(rois).to(cuda))
    ref = oracle.roi_align_forward(feat, rois, 7, 1 / 16)
    _close(out.cpu().numpy(), ref, 'single-level roi_align')
    out2 = ops.roi_align(torch.from_numpy(feat).to(cuda), torch.from_numpy(rois).to(cuda), 7, 1 / 16,
                         sampling_ratio=2)
    _close(out2.cpu().numpy(), oracle.roi_align_forward(feat, rois, 7, 1 / 16, sampling_ratio=2),
           'sampling_ratio=2')


def test_layout_helpers_roundtrip(cuda):
    x = torch.randn(3, 20, 13, 21, device=cuda)
    y = ops.to_nhwc(x)
    assert y.shape == (3, 13, 21, 20) and y.is_contiguous()
    assert torch.equal(y, x.permute(0, 2, 3, 1).contiguous())
    assert torch.equal(ops.nhwc_to_nchw(y), x)


@pytest.mark.parametrize('B,C,sizes', [
    (2, 256, [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]),   # UTDAC pyramid
    (3, 20, [(13, 21), (5, 3), (1, 1)]),                              # odd hw, C % 64 != 0
    (1, 68, [(64, 64), (9, 130)]),
])
def test_pyramid_transposes_one_launch(cuda, B, C, sizes):
    from boosting_rcnn_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device='cpu').manual_seed(B * 1000 + C)
    feats = [torch.randn(B, C, h, w, generator=g).to(cuda) for h, w in sizes]
    l0 = lib.brcnn_launch_count()
    nhwc = ops.pyramid_to_nhwc(feats)
    assert lib.brcnn_launch_count() - l0 == 1
    for f, y in zip(feats, nhwc):
        assert y.is_contiguous() and torch.equal(y, f.permute(0, 2, 3, 1).contiguous())
    back = ops.pyramid_to_nchw(nhwc)
    for f, y in zip(feats, back):
        assert y.is_contiguous() and torch.equal(y, f)
    # channels_last inputs are free views (no launch)
    cl = [f.contiguous(memory_format=torch.channels_last) for f in feats]
    l0 = lib.brcnn_launch_count()
    views = ops.pyramid_to_nhwc(cl)
    assert lib.brcnn_launch_count() == l0
    assert all(v.data_ptr() == f.data_ptr() for v, f in zip(views, cl))


def test_bbox2roi_padded_matches_reference_bbox2roi(cuda):
    from boosting_rcnn_b200.roi_head import padded_rois
    from boosting_rcnn_b200.rpn_head import PaddedProposals
    g = torch.Generator().manual_seed(5)
    boxes = torch.rand(4, 37, 5, generator=g).to(cuda) * 300
    num = torch.tensor([37, 0, 5, 36], dtype=torch.int32, device=cuda)
    rois, prior = ops.bbox2roi_padded(boxes, num)
    assert torch.equal(rois, padded_rois(PaddedProposals(boxes, num)))
    assert torch.equal(prior, boxes[..., 4].reshape(-1))


def test_roi_forward_wide_footprints_multi_pass(cuda):
    # elongated RoIs: footprints wider than one ring slot -> several x-chunk
    # passes of the TMA row-streaming kernel; also whole-image RoIs
    ex = [[0, 2, 100, 1330, 160], [1, 0, 0, 1333, 800], [0, 10, 300, 1300, 330],
          [1, 600, 2, 660, 798], [0, 0, 0, 1344, 60], [1, 3, 3, 900, 120]]
    _case(cuda, 2, (800, 1344), 256, 20, seed=21, extra_rois=np.array(ex, dtype=np.float32))


@pytest.mark.parametrize('C', [4, 64, 320, 516])
def test_roi_forward_channel_slabs(cuda, C):
    # C < 256 (partial slab), C > 256 (several slabs, strided per-pixel TMA copies)
    _case(cuda, 2, (256, 320), C, 30, seed=30 + C)


def test_roi_forward_pooled_sizes(cuda):
    # non-7x7 outputs: 5x3 on the TMA kernel, 14x14 on the register-tile kernel
    feat = synth.fpn_feats(2, 32, [(50, 84)], seed=4)[0]
    rois = synth.random_rois(2, 25, 800, 1344, seed=12)
    tf, tr = torch.from_numpy(feat).to(cuda), torch.from_numpy(rois).to(cuda)
    for osz in [(5, 3), (14, 14), (1, 1)]:
        out = ops.roi_align(tf, tr, osz, 1 / 16)
        _close(out.cpu().numpy(), oracle.roi_align_forward(feat, rois, osz, 1 / 16), f'out {osz}')


def test_roi_forward_v1_and_v2_kernels_agree(cuda):
    """BRCNN_ROI_FWD=v1 (register-tile kernel) vs the default TMA kernel: run in
    a subprocess because the switch is read once per process."""
    import os
    import subprocess
    import sys
    code = '''
import sys, numpy as np, torch
sys.path[:0] = [%r, %r]
import synth
from boosting_rcnn_b200 import ops
sizes = synth.featmap_sizes(800, 1344)
feats = [torch.from_numpy(f).cuda() for f in synth.fpn_feats(1, 256, sizes, seed=7)]
rois = torch.from_numpy(synth.random_rois(1, 300, 800, 1344, seed=8)).cuda()
out = ops.roi_extract(feats, rois, [1.0 / s for s in synth.STRIDES], 7)
np.save(sys.argv[1], out.cpu().numpy())
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
       os.path.dirname(os.path.abspath(__file__)))
    import tempfile
    outs = []
    with tempfile.TemporaryDirectory() as d:
        for mode in ('v1', 'v2'):
            path = os.path.join(d, mode + '.npy')
            env = dict(os.environ, BRCNN_ROI_FWD=mode)
            subprocess.run([sys.executable, '-c', code, path], check=True, env=env)
            outs.append(np.load(path))
    _close(outs[0], outs[1], 'v1 vs v2')


@pytest.mark.parametrize('C,n', [(4, 40), (132, 60), (320, 1500)])
def test_roi_backward_channel_slabs_and_long_lists(cuda, C, n):
    # partial / multiple 128-channel slabs; > 1024 RoIs on one tile (list rounds);
    # padding rows, degenerate and whole-image RoIs in the backward pass
    ex = np.array([[-1, 0, 0, 0, 0], [0, 5, 5, 5, 5], [0, 0, 0, 320, 256], [0, 2, 2, 318, 6],
                   [0, 300, 250, 330, 270]], dtype=np.float32)
    tf, feats, rois, scales, out = _case(cuda, 1, (256, 320), C, n, seed=40 + C, clustered=True,
                                         extra_rois=ex)
    g = np.random.RandomState(41).normal(0, 1, out.shape).astype(np.float32)
    out.backward(torch.from_numpy(g).to(cuda))
    live = rois[:, 0] >= 0
    ref = oracle.roi_extract_backward(g[live], rois[live], [f.shape for f in feats], scales)
    for l, (t, r) in enumerate(zip(tf, ref)):
        _close(t.grad.cpu().numpy(), r, f'grad level {l}')


# ---------------------------------------------------------------------------
# (R,7,7,C) feature hand-off: persistent forward kernel (roi_align_fwd3.cuh) and the
# backward gather reading bin-major gradients (roi_align_bwd5.cuh)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize('pad_hw,C,n,clustered', [
    ((256, 320), 64, 50, False), ((800, 1344), 256, 256, True), ((256, 320), 4, 30, False),
    ((256, 320), 320, 30, False), ((256, 320), 516, 30, True), ((800, 1344), 256, 2000, False)])
def test_roi_forward_hwc_layout_matches_nchw_layout(cuda, pad_hw, C, n, clustered):
    """The persistent kernel does the same per-row arithmetic as the per-RoI TMA kernel; only
    the x-chunking of wide footprints differs (fixed 16-pixel slots), so the two layouts agree
    to the last few ulp (and both match the oracle)."""
    ex = np.array([[-1, 0, 0, 0, 0], [0, 5, 5, 5, 5], [1, 0, 0, pad_hw[1], pad_hw[0]],
                   [0, 2, 2, pad_hw[1] - 2, 6], [1, pad_hw[1] - 20, pad_hw[0] - 6, pad_hw[1] + 10,
                                                  pad_hw[0] + 14], [7, 1, 1, 50, 50]],
                  dtype=np.float32)
    sizes = synth.featmap_sizes(*pad_hw)
    scales = [1.0 / s for s in synth.STRIDES]
    feats = [torch.from_numpy(f).to(cuda) for f in synth.fpn_feats(2, C, sizes, seed=50 + C)]
    rois = np.concatenate([synth.random_rois(2, n, pad_hw[0], pad_hw[1], seed=51 + C,
                                             clustered=clustered), ex], 0).astype(np.float32)
    tr = torch.from_numpy(rois).to(cuda)
    a, la = ops.roi_extract(feats, tr, scales, 7, return_levels=True)
    b, lb = ops.roi_extract(feats, tr, scales, 7, return_levels=True, channels_last_out=True)
    assert b.permute(0, 2, 3, 1).is_contiguous() and a.is_contiguous()
    assert torch.equal(la, lb)
    an, bn = a.cpu().numpy(), b.cpu().numpy()
    assert float(np.abs(an - bn).max()) <= 2e-6 * max(float(np.abs(an).max()), 1e-6)
    live = (rois[:, 0] >= 0) & (rois[:, 0] < 2)
    ref, _ = oracle.roi_extract_forward([f.cpu().numpy() for f in feats], rois[live], scales)
    _close(b.cpu().numpy()[live], ref, 'hwc roi features')
    assert float(b[~torch.from_numpy(live).to(cuda)].abs().max()) == 0.0


def test_roi_forward_hwc_other_pooled_sizes(cuda):
    feat = synth.fpn_feats(2, 32, [(50, 84)], seed=4)[0]
    rois = synth.random_rois(2, 25, 800, 1344, seed=12)
    tf, tr = torch.from_numpy(feat).to(cuda), torch.from_numpy(rois).to(cuda)
    for osz in [(5, 3), (1, 1), (7, 2)]:
        out = ops.roi_extract([tf], tr, [1 / 16], osz, channels_last_out=True)
        _close(out.cpu().numpy(), oracle.roi_align_forward(feat, rois, osz, 1 / 16), f'out {osz}')


@pytest.mark.parametrize('grad_hwc', [False, True])
@pytest.mark.parametrize('pad_hw,C,n', [((256, 320), 64, 120), ((256, 320), 132, 60),
                                        ((256, 320), 4, 40), ((800, 1344), 256, 512),
                                        ((256, 320), 320, 1500)])
def test_roi_backward_both_gradient_layouts(cuda, pad_hw, C, n, grad_hwc):
    """grad_out arriving (R,C,7,7)-contiguous (transposed inside the library) or bin-major
    (R,7,7,C) (the permuted-FC hand-off; consumed as is): same gradients, <= 1e-5 vs oracle.
    The 1500-RoI case puts > 512 RoIs on one (image, level) bucket (windowed list rounds)."""
    ex = np.array([[-1, 0, 0, 0, 0], [0, 5, 5, 5, 5], [0, 0, 0, pad_hw[1], pad_hw[0]],
                   [0, 2, 2, pad_hw[1] - 2, 6]], dtype=np.float32)
    B = 1 if n > 1000 else 2
    tf, feats, rois, scales, out = _case(cuda, B, pad_hw, C, n, seed=60 + C, clustered=True,
                                         extra_rois=ex, channels_last_out=grad_hwc)
    g = np.random.RandomState(61).normal(0, 1, out.shape).astype(np.float32)
    gt = torch.from_numpy(g).to(cuda)
    if grad_hwc:
        gt = gt.contiguous(memory_format=torch.channels_last)
    out.backward(gt)
    live = rois[:, 0] >= 0
    ref = oracle.roi_extract_backward(g[live], rois[live], [f.shape for f in feats], scales)
    for l, (t, r) in enumerate(zip(tf, ref)):
        assert t.grad is not None and t.grad.shape == r.shape
        _close(t.grad.cpu().numpy(), r, f'grad level {l}')


def test_roi_backward_hwc_deterministic(cuda):
    res = []
    for _ in range(3):
        tf, feats, rois, scales, out = _case(cuda, 2, (256, 320), 128, 700, seed=70,
                                             clustered=True, channels_last_out=True)
        g = np.random.RandomState(71).normal(0, 1, out.shape).astype(np.float32)
        out.backward(torch.from_numpy(g).to(cuda).contiguous(memory_format=torch.channels_last))
        res.append([t.grad.cpu().numpy() for t in tf])
    for other in res[1:]:
        for a, b in zip(res[0], other):
            np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32))


def test_head_consumes_hwc_features_as_a_view(cuda):
    """ProbConvFCBBoxHead: channels-last RoI features feed fc1 without a copy, give the
    same logits as the reference-ordered flatten, and the gradient returns bin-major."""
    from boosting_rcnn_b200 import configs
    torch.manual_seed(0)
    _, roi_head, _ = configs.build_hot_path('utdac')
    head = roi_head.bbox_head.to(cuda)
    x = torch.randn(64, 256, 7, 7, device=cuda)
    xcl = x.contiguous(memory_format=torch.channels_last).requires_grad_(True)
    assert head._flatten(xcl).data_ptr() == xcl.data_ptr()
    cs0, bp0 = head(x)
    cs1, bp1 = head(xcl)
    assert torch.allclose(cs0, cs1, rtol=1e-5, atol=1e-6) and torch.allclose(bp0, bp1, rtol=1e-5, atol=1e-6)
    w_ref = head.state_dict()['shared_fcs.0.weight']
    y_ref = torch.nn.functional.linear(x.flatten(1), w_ref, head.shared_fcs[0].bias)
    y = torch.nn.functional.linear(head._flatten(xcl), head.shared_fcs[0].weight,
                                   head.shared_fcs[0].bias)
    assert torch.allclose(y, y_ref, rtol=1e-4, atol=1e-4)
    (cs1.sum() + bp1.sum()).backward()
    assert xcl.grad.permute(0, 2, 3, 1).is_contiguous()


def test_roi_forward_schedule_is_invisible(cuda):
    """brcnn_roi_extract_forward pulls RoIs from an atomic counter when the caller passes the
    scheduling scratch and strides statically without it: same bits either way, the scratch is
    left zero-filled (one buffer serves every later call on the stream), and a scratch that
    is too small is refused."""
    from boosting_rcnn_b200 import _lib
    lib = _lib.load()
    B, C, pad_hw = 2, 256, (256, 320)
    sizes = synth.featmap_sizes(*pad_hw)
    scales = [1.0 / s for s in synth.STRIDES]
    feats = [torch.from_numpy(f).to(cuda).permute(0, 2, 3, 1).contiguous()
             for f in synth.fpn_feats(B, C, sizes, seed=5)]
    rois = torch.from_numpy(synth.random_rois(B, 700, pad_hw[0], pad_hw[1], seed=6)).to(cuda)
    R = rois.size(0)
    p = ops.make_roi_params(B, C, sizes, scales, 7, 0, True, 56, out_layout=1)
    ptrs = ops.ptr_array([f.data_ptr() for f in feats])
    stream = torch.cuda.current_stream().cuda_stream
    nbytes = lib.brcnn_roi_extract_forward_workspace_bytes(p)
    assert 0 < nbytes <= 4096
    outs = []
    ws = torch.zeros(nbytes // 4, dtype=torch.int32, device=cuda)
    for scratch in (None, ws, ws):           # the second dynamic call reuses the buffer as left
        out = torch.full((R, 7, 7, C), float('nan'), device=cuda)
        lv = torch.empty(R, dtype=torch.int32, device=cuda)
        rc = lib.brcnn_roi_extract_forward(p, ptrs, rois.data_ptr(), R, out.data_ptr(),
                                           lv.data_ptr(),
                                           scratch.data_ptr() if scratch is not None else None,
                                           nbytes if scratch is not None else 0, stream)
        assert rc == 0
        torch.cuda.synchronize()
        assert int(ws.abs().sum()) == 0, 'scheduling scratch must be left zero-filled'
        outs.append((out, lv))
    for out, lv in outs[1:]:
        assert torch.equal(out.view(torch.int32), outs[0][0].view(torch.int32))
        assert torch.equal(lv, outs[0][1])
    out = torch.empty((R, 7, 7, C), device=cuda)
    rc = lib.brcnn_roi_extract_forward(p, ptrs, rois.data_ptr(), R, out.data_ptr(), None,
                                       ws.data_ptr(), 16, stream)
    assert rc == -2          # BRCNN_ERR_WORKSPACE
