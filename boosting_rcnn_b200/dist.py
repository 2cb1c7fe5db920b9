"""Multi-GPU plumbing of the hot path (torch.distributed only carries bytes).

The path shards by image with NO data-path collective:
``shard_range`` gives each rank its images.  The only collectives are the
ones the reference issues around the path in training:
  * the gradient all-reduce (left to torch DDP / NCCL, mmdet/apis/train.py:75-83);
  * scalar reductions — ``reduce_mean`` x2 in the RPN loss
    (atss_rpn_head.py:441-444,458-459) and one all-reduce per logged key in
    ``_parse_losses`` (mmdet/models/detectors/base.py:201-207), each followed
    by ``.item()``.  ``fused_scalar_allreduce`` replaces those 2 + 7 tiny
    all-reduces by ONE all-reduce of a packed <=16-float vector.
"""
from collections import OrderedDict

import torch
import torch.distributed as dist


def get_dist_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(num_images, rank=None, world_size=None):
    """Contiguous image shard [lo, hi) of this rank; remainders go to the
    first ranks (every image is processed exactly once)."""
    if rank is None or world_size is None:
        rank, world_size = get_dist_info()
    base, rem = divmod(num_images, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_mean(tensor):
    """mmdet/core/utils/dist_utils.py:67-73: all-reduce of tensor / world."""
    _, world = get_dist_info()
    if world == 1:
        return tensor
    tensor = tensor.clone()
    dist.all_reduce(tensor.div_(world), op=dist.ReduceOp.SUM)
    return tensor


def fused_scalar_allreduce(named_scalars):
    """Average every scalar of an ordered {name: 0-dim tensor} dict over the
    ranks with a single all-reduce; returns an OrderedDict of 0-dim tensors.
    Same values as calling ``reduce_mean`` once per entry (base.py:201-207),
    one collective instead of len(named_scalars)."""
    named_scalars = OrderedDict(named_scalars)
    _, world = get_dist_info()
    if world == 1 or not named_scalars:
        return named_scalars
    vals = list(named_scalars.values())
    packed = torch.stack([v.detach().reshape(()).float() for v in vals])
    dist.all_reduce(packed.div_(world), op=dist.ReduceOp.SUM)
    return OrderedDict((k, packed[i]) for i, k in enumerate(named_scalars))


def collect_detections(det_bboxes, det_labels, num_dets, size=None, interleaved=False):
    """Result-side wire of the multi-GPU test loop (mmdet/apis/test.py:273-340
    ``collect_results_cpu/gpu``): gather every rank's detections on rank 0.

    The reference pickles per-image python lists, all-gathers their byte lengths, pads and
    all-gathers the bytes (2 collectives + pickle on every rank).  The hot path already holds
    fixed-capacity tensors — det_bboxes (b,M,5), det_labels (b,M), num_dets (b,) as returned by
    ``simple_test_bboxes_padded`` — so ONE ``all_gather`` of a packed (b_max, M, 6) tensor +
    one of the counts is enough.  Ranks may own different numbers of images (``shard_range``).

    Returns on rank 0 a list with one (k,6) float32 tensor [x1,y1,x2,y2,score,label] per image,
    in dataset order: concatenation of the contiguous shards, or the reference's round-robin
    order (``interleaved=True``, DistributedSampler + ``zip(*part_list)``); truncated to
    ``size`` (the sampler may pad).  Other ranks get None (like the reference)."""
    rank, world = get_dist_info()
    packed = torch.cat([det_bboxes, det_labels.to(det_bboxes.dtype).unsqueeze(-1)], dim=-1)
    counts = num_dets.to(torch.int64)
    if world > 1:
        nb = torch.tensor([packed.size(0)], dtype=torch.int64, device=packed.device)
        nbs = [torch.zeros_like(nb) for _ in range(world)]
        dist.all_gather(nbs, nb)
        bmax = int(max(int(x) for x in nbs))
        pad = packed.new_zeros((bmax,) + tuple(packed.shape[1:]))
        pad[:packed.size(0)] = packed
        cpad = counts.new_zeros((bmax,))
        cpad[:counts.numel()] = counts
        parts = [torch.zeros_like(pad) for _ in range(world)]
        cparts = [torch.zeros_like(cpad) for _ in range(world)]
        dist.all_gather(parts, pad)
        dist.all_gather(cparts, cpad)
        if rank != 0:
            return None
        per_rank = [[p[i, :int(c[i])].cpu() for i in range(int(n))]
                    for p, c, n in zip(parts, cparts, nbs)]
    else:
        per_rank = [[packed[i, :int(counts[i])].cpu() for i in range(packed.size(0))]]
    if interleaved:
        out, i = [], 0
        while any(i < len(r) for r in per_rank):
            out.extend(r[i] for r in per_rank if i < len(r))
            i += 1
    else:
        out = [x for r in per_rank for x in r]
    return out if size is None else out[:size]


def max_over_ranks(value, device=None):
    """Timing rule of bench.py: device time of a step is the max over ranks."""
    _, world = get_dist_info()
    if world == 1:
        return float(value)
    t = torch.tensor([float(value)], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def bind_to_gpu_cpus(gpu_index):
    """Pin this process to the CPU cores NVML reports as local to ``gpu_index`` (same
    NUMA node / PCIe root) BEFORE pinned host buffers are allocated, so that first-touch
    places them next to the GPU and the H2D copies of 8 concurrent ranks do not cross
    sockets.  Returns the CPU set, or None when NVML / affinity are unavailable."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(gpu_index))
        n_words = (os.cpu_count() + 63) // 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return sorted(cpus)
    except Exception:  # noqa: BLE001 - best effort, never fatal
        return None
    return None
