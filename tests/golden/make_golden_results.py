#!/usr/bin/env python
"""Golden vectors for the result-side wire by EXECUTING the reference:

  * ``bbox2result`` (mmdet/core/bbox/transforms.py:100-117) — per-class numpy split;
  * ``collect_results_cpu`` (mmdet/apis/test.py:273-313) — the rank-0 gather of the per-rank
    result lists: ``zip(*part_list)`` interleave + truncation to the dataset size.

Both are lifted out of /root/reference with ``ast`` and run unmodified.  ``collect_results_cpu``
is executed once per simulated rank (rank 1 first, rank 0 last) with a real tmpdir; its imports
are served by stand-ins: ``mmcv.dump / load / mkdir_or_exist`` = pickle / os.makedirs,
``get_dist_info`` = the simulated (rank, world), ``dist.barrier`` = no-op.

    python tests/golden/make_golden_results.py     # rewrites reference_golden_results.npz

World 2, dataset of 5 images: DistributedSampler (round-robin, padded to 6) gives rank 0 the
images 0, 2, 4 and rank 1 the images 1, 3, 0 (the pad repeats image 0; ``[:size]`` drops it).
"""
import os
import os.path as osp
import pickle
import shutil
import sys
import tempfile
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import OUT, lift  # noqa: E402

WORLD, SIZE, CAP, NUM_CLASSES = 2, 5, 8, 4


def shards():
    idx = list(range(SIZE)) + list(range((-SIZE) % WORLD))      # DistributedSampler padding
    return [idx[r::WORLD] for r in range(WORLD)]


def inputs():
    rng = np.random.RandomState(7)
    num = np.array([5, 0, 8, 3, 1], dtype=np.int32)             # image 1 has no detection
    det = np.zeros((SIZE, CAP, 5), dtype=np.float32)
    lab = np.zeros((SIZE, CAP), dtype=np.int64)
    for i in range(SIZE):
        k = num[i]
        xy = rng.uniform(0, 200, (k, 2))
        det[i, :k, :2] = xy
        det[i, :k, 2:4] = xy + rng.uniform(1, 100, (k, 2))
        det[i, :k, 4] = np.sort(rng.uniform(0.05, 1, k))[::-1]
        lab[i, :k] = rng.randint(0, NUM_CLASSES, k)
    return det, lab, num


def main():
    det, lab, num = inputs()
    ns = dict(torch=torch, np=np)
    lift('mmdet/core/bbox/transforms.py', ['bbox2result'], ns)
    parts = [[ns['bbox2result'](torch.from_numpy(det[i, :num[i]]), torch.from_numpy(lab[i, :num[i]]),
                                NUM_CLASSES) for i in sh] for sh in shards()]
    tmp = tempfile.mkdtemp()
    state = {}
    mmcv = types.SimpleNamespace(
        mkdir_or_exist=lambda d: os.makedirs(d, exist_ok=True),
        dump=lambda obj, f: pickle.dump(obj, open(f, 'wb')),
        load=lambda f: pickle.load(open(f, 'rb')))
    cns = dict(torch=torch, mmcv=mmcv, osp=osp, shutil=shutil, tempfile=tempfile, pickle=pickle,
               dist=types.SimpleNamespace(barrier=lambda: None),
               get_dist_info=lambda: (state['rank'], WORLD))
    lift('mmdet/apis/test.py', ['collect_results_cpu'], cns)
    ordered = None
    for rank in reversed(range(WORLD)):
        state['rank'] = rank
        out = cns['collect_results_cpu'](parts[rank], SIZE, tmpdir=osp.join(tmp, 'parts'))
        assert (out is None) == (rank != 0)
        ordered = out if rank == 0 else ordered
    shutil.rmtree(tmp, ignore_errors=True)
    assert len(ordered) == SIZE
    gold = dict(det=det, lab=lab, num=num, world=np.int32(WORLD), num_classes=np.int32(NUM_CLASSES),
                shards=np.array(shards(), dtype=np.int32))
    for i, per_class in enumerate(ordered):
        assert len(per_class) == NUM_CLASSES
        for c, arr in enumerate(per_class):
            gold[f'out/{i}/{c}'] = np.asarray(arr, dtype=np.float32).reshape(-1, 5)
    path = os.path.join(OUT, 'reference_golden_results.npz')
    np.savez_compressed(path, **gold)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
