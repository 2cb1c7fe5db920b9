#!/usr/bin/env python
"""All-reduce time of the training step's gradient bucket (57 MB fp32) at this world size:
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/nccl_probe.py
Prints the device time per all-reduce (CUDA events, max over ranks) and the bus bandwidth;
run with NCCL_DEBUG=INFO to see the transport (P2P/NVL, NVLS, SHM)."""
import os
import torch
import torch.distributed as dist

rank, local = int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
dev = torch.device('cuda', local)
torch.cuda.set_device(dev)
dist.init_process_group('nccl', device_id=dev)
world = dist.get_world_size()
for nbytes in (57226820, 4 << 20, 8):
    t = torch.ones(nbytes // 4, device=dev)
    for _ in range(5):
        dist.all_reduce(t)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        dist.all_reduce(t)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 20], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        m = float(ms.item())
        print(f'all_reduce {nbytes} B x{world}: {m:.4f} ms  busbw {nbytes * 2 * (world - 1) / world / (m * 1e-3) / 1e9:.1f} GB/s',
              flush=True)
dist.destroy_process_group()
