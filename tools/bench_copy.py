#!/usr/bin/env python
"""Bare pinned host<->device copy benchmark: the ceiling of the end-to-end path.

Every rank (one per GPU, torchrun) moves the same bytes per step as bench.py's end-to-end arm --
`--in-bytes` / `--out-bytes` per structure x `--structures` per step, chunked over 3 streams exactly like
beamfe2_b200.batch.HostPipeline -- but launches NO kernel.  Timed with CUDA events between barriers, max over
ranks; rank 0 prints one JSON line with the aggregate GB/s and the structures/s that bandwidth would allow.

    python tools/bench_copy.py                      # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_copy.py
"""
import argparse
import json
import os

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--structures', type=int, default=100000)
    ap.add_argument('--in-bytes', type=int, default=5704)        # dense P20 inputs
    ap.add_argument('--out-bytes', type=int, default=11408)      # dense P20 outputs
    ap.add_argument('--chunk', type=int, default=3125)
    ap.add_argument('--copies-in', type=int, default=1, help='cudaMemcpyAsync calls per chunk, host->device')
    ap.add_argument('--copies-out', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    a = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    dev = torch.device('cuda', local)
    B = a.structures
    wi, wo = a.in_bytes // 8, a.out_bytes // 8

    def split(w, parts):
        edges = [w * k // parts for k in range(parts + 1)]
        return [edges[k + 1] - edges[k] for k in range(parts) if edges[k + 1] > edges[k]]

    # `copies` separate contiguous tensors per direction (the round-1 pipeline moved 5 + 7 tensors per chunk)
    h_in = [torch.ones((B, w), dtype=torch.float64).pin_memory() for w in split(wi, a.copies_in)]
    h_out = [torch.empty((B, w), dtype=torch.float64).pin_memory() for w in split(wo, a.copies_out)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(3)]
    d_in = [[torch.empty((a.chunk, t.shape[1]), dtype=torch.float64, device=dev) for t in h_in] for _ in streams]
    d_out = [[torch.zeros((a.chunk, t.shape[1]), dtype=torch.float64, device=dev) for t in h_out] for _ in streams]

    def step():
        cur = torch.cuda.current_stream(dev)
        for s in streams:
            s.wait_stream(cur)
        for ci, lo in enumerate(range(0, B, a.chunk)):
            hi = min(B, lo + a.chunk)
            m = hi - lo
            k = ci % len(streams)
            with torch.cuda.stream(streams[k]):
                for h, d in zip(h_in, d_in[k]):
                    d[:m].copy_(h[lo:hi], non_blocking=True)
                for h, d in zip(h_out, d_out[k]):
                    h[lo:hi].copy_(d[:m], non_blocking=True)
        for s in streams:
            cur.wait_stream(s)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item()) / a.steps
    if rank == 0:
        total = world * B * (a.in_bytes + a.out_bytes)
        print(json.dumps({'bench': 'pinned host<->device copies, no kernel', 'n_gpus': world, 'structures_per_gpu': B,
                          'in_bytes_per_structure': a.in_bytes, 'out_bytes_per_structure': a.out_bytes,
                          'copies_per_chunk': [a.copies_in, a.copies_out], 'chunk': a.chunk,
                          'ms_per_step': ms, 'aggregate_GBps': total / ms / 1e6,
                          'per_gpu_GBps': total / ms / 1e6 / world,
                          'structures_per_s_ceiling': world * B / ms * 1e3,
                          'host_cpus': os.cpu_count()}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
