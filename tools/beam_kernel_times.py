#!/usr/bin/env python
"""Per-kernel device times of one config-4 solve (n = 1e5, double-double; 1 and 64 right-hand sides) from the
torch profiler: automatic two-level partitioning and a one-level partitioning beside it."""
import sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from beamfe2_b200 import sweep
from beamfe2_b200.beam import LargeBeam, Q
from torch.profiler import profile, ProfilerActivity
pk = sweep.beam_pack(100000, device='cuda:0')
for parts in (0, 1024):
    beam = LargeBeam(pk['coords'][0], pk['props'][0], [0, 1, 2], n_partitions=parts)
    X = torch.randn((beam.n, Q), dtype=torch.float64, device='cuda:0'); X[:3] = 0
    x1 = X[:, 0].contiguous()
    for name, rhs in (('1 rhs', x1), ('64 rhs', X)):
        beam.solve(rhs); torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                beam.solve(rhs)
            torch.cuda.synchronize()
        print('P', beam.n_partitions, name)
        for e in prof.key_averages():
            if e.device_time_total > 0:
                print('   %-60s %8.1f us x %d' % (e.key[:60], e.device_time_total / e.count, e.count))
