#!/usr/bin/env python
"""Where the time of the config-4 modal solve goes: CUDA-event timing of the building blocks at n = 1e5, 64 vectors."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from beamfe2_b200 import sweep                     # noqa: E402
from beamfe2_b200.beam import LargeBeam, Q         # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
pk = sweep.beam_pack(n, device='cuda:0')
for parts in (0, 1024, 2048):
    beam = LargeBeam(pk['coords'][0], pk['props'][0], [0, 1, 2], n_partitions=parts)
    X = torch.randn((beam.n, Q), dtype=torch.float64, device='cuda:0')
    X[:3] = 0
    Y = beam.matvec(X, mass=True)
    Z = torch.empty_like(X)
    A = beam.gram(X, Y)
    Bm = A @ A.t() + 64 * torch.eye(64, dtype=torch.float64, device='cuda:0')
    print('partitions %d:  solve 64 rhs %.2f ms | solve 1 rhs %.2f ms | matvec M %.2f ms | gram %.2f ms | update %.2f ms | '
          'small_eig %.2f ms | rayleigh_ritz (host logic + eig) %.2f ms'
          % (beam.n_partitions, timed(lambda: beam.solve(Y, out=Z)), timed(lambda: beam.solve(Y[:, 0].contiguous())),
             timed(lambda: beam.matvec(X, mass=True)), timed(lambda: beam.gram(X, Y)), timed(lambda: beam.update(X, A)),
             timed(lambda: beam.small_eig(Bm, Bm + torch.eye(64, dtype=torch.float64, device='cuda:0'))),
             timed(lambda: beam._rayleigh_ritz(Bm, Bm + torch.eye(64, dtype=torch.float64, device='cuda:0')))))

import time
beam = LargeBeam(pk['coords'][0], pk['props'][0], [0, 1, 2])
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mo = beam.modal(n_modes=50, tol=1e-10, maxit=60)
    torch.cuda.synchronize()
    print('modal() call %d: %.3f s, %d iterations, converged %s' % (rep, time.perf_counter() - t0, mo['iterations'], mo['converged']))
