#!/usr/bin/env python
"""Batched first-k modes of mid-size frames (n_free > 116): BatchedFrameModes on B variants of a multi-storey frame
against a loop of single-structure FrameModes and against scipy.linalg.eigh (dense, all modes) on the host.
    python tools/bench_frames.py [--storeys 3 --bays 4 --batch 256 --modes 12]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from beamfe2_b200.frame_modes import BatchedFrameModes, FrameModes   # noqa: E402
from beamfe2_b200.plan import Plan                                   # noqa: E402
from oracle import beamfe2_oracle as orc                             # noqa: E402


def frame(storeys, bays, H=3000.0, Lb=5000.0):
    nodes, idx = [], {}

    def node(x, y):
        key = (round(x, 6), round(y, 6))
        if key not in idx:
            idx[key] = len(nodes)
            nodes.append((x, y))
        return idx[key]
    conn, kinds = [], []
    for b in range(bays + 1):
        for s_ in range(storeys):
            ys = [s_ * H, s_ * H + H / 2, (s_ + 1) * H]
            for a, c in zip(ys[:-1], ys[1:]):
                conn.append((node(b * Lb, a), node(b * Lb, c)))
                kinds.append(0)
    for s_ in range(1, storeys + 1):
        for b in range(bays):
            xs = [b * Lb, b * Lb + Lb / 2, (b + 1) * Lb]
            for a, c in zip(xs[:-1], xs[1:]):
                conn.append((node(a, s_ * H), node(c, s_ * H)))
                kinds.append(1)
    xy = np.array(nodes)
    fixed = sorted(3 * i + d for i, (x, y) in enumerate(nodes) if y == 0.0 for d in range(3))
    return xy, np.array(conn, dtype=np.int32), np.array(kinds), fixed


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--storeys', type=int, default=3)
    ap.add_argument('--bays', type=int, default=4)
    ap.add_argument('--batch', type=int, default=256)
    ap.add_argument('--modes', type=int, default=12)
    a = ap.parse_args()
    dev = 'cuda:0'
    xy, conn, kinds, fixed = frame(a.storeys, a.bays)
    plan = Plan(len(xy), conn, fixed, reorder=True)
    B = a.batch
    rng = np.random.default_rng(5)
    coords = xy[None] + rng.uniform(-50, 50, (B,) + xy.shape) * (xy[None, :, 1:2] > 0)
    base = np.where(kinds[:, None] == 0, [[2.1e5, 9.0e4, 300.0 ** 4 / 12, 7.85e-9]], [[2.1e5, 1.0e5, 200.0 * 500.0 ** 3 / 12, 7.85e-9]])
    props = base[None] * rng.uniform(0.7, 1.4, (B, len(conn), 1))
    mass = np.zeros((B, len(xy), 2))
    mass[:, xy[:, 1] > 0, :] = rng.uniform(0.5, 4.0, (B, 1, 1))
    T = lambda x: torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64, device=dev)   # noqa: E731
    tc, tp, tm = T(coords), T(props), T(mass)
    BatchedFrameModes(plan, tc[:2], tp[:2], tm[:2]).modes(n_modes=a.modes, maxit=3)           # warm-up: lazy kernel loading
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = BatchedFrameModes(plan, tc, tp, tm).modes(n_modes=a.modes)
    torch.cuda.synchronize()
    t_batched = time.perf_counter() - t0
    ns = min(B, 8)
    t0 = time.perf_counter()
    for b in range(ns):
        FrameModes(plan, tc[b], tp[b], tm[b]).modes(n_modes=a.modes)
    torch.cuda.synchronize()
    t_single = (time.perf_counter() - t0) / ns
    import scipy.linalg as sla
    t0 = time.perf_counter()
    err = 0.0
    for b in range(ns):
        o = orc.solve_batch(conn, fixed, coords[b], props[b], nodal_mass=mass[b])
        w = sla.eigh(o['Kc'], o['Mc'], eigvals_only=True)
        err = max(err, float(np.abs(out['lam'][b].cpu().numpy() / w[:a.modes] - 1).max()))
    t_cpu = (time.perf_counter() - t0) / ns
    print(json.dumps({'n_free': plan.n_free, 'hbw': plan.hbw, 'batch': B, 'modes': a.modes, 'iterations': out['iterations'],
                      'converged': bool(out['converged'].all()), 'max_scaled_residual': float(out['residual'].max()),
                      'max_rel_err_vs_scipy_eigh': err, 'batched_ms_per_structure': 1e3 * t_batched / B,
                      'single_structure_route_ms_per_structure': 1e3 * t_single,
                      'host_assembly_plus_scipy_eigh_all_modes_ms_per_structure': 1e3 * t_cpu}, indent=1))


if __name__ == '__main__':
    main()
