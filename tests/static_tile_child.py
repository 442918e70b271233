"""Child process of tests/test_gpu_static_tile.py: runs with BFE2_STATIC_TILE=1 (the switch is read once per process).
Static-only requests with ONE load case on chains go through the one-lane-per-structure tile kernel; results are
compared with the CPU oracle on a few variants and with the group kernel (behind the unfused K2 -> K3 entry) on all."""
import sys

import numpy as np
import torch

from beamfe2_b200 import _lib, batch
from beamfe2_b200.plan import Plan
from oracle import beamfe2_oracle as orc
from tests import util

DEV = 'cuda:0'
T = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=DEV)     # noqa: E731


def run(n_nodes, B, with_loads, seed, expect_tile=1):
    rng = np.random.default_rng(seed)
    xy = np.stack([np.arange(n_nodes) * 700.0, 300.0 * np.sin(np.arange(n_nodes))], axis=1)
    conn = np.array([(k, k + 1) for k in range(n_nodes - 1)], dtype=np.int32)
    fixed = [0, 1, 2] + ([3 * (n_nodes - 1) + 1] if n_nodes > 4 else [])
    nE = len(conn)
    plan = Plan(n_nodes, conn, fixed)
    assert plan.hbw <= 5
    coords = xy[None] + rng.uniform(-20, 20, (B, n_nodes, 2))
    props = np.stack([rng.uniform(1.9e5, 2.2e5, (B, nE)), rng.uniform(2e3, 2e4, (B, nE)),
                      rng.uniform(1e6, 1e8, (B, nE)), rng.uniform(7e-9, 8.5e-9, (B, nE))], axis=-1)
    f_nodal = rng.uniform(-1e4, 1e4, (B, 1, 3 * n_nodes)) * (rng.uniform(size=(1, 1, 3 * n_nodes)) < 0.3)
    f_nodal[:, :, 0] += 1.0
    r_local = None
    if with_loads:
        L, _ = orc.length_direction(coords, conn)
        q = rng.uniform(-5, 5, (B, 1, nE))
        Lc = np.broadcast_to(L[:, None, :], q.shape)
        r_local = np.zeros((B, 1, nE, 6))
        r_local[..., 1] = q * Lc / 2
        r_local[..., 2] = q * (Lc ** 2) / 12
        r_local[..., 4] = q * Lc / 2
        r_local[..., 5] = -1 * q * (Lc ** 2) / 12.
    bad = [5, B - 1]                                         # not positive definite: negative modulus
    props[bad, :, 0] *= -1.0
    lib = _lib.lib()
    n0 = lib.bfe2_static_tile_launches()
    rl = None if r_local is None else T(r_local)
    tile = batch.solve_fused(plan, T(coords), T(props), None, T(f_nodal), rl, modal=False)
    assert lib.bfe2_static_tile_launches() == n0 + expect_tile, 'tile kernel launches'
    Kb, _ = batch.assemble_banded(plan, T(coords), T(props), None)
    group = batch.static_solve(plan, Kb, T(coords), T(props), T(f_nodal), rl)
    assert lib.bfe2_static_tile_launches() == n0 + expect_tile
    good = np.setdiff1d(np.arange(B), bad)
    info = tile['info'].cpu().numpy()
    assert (info[good] == 0).all() and (info[bad] == 1).all(), info[bad]
    pick = good[[0, 1, len(good) // 2, -2, -1]]
    o = orc.solve_batch(conn, fixed, coords[pick], props[pick], None, f_nodal[pick],
                        np.zeros((len(pick), 1, nE, 6)) if r_local is None else r_local[pick])
    fp = orc.free_positions(n_nodes, fixed)
    tol = max(1e-10, 50 * np.finfo(float).eps * np.linalg.cond(o['K'][:, fp][:, :, fp]).max())   # as in test_gpu_random_frames
    for key in ('u', 'endforces', 'reactions'):
        a, g = tile[key].cpu().numpy(), group[key].cpu().numpy()
        assert np.isnan(a[bad]).all(), key
        assert np.isfinite(a[good]).all(), key
        for k, b in enumerate(pick):
            assert util.relmax(a[b], o[key][k]) < tol, (key, int(b), util.relmax(a[b], o[key][k]), tol)
        scale = np.abs(g[good]).max(axis=tuple(range(1, g.ndim)), keepdims=True)
        assert (np.abs(a[good] - g[good]) / scale).max() < tol, (key, (np.abs(a[good] - g[good]) / scale).max(), tol)
    # optional outputs: displacements do not depend on what else is requested
    u_only = batch.solve_fused(plan, T(coords), T(props), None, T(f_nodal), rl, modal=False, want_endforces=False,
                               want_reactions=False)
    assert lib.bfe2_static_tile_launches() == n0 + 2 * expect_tile
    assert torch.equal(u_only['u'][torch.as_tensor(good, device=DEV)], tile['u'][torch.as_tensor(good, device=DEV)])


if __name__ == '__main__':
    run(21, 1000 + 19, True, 1)          # ragged last tile
    run(21, 4096, False, 2)              # no beam loads
    run(8, 333, True, 3)
    run(3, 40, True, 4)                  # one free node
    run(40, 64, True, 5, expect_tile=0)  # n_free = 116: two tiles do not fit an SM -> the group kernel answers
    print('OK')
    sys.exit(0)
