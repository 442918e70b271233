import sys, torch, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beamfe2_b200 import batch
from beamfe2_b200.plan import Plan
from oracle import beamfe2_oracle as orc
EPS = np.finfo(float).eps
for nn, B in ((23, 20000), (31, 20000), (39, 10000)):
    xy = np.stack([np.arange(nn) * 500.0, 100.0 * np.sin(np.arange(nn))], 1)
    conn = np.array([(k, k + 1) for k in range(nn - 1)], dtype=np.int32)
    fixed = [0, 1, 2]
    plan = Plan(nn, conn, fixed)
    rng = np.random.default_rng(nn)
    coords = xy[None] + rng.uniform(-30, 30, (B, nn, 2))
    props = np.tile(np.array([2.1e5, 30., 22.5, 7.85e-9]), (B, nn - 1, 1)) * rng.uniform(0.7, 1.3, (B, nn - 1, 4))
    mass = rng.uniform(0, 1e-5, (B, nn, 2))
    T = lambda a: torch.as_tensor(np.ascontiguousarray(a), device='cuda')
    out = batch.solve_fused(plan, T(coords), T(props), T(mass), None, None, n_modes=6)
    assert int(out['info'].abs().sum()) == 0
    lam = out['lam'].cpu().numpy()
    assert (np.diff(lam, axis=1) >= 0).all() and (lam > 0).all()
    idx = rng.choice(B, 8, replace=False)
    o = orc.solve_batch(conn, fixed, coords[idx], props[idx], mass[idx], n_modes=6)
    rel = np.abs(lam[idx] / o['lam'] - 1) / (1e-10 + 64 * EPS * o['lam'][:, -1:] / o['lam'])
    print('n_free %d: B %d, max sweeps %d, eigenvalue error / LAPACK bound %.3f' % (plan.n_free, B, int(out['sweeps'].max()), rel.max()))
    assert rel.max() < 1.0
print('mid-size soak ok')
