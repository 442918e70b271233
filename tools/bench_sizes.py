"""Fused modal + 10 shapes on chains of growing size: structures/s per n_free, for either modal engine.
    python tools/bench_sizes.py [jacobi|two_ended]"""
import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beamfe2_b200 import batch
from beamfe2_b200.plan import Plan
ENGINE = sys.argv[1] if len(sys.argv) > 1 else 'jacobi'
batch.set_modal_engine(ENGINE)
print('# engine', ENGINE)


def run(nn, B):
    xy = np.stack([np.arange(nn) * 500.0, np.zeros(nn)], 1)
    conn = np.array([(k, k + 1) for k in range(nn - 1)], dtype=np.int32)
    plan = Plan(nn, conn, [0, 1, 2])
    rng = np.random.default_rng(0)
    coords = torch.as_tensor(xy[None] + rng.uniform(-10, 10, (B, nn, 2)), device='cuda')
    props = torch.as_tensor(np.tile(np.array([2.1e5, 30., 22.5, 7.85e-9]), (B, nn - 1, 1)) * rng.uniform(0.9, 1.1, (B, nn - 1, 1)), device='cuda')
    out = batch.solve_fused(plan, coords, props, None, None, None, n_modes=10)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        batch.solve_fused(plan, coords, props, None, None, None, n_modes=10, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print('n_free %3d  B %d  %.2f ms  %.0f structures/s  sweeps %.2f  info %d' % (plan.n_free, B, ms, B / ms * 1e3, out['sweeps'].double().mean().item(), int(out['info'].abs().sum())))
SIZES = ((11, 100000), (20, 100000), (21, 50000), (22, 50000), (23, 20000), (25, 20000), (27, 20000), (31, 20000), (34, 10000), (39, 10000))
if len(sys.argv) > 2:                                   # python tools/bench_sizes.py jacobi 23,25  -> only these node counts
    want = {int(v) for v in sys.argv[2].split(',')}
    SIZES = tuple((nn, B) for nn, B in SIZES if nn in want)
for nn, B in SIZES:
    if ENGINE == 'two_ended' and 3 * (nn - 1) > 64:
        continue
    run(nn, B)
