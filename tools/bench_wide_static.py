import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
from beamfe2_b200 import batch
from beamfe2_b200.plan import Plan
from tests.test_gpu_random_frames import random_frame
for n_nodes, extra in ((12, 3), (20, 5), (30, 6)):
    rng = np.random.default_rng(n_nodes)
    xy, conn, fixed = random_frame(rng, n_nodes, extra, 1, 1)
    plan = Plan(n_nodes, conn, fixed, reorder=True)
    B, C, nE = 20000, 3, len(conn)
    coords = torch.as_tensor(xy[None] + rng.uniform(-20, 20, (B, n_nodes, 2)), device='cuda')
    props = torch.as_tensor(np.stack([rng.uniform(1.9e5, 2.2e5, (B, nE)), rng.uniform(2e3, 2e4, (B, nE)), rng.uniform(1e6, 1e8, (B, nE)), rng.uniform(7e-9, 8.5e-9, (B, nE))], -1), device='cuda')
    f = torch.as_tensor(rng.uniform(-1e4, 1e4, (B, C, 3 * n_nodes)), device='cuda')
    out = batch.solve_fused(plan, coords, props, None, f, None, modal=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        batch.solve_fused(plan, coords, props, None, f, None, modal=False, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print('static: nodes %d n_free %d hbw %d: %.2f ms per %d -> %.0f structures/s (info %d)' % (n_nodes, plan.n_free, plan.hbw, ms, B, B / ms * 1e3, int(out['info'].abs().sum())))
    mass = torch.as_tensor(rng.uniform(0, 1e-3, (B, n_nodes, 2)), device='cuda')
    om = batch.solve_fused(plan, coords, props, mass, f, None, n_modes=8)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        batch.solve_fused(plan, coords, props, mass, f, None, n_modes=8, out=om)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print('static + modal:                          %.2f ms per %d -> %.0f structures/s (sweeps %.2f, info %d)' % (ms, B, B / ms * 1e3, om['sweeps'].double().mean().item(), int(om['info'].abs().sum())))
