import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
from beamfe2_b200 import batch
from beamfe2_b200.plan import Plan
nn, B = 31, 10000
xy = np.stack([np.arange(nn) * 500.0, np.zeros(nn)], 1)
conn = np.array([(k, k + 1) for k in range(nn - 1)], dtype=np.int32)
plan = Plan(nn, conn, [0, 1, 2])
rng = np.random.default_rng(0)
coords = torch.as_tensor(xy[None] + rng.uniform(-10, 10, (B, nn, 2)), device='cuda')
props = torch.as_tensor(np.tile(np.array([2.1e5, 30., 22.5, 7.85e-9]), (B, nn - 1, 1)) * rng.uniform(0.9, 1.1, (B, nn - 1, 1)), device='cuda')
for _ in range(3):
    out = batch.solve_fused(plan, coords, props, None, None, None, n_modes=10)
torch.cuda.synchronize(); print('ok')
