import sys, torch
sys.path.insert(0, '/root/repo')
from beamfe2_b200 import batch, sweep
from beamfe2_b200.plan import Plan
B = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
conn, sup = sweep.p20_topology()
plan = Plan.from_supports(21, conn, sup)
pk = sweep.p20_pack(torch.from_numpy(sweep.p20_draw_params(B)).cuda())
so = batch.solve_fused(plan, pk['coords'], pk['props'], None, pk['f_nodal'], pk['r_local'], modal=False)
for _ in range(3):
    batch.solve_fused(plan, pk['coords'], pk['props'], None, pk['f_nodal'], pk['r_local'], modal=False, out=so)
torch.cuda.synchronize()
print('ok', int(so['info'].abs().sum()))
