"""Config 3 in miniature: a parametric sweep of 20-element portal frames through the batched front end,
then the post-processing kernels (moment envelopes, modal participation)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beamfe2_b200 import batch, sweep                    # noqa: E402
from beamfe2_b200.plan import Plan                        # noqa: E402

B = 10000
conn, supports = sweep.p20_topology()
plan = Plan.from_supports(21, conn, supports)
params = sweep.p20_draw_params(B)
pk = sweep.p20_pack(torch.from_numpy(params).cuda())
out = batch.solve_fused(plan, pk['coords'], pk['props'], pk['nodal_mass'], pk['f_nodal'], pk['r_local'], n_modes=10)
assert bool((out['info'] == 0).all())
print('first frequency [Hz]: min %.3f  max %.3f' % tuple((out['lam'][:, 0].sqrt() / 6.283185307179586).aminmax()))
print('max |u| over the sweep: %.3f mm' % out['u'].abs().max().item())
s = torch.from_numpy(params[:, 9]).cuda()
qp = torch.zeros((B, 3, 20), dtype=torch.float64, device='cuda')
qp[:, 1:, 7:13] = (-s * 10.0)[:, None, None]
_, env = batch.internal_actions(plan, pk['coords'], out['endforces'], None, qp, npts=21, want_field=False)
print('largest girder moment of load case 3: %.4g Nmm' % env[:, 2, 7:13, 2].max().item())
full = Plan(21, conn, [])
_, Mfull = batch.assemble_banded(full, pk['coords'], pk['props'], pk['nodal_mass'])
gam = batch.modal_participation(full, Mfull, out['phi'])
print('mean share of the sway mass in mode 1: %.3f' % (gam[:, 0, 0] ** 2 / (gam[:, :, 0] ** 2).sum(dim=1)).mean().item())
