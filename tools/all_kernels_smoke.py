#!/usr/bin/env python
"""Tiny end-to-end invocation of EVERY kernel of libbfe2 (was meant for compute-sanitizer, which is closed on this GPU
pool: used as a plain smoke run)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from beamfe2_b200 import batch, sweep              # noqa: E402
from beamfe2_b200.plan import Plan                  # noqa: E402
from beamfe2_b200.beam import LargeBeam             # noqa: E402

dev = 'cuda:0'
B = 8
P = sweep.p20_draw_params(B)
pk = sweep.p20_pack(torch.from_numpy(P).to(dev))
conn, sup = sweep.p20_topology()
plan = Plan.from_supports(21, conn, sup)
out = batch.solve_fused(plan, pk['coords'], pk['props'], pk['nodal_mass'], pk['f_nodal'], pk['r_local'], n_modes=10)
Kb, Mb = batch.assemble_banded(plan, pk['coords'], pk['props'], pk['nodal_mass'])
st = batch.static_solve(plan, Kb, pk['coords'], pk['props'], pk['f_nodal'], pk['r_local'])
mo = batch.modal_solve(plan, Kb, Mb, n_modes=57)
act, mx = batch.internal_actions(plan, pk['coords'], out['endforces'], None, None, npts=7)
full = Plan(21, conn, [])
Kf, Mf = batch.assemble_banded(full, pk['coords'], pk['props'], pk['nodal_mass'])
gam = batch.modal_participation(full, Mf, out['phi'])
c, p = pk['coords'], pk['props']
Kg, Mg = batch.element_km(c[:, :-1, 0].reshape(-1).contiguous(), c[:, :-1, 1].reshape(-1).contiguous(),
                          c[:, 1:, 0].reshape(-1).contiguous(), c[:, 1:, 1].reshape(-1).contiguous(),
                          *(p[:, :, k].reshape(-1).contiguous() for k in range(4)))
# odd / tiny / wide-band topologies: Chopra 10.5 (hbw 8), 2-element cantilever (n = 6), rod with n = 90 (smem Jacobi)
for nE, fixed in ((2, [0, 1, 2]), (30, [0, 1, 2]), (11, [0, 1, 32])):
    cn = np.stack([np.arange(nE), np.arange(1, nE + 1)], 1).astype(np.int32)
    pl = Plan(nE + 1, cn, fixed)
    bp = sweep.beam_pack(nE, length=100.0, device=dev)
    f = bp['f_nodal'].repeat(3, 1, 1)
    o = batch.solve_fused(pl, bp['coords'].repeat(3, 1, 1), bp['props'].repeat(3, 1, 1), None, f, None, n_modes=min(4, pl.n_free))
    assert int(o['info'].abs().max()) == 0
bp = sweep.beam_pack(300, device=dev)
beam = LargeBeam(bp['coords'][0], bp['props'][0], [0, 1, 2], n_partitions=7)
u = beam.static(bp['f_nodal'][0, 0])
m = beam.modal(n_modes=6, maxit=6)
# round 2: packed records, the two-ended engine, concentrated load curves, geometric stiffness / buckling, condensed
# lumped mass, band products, double-double residual
io = batch.PackedIO.from_dense(plan, pk['f_nodal'], pk['r_local'], n_modes=10)
rec = torch.zeros((B, io.w_in), dtype=torch.float64, device=dev)
io.pack_inputs(rec, pk['coords'], pk['props'], pk['nodal_mass'], pk['f_nodal'], pk['r_local'])
pv = io.out_views(io.solve(rec))
prev = batch.set_modal_engine('two_ended')
te = batch.solve_fused(plan, pk['coords'], pk['props'], pk['nodal_mass'], pk['f_nodal'], pk['r_local'], n_modes=10)
tp = io.out_views(io.solve(rec))
batch.set_modal_engine(prev)
assert int(te['info'].abs().max()) == 0 and int(te['sweeps'].max()) == 0 and torch.equal(tp['lam'], te['lam'])
assert torch.equal(pv['lam'], out['lam']) and (te['lam'] / out['lam'] - 1).abs().max().item() < 2e-10
conc = torch.zeros((B, 3, 20, 1, 3), dtype=torch.float64, device=dev)
conc[:, :, 9, 0] = torch.tensor([1.0, -5.0, 0.4], dtype=torch.float64, device=dev)
act2, mx2 = batch.internal_actions(plan, pk['coords'], out['endforces'], None, None, npts=7, conc=conc)
bk = batch.buckling(plan, pk['coords'], pk['props'], pk['f_nodal'][:, 2:3].contiguous(), None, n_modes=2)
assert int(bk['info'].abs().max()) == 0
lp = Plan.from_supports(21, conn, sup, lumped=np.ones(20, dtype=np.uint8))
lm = batch.modal_lumped_condensed(lp, pk['coords'], pk['props'], pk['nodal_mass'], n_modes=4)
assert int(lm['info'].abs().max()) == 0 and bool((lm['lam'] > 0).all())
from beamfe2_b200.frame_modes import FrameModes
fm = FrameModes(plan, pk['coords'][0], pk['props'][0], pk['nodal_mass'][0])
X = torch.randn((plan.n_free, 64), dtype=torch.float64, device=dev)
Kd = torch.as_tensor(plan.band_to_dense(Kb[0].cpu().numpy()), device=dev)
assert ((fm.matvec(X) - Kd @ X).abs().max() / (Kd @ X).abs().max()).item() < 1e-14
u2, ulo = beam.static(bp['f_nodal'][0, 0], with_lo=True)
r = beam.residual(bp['f_nodal'][0, 0], u2, ulo)
torch.cuda.synchronize()
assert int(out['info'].abs().max()) == 0 and int(mo['info'].abs().max()) == 0
print('sanitize_case ok', float(out['lam'][0, 0]), float(u[-2]), float(m['lam'][0]))
