#!/usr/bin/env python
"""Static-only fused path on the P20 sweep: tile kernel (one lane per structure) against the group kernel behind the
unfused K2 -> K3 entry, and its time per batch.  usage: static_tile_probe.py [B [C]]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beamfe2_b200 import _lib, batch, sweep       # noqa: E402
from beamfe2_b200.plan import Plan                # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
conn, sup = sweep.p20_topology()
plan = Plan.from_supports(21, conn, sup)
pk = sweep.p20_pack(torch.from_numpy(sweep.p20_draw_params(B)).cuda())
C = int(sys.argv[2]) if len(sys.argv) > 2 else 3              # load cases kept (P20 has three)
pk['f_nodal'] = pk['f_nodal'][:, :C].contiguous()
pk['r_local'] = pk['r_local'][:, :C].contiguous()
L = _lib.lib()
n0 = L.bfe2_static_tile_launches()
so = batch.solve_fused(plan, pk['coords'], pk['props'], None, pk['f_nodal'], pk['r_local'], modal=False)
torch.cuda.synchronize()
print('tile launches:', L.bfe2_static_tile_launches() - n0, ' info != 0:', int((so['info'] != 0).sum()))
Kb, _ = batch.assemble_banded(plan, pk['coords'], pk['props'], None)
ref = batch.static_solve(plan, Kb, pk['coords'], pk['props'], pk['f_nodal'], pk['r_local'])
for key in ('u', 'endforces', 'reactions'):
    a, b = so[key], ref[key]
    scale = b.abs().amax(dim=tuple(range(1, b.dim())), keepdim=True).clamp_min(1e-300)
    print('%-10s max |tile - group| / max|group| per structure: %.3e   nan: %d' % (key, float(((a - b).abs() / scale).max()), int(torch.isnan(a).sum())))
for name, fn in (('fused static-only', lambda: batch.solve_fused(plan, pk['coords'], pk['props'], None, pk['f_nodal'], pk['r_local'], modal=False, out=so)),
                 ('unfused K3', lambda: batch.static_solve(plan, Kb, pk['coords'], pk['props'], pk['f_nodal'], pk['r_local'], out=ref) if False else batch.static_solve(plan, Kb, pk['coords'], pk['props'], pk['f_nodal'], pk['r_local']))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print('%-18s %.3f ms per %d structures = %.1f M structures/s' % (name, ms, B, B / ms / 1e3))
