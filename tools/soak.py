#!/usr/bin/env python
"""Soak run: many shards of the P20 sweep through the fused kernel; cheap invariants on every structure and an
oracle comparison (numpy.linalg.solve / scipy.linalg.eigh with LAPACK's error bound) on a random subset."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from beamfe2_b200 import batch, sweep            # noqa: E402
from beamfe2_b200.plan import Plan                # noqa: E402
from oracle import beamfe2_oracle as orc          # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--shards', type=int, default=16)
ap.add_argument('--batch', type=int, default=100000)
ap.add_argument('--check', type=int, default=32)
ap.add_argument('--engine', default='jacobi', choices=['jacobi', 'two_ended'])
args = ap.parse_args()
batch.set_modal_engine(args.engine)
dev = 'cuda:0'
conn, sup = sweep.p20_topology()
plan = Plan.from_supports(21, conn, sup)
fixed = orc.positions_to_eliminate(sup)
free = torch.as_tensor(plan.pos_of_free.astype(np.int64), device=dev)
eps = np.finfo(float).eps
worst = dict(u=0.0, lam=0.0, lam_refined=0.0, sweeps=0, eq=0.0, static_vs_fused=0.0)
io = None
for shard in range(100, 100 + args.shards):
    P = sweep.p20_draw_params(args.batch, shard=shard)
    pk = sweep.p20_pack(torch.from_numpy(P).to(dev))
    out = batch.solve_fused(plan, pk['coords'], pk['props'], pk['nodal_mass'], pk['f_nodal'], pk['r_local'], n_modes=10)
    assert bool((out['info'] == 0).all()), 'info != 0 in shard %d' % shard
    lam = out['lam']
    assert bool((lam[:, 1:] >= lam[:, :-1]).all()) and bool((lam > 0).all())
    react = out['reactions']
    eq = float((react[:, :, free].abs().amax(dim=(1, 2)) / react.abs().amax(dim=(1, 2))).max())
    worst['eq'] = max(worst['eq'], eq)
    worst['sweeps'] = max(worst['sweeps'], int(out['sweeps'].max()))
    # the static-only kernel (16 lanes per structure) against the static part of the fused kernel, every structure
    so = batch.solve_fused(plan, pk['coords'], pk['props'], None, pk['f_nodal'], pk['r_local'], modal=False)
    assert bool((so['info'] == 0).all())
    for key in ('u', 'endforces', 'reactions'):
        dv = float(((so[key] - out[key]).abs().amax(dim=tuple(range(1, out[key].dim()))) /
                    out[key].abs().amax(dim=tuple(range(1, out[key].dim())))).max())
        worst['static_vs_fused'] = max(worst['static_vs_fused'], dv)
    assert worst['static_vs_fused'] < 1e-9, worst
    # the packed-record entry (compact loads / results) against the dense one: bit for bit, every structure
    if io is None:
        io = batch.PackedIO.from_dense(plan, pk['f_nodal'], pk['r_local'], n_modes=10)
    rec = torch.zeros((args.batch, io.w_in), dtype=torch.float64, device=dev)
    io.pack_inputs(rec, pk['coords'], pk['props'], pk['nodal_mass'], pk['f_nodal'], pk['r_local'])
    pv = io.out_views(io.solve(rec))
    assert torch.equal(pv['lam'], out['lam']) and torch.equal(pv['phi'], out['phi'][:, :, free])
    assert torch.equal(pv['u'], out['u'][:, :, free]) and torch.equal(pv['endforces'], out['endforces'])
    del rec, pv
    idx = np.random.default_rng(shard).choice(args.batch, args.check, replace=False)
    h = {k: v[idx].cpu().numpy() for k, v in pk.items()}
    o = orc.solve_batch(conn, fixed, h['coords'], h['props'], h['nodal_mass'], h['f_nodal'], h['r_local'], n_modes=57)
    u = out['u'][idx].cpu().numpy()
    eu = (np.abs(u - o['u']).max(axis=(1, 2)) / np.abs(o['u']).max(axis=(1, 2))).max()
    lm = lam[idx].cpu().numpy()
    el = (np.abs(lm / o['lam'] - 1) / (1e-10 + 64 * eps * o['lam'][:, -1:] / o['lam'])).max()
    # every one of the 57 eigenvalues against scipy's eigenpairs refined by an extended-precision Rayleigh quotient
    K, M = o['Kc'].astype(np.longdouble), o['Mc'].astype(np.longdouble)
    V = o['phi'][:, :, plan.pos_of_free].astype(np.longdouble)
    rq = (np.einsum('bmi,bij,bmj->bm', V, K, V) / np.einsum('bmi,bij,bmj->bm', V, M, V)).astype(np.float64)
    er = float(np.abs(lm / rq - 1).max())
    worst['lam_refined'] = max(worst['lam_refined'], er)
    assert er < 1e-10, (shard, er)
    worst['u'] = max(worst['u'], float(eu))
    worst['lam'] = max(worst['lam'], float(el))
    assert eu < 1e-10 and el < 1.0, (shard, eu, el)
print('soak ok (engine %s): %d structures, worst %s' % (args.engine, args.shards * args.batch, worst))
