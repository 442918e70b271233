#!/usr/bin/env python
"""Per-kernel timing of the unfused path K1..K4 and of the fused kernel on P20 variants, each against
the roofline that bounds it.  Prints one JSON object; run on a B200:
    python tools/bench_kernels.py [--batch 100000] [--reps 10]
Algorithmic bytes per unit:
    K1  640 B / element            (8 f64 in, 2 x 36 f64 out)
    K2  7.1 KB / structure         (coords+props+masses in, 2 band matrices out)
    K3  8.5 KB / structure         (band + rhs in, u + end forces out; reactions add 1.5 KB)
    K4  2.10 MFLOP / structure     (canonical dense generalized eigenproblem, 11.33 n^3, n = 57)
    fused 12.7 KB and 2.11 MFLOP / structure
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from beamfe2_b200 import batch, sweep            # noqa: E402
from beamfe2_b200.plan import Plan                # noqa: E402


def timeit(fn, reps, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.add_(1.0)                            # write > L2 worth of data between iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=100000)
    ap.add_argument('--reps', type=int, default=10)
    args = ap.parse_args()
    dev = 'cuda:0'
    B = args.batch
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(
        os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {}
    hbm = float(peaks.get('hbm_gbs', 6650.0))
    conn, sup = sweep.p20_topology()
    plan = Plan.from_supports(21, conn, sup)
    pk = sweep.p20_pack(torch.from_numpy(sweep.p20_draw_params(B)).to(dev))
    flush = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device=dev)      # 256 MB > 126 MB L2
    out = {'batch': B, 'hbm_peak_gbs': hbm, 'kernels': {}}

    # K1 on the B*20 elements of the batch, SoA
    c, p = pk['coords'], pk['props']
    ci = torch.as_tensor(conn[:, 0].astype('int64'), device=dev)
    cj = torch.as_tensor(conn[:, 1].astype('int64'), device=dev)
    xi, yi = c[:, ci, 0].reshape(-1).contiguous(), c[:, ci, 1].reshape(-1).contiguous()
    xj, yj = c[:, cj, 0].reshape(-1).contiguous(), c[:, cj, 1].reshape(-1).contiguous()
    E, A, I, rho = (p[:, :, k].reshape(-1).contiguous() for k in range(4))
    nE = xi.numel()
    import ctypes
    from beamfe2_b200 import _lib
    Kg = torch.empty((nE, 6, 6), dtype=torch.float64, device=dev)
    Mg = torch.empty_like(Kg)
    L = _lib.lib()

    def k1():
        _lib.check(L.bfe2_element_km(nE, *[_lib.ptr(t) for t in (xi, yi, xj, yj, E, A, I, rho)], None,
                                     _lib.ptr(Kg), _lib.ptr(Mg), torch.cuda.current_stream().cuda_stream))
    ms = timeit(k1, args.reps, flush)
    gbs = 640.0 * nE / (ms * 1e-3) / 1e9
    out['kernels']['K1_element_km'] = {'ms': ms, 'units': nE, 'unit': 'elements', 'bytes_per_unit': 640,
                                       'achieved_gbs': gbs, 'frac_of_hbm': gbs / hbm, 'bound': 'hbm'}

    # K2
    Kb = torch.empty((B, plan.hbw + 1, plan.n_free), dtype=torch.float64, device=dev)
    Mb = torch.empty_like(Kb)

    def k2():
        _lib.check(L.bfe2_assemble_banded(plan.handle, B, _lib.ptr(pk['coords']), _lib.ptr(pk['props']),
                                          _lib.ptr(pk['nodal_mass']), _lib.ptr(Kb), _lib.ptr(Mb),
                                          torch.cuda.current_stream().cuda_stream))
    ms = timeit(k2, args.reps, flush)
    bpu = (21 * 2 + 20 * 4 + 21 * 2) * 8 + 2 * 6 * 57 * 8
    gbs = bpu * B / (ms * 1e-3) / 1e9
    out['kernels']['K2_assemble_banded'] = {'ms': ms, 'units': B, 'unit': 'structures', 'bytes_per_unit': bpu,
                                            'achieved_gbs': gbs, 'frac_of_hbm': gbs / hbm, 'bound': 'hbm'}

    # K3
    st = batch.static_solve(plan, Kb, pk['coords'], pk['props'], pk['f_nodal'], pk['r_local'])

    def k3():
        _lib.check(L.bfe2_static_solve(plan.handle, B, 3, _lib.ptr(Kb), _lib.ptr(pk['coords']), _lib.ptr(pk['props']),
                                       _lib.ptr(pk['f_nodal']), _lib.ptr(pk['r_local']), _lib.ptr(st['u']),
                                       _lib.ptr(st['endforces']), _lib.ptr(st['reactions']), _lib.ptr(st['info']),
                                       torch.cuda.current_stream().cuda_stream))
    ms = timeit(k3, args.reps, flush)
    bpu = 6 * 57 * 8 + (21 * 2 + 20 * 4) * 8 + 3 * 63 * 8 + 3 * 120 * 8 + 3 * 63 * 8 + 3 * 120 * 8 + 3 * 63 * 8
    gbs = bpu * B / (ms * 1e-3) / 1e9
    out['kernels']['K3_static_solve'] = {'ms': ms, 'units': B, 'unit': 'structures', 'bytes_per_unit': bpu,
                                         'achieved_gbs': gbs, 'frac_of_hbm': gbs / hbm, 'bound': 'hbm (unfused)',
                                         'structures_per_s': B / (ms * 1e-3)}

    # K4
    mo = batch.modal_solve(plan, Kb, Mb, n_modes=10)

    def k4():
        _lib.check(L.bfe2_modal_solve(plan.handle, B, 10, _lib.ptr(Kb), _lib.ptr(Mb), _lib.ptr(mo['lam']),
                                      _lib.ptr(mo['phi']), _lib.ptr(mo['info']), _lib.ptr(mo['sweeps']),
                                      torch.cuda.current_stream().cuda_stream))
    ms = timeit(k4, args.reps, flush)
    out['kernels']['K4_modal_solve'] = {'ms': ms, 'units': B, 'unit': 'structures', 'flop_per_unit': 2.10e6,
                                        'achieved_tflops_algorithmic': 2.10e6 * B / (ms * 1e-3) / 1e12,
                                        'bound': 'fp64 / issue', 'structures_per_s': B / (ms * 1e-3),
                                        'mean_sweeps': float(mo['sweeps'].double().mean().item())}

    # fused
    fo = batch.solve_fused(plan, pk['coords'], pk['props'], pk['nodal_mass'], pk['f_nodal'], pk['r_local'], n_modes=10)

    def kf():
        batch.solve_fused(plan, pk['coords'], pk['props'], pk['nodal_mass'], pk['f_nodal'], pk['r_local'],
                          n_modes=10, out=fo)
    ms = timeit(kf, args.reps, flush)
    gbs = 12.7e3 * B / (ms * 1e-3) / 1e9
    out['kernels']['fused'] = {'ms': ms, 'units': B, 'unit': 'structures', 'bytes_per_unit': 12700,
                               'achieved_gbs': gbs, 'frac_of_hbm': gbs / hbm, 'structures_per_s': B / (ms * 1e-3),
                               'achieved_tflops_algorithmic': 2.11e6 * B / (ms * 1e-3) / 1e12}

    # fused, static only (coords/props in, u / end forces / reactions out): the linear-static sweep
    so = batch.solve_fused(plan, pk['coords'], pk['props'], None, pk['f_nodal'], pk['r_local'], modal=False)

    def ks():
        batch.solve_fused(plan, pk['coords'], pk['props'], None, pk['f_nodal'], pk['r_local'], modal=False, out=so)
    ms = timeit(ks, args.reps, flush)
    bpu = (21 * 2 + 20 * 4) * 8 + 3 * 63 * 8 + 3 * 120 * 8 + 3 * 63 * 8 + 3 * 120 * 8 + 3 * 63 * 8
    gbs = bpu * B / (ms * 1e-3) / 1e9
    out['kernels']['fused_static_only'] = {'ms': ms, 'units': B, 'unit': 'structures', 'bytes_per_unit': bpu,
                                           'achieved_gbs': gbs, 'frac_of_hbm': gbs / hbm, 'bound': 'hbm',
                                           'structures_per_s': B / (ms * 1e-3)}
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
