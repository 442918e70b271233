"""Generate tests/golden/* from the REAL reference (imported read-only from /root/reference).

Run in the build container only:   python -m oracle.make_golden
Outputs (committed):
    tests/golden/ref_cases.json   the model dicts (inputs) of every case
    tests/golden/ref_cases.npz    what the reference computed for them ("<case>/<quantity>")
    tests/golden/p20_truth.npz    40-digit (mpmath) solutions of the assembled P20 matrices of 64 variants of
                                  the config-3 generator (made by oracle/make_truth.py): a truth that is
                                  independent of any FP64 LAPACK route (scipy's own eigenvalues are
                                  only good to ~1e-10 relative on these pencils)

Cases follow the reference's own tests / examples (file:line cited per case).
"""
from __future__ import annotations

import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh            # noqa: E402
from oracle import beamfe2_oracle as orc         # noqa: E402

STEEL_E, STEEL_RHO = 2.1e5, 7.85e-9              # Material.py:59-62


def rect(h, w):
    """BeamSections.Rectangle (BeamSections.py:41-59): A = w h, I.x = w h^3 / 12."""
    return w * h, w * h ** 3 / 12.


def cases():
    out = {}
    # --- element known answers: Tests/test_beam_linear_elastic.py:37-46 (unit beam), :140-151 (rotated)
    out['unit_beam'] = dict(nodes=[(0, 0), (1, 0)], beams=[dict(i=1, j=2, E=1., A=1., I=1., rho=1.)],
                            supports={1: ['ux', 'uy', 'rotz']},
                            load_cases=[dict(nodal=[(2, {'FX': 3.0, 'FY': -2.0, 'MZ': 1.5})])])
    out['rot_beam'] = dict(nodes=[(0, 0), (2, 1)], beams=[dict(i=1, j=2, E=1., A=1., I=1., rho=1.)],
                           supports={1: ['ux', 'uy', 'rotz']},
                           load_cases=[dict(nodal=[(2, {'FX': 1.0, 'FY': 1.0, 'MZ': 1.0})])])
    out['vertical_beam'] = dict(nodes=[(0, 0), (0, 3.5)], beams=[dict(i=1, j=2, E=210., A=2., I=3., rho=.5)],
                                supports={1: ['ux', 'uy', 'rotz']},
                                load_cases=[dict(nodal=[(2, {'FX': 1.0})],
                                                 beam=[(0, 'concentrated moment', dict(value=7.0, position=0.25))])])
    # --- config 1: Examples/cantilever_beam.py:22-53 (with I['x'] -> I.x)
    A, I = rect(3, 10)
    out['cantilever_example'] = dict(
        nodes=[(0, 0), (200.0, 0)], beams=[dict(i=1, j=2, E=STEEL_E, A=A, I=I, rho=STEEL_RHO)],
        supports={1: ['ux', 'uy', 'rotz']},
        load_cases=[dict(beam=[(0, 'concentrated axial force', dict(value=200, position=0.3)),
                               (0, 'uniform perpendicular force', dict(value=-10))])])
    # --- config 2: Examples/Chopra_102.py:14-42, Chopra_103.py:14-42
    L, mass = 10, 100000 / 386.
    b = dict(E=29000, I=320, A=63.41, rho=0.1)
    out['chopra_102'] = dict(nodes=[(0, 0), (L / 2., 0), (L, 0)], beams=[dict(i=1, j=2, **b), dict(i=2, j=3, **b)],
                             supports={1: ['ux', 'uy', 'rotz']},
                             masses=[(2, {'mx': mass * L / 2., 'my': mass * L / 2.}),
                                     (3, {'mx': mass * L / 4., 'my': mass * L / 4.})])
    out['chopra_103'] = dict(nodes=[(0, 0), (0, L), (L, L)], beams=[dict(i=1, j=2, **b), dict(i=2, j=3, **b)],
                             supports={1: ['ux', 'uy', 'rotz']},
                             masses=[(2, {'mx': 2 * mass, 'my': 2 * mass}), (3, {'mx': mass, 'my': mass})])
    # --- Examples/Chopra_105.py:17-60 two-storey frame, NOT chain numbered (half bandwidth 8)
    h = 120
    L5 = 2 * h
    m1 = 1000000 / 386.
    E5, I5, A5, r5 = 29000, 320, 63.41, 0.01
    beams5 = [dict(i=1, j=3, E=E5, I=2 * I5, A=A5, rho=r5), dict(i=3, j=5, E=E5, I=I5, A=A5, rho=r5),
              dict(i=2, j=4, E=E5, I=2 * I5, A=A5, rho=r5), dict(i=4, j=6, E=E5, I=I5, A=A5, rho=r5),
              dict(i=3, j=4, E=E5, I=2 * I5, A=I5, rho=r5), dict(i=5, j=6, E=E5, I=2 * I5, A=I5, rho=r5)]
    out['chopra_105'] = dict(
        nodes=[(0, 0), (L5, 0), (0, h), (L5, h), (0, 2 * h), (L5, 2 * h)], beams=beams5,
        supports={1: ['ux', 'uy', 'rotz'], 2: ['ux', 'uy', 'rotz']},
        masses=[(3, {'mx': m1, 'my': m1}), (4, {'mx': m1, 'my': m1}),
                (5, {'mx': m1 / 2., 'my': m1 / 2.}), (6, {'mx': m1 / 2., 'my': m1 / 2.})],
        load_cases=[dict(beam=[(k, 'uniform perpendicular force', dict(value=5.00)) for k in range(6)])])
    # --- Tests/test_beam_allresults.py:18-41 continuous beam, mixed supports
    A, I = rect(3, 10)
    cb = [dict(i=k + 1, j=k + 2, E=STEEL_E, A=A, I=I, rho=STEEL_RHO) for k in range(6)]
    out['allresults_1'] = dict(
        nodes=[(500 * x, 0) for x in range(7)], beams=cb,
        supports={1: ['uy'], 3: ['uy'], 5: ['uy'], 7: ['ux', 'uy', 'rotz']},
        load_cases=[dict(nodal=[(1, {'FX': -1000}), (2, {'FY': -1000}), (4, {'MZ': 500000})],
                         beam=[(4, 'uniform perpendicular force', dict(value=-10.0)),
                               (5, 'uniform perpendicular force', dict(value=-20.0))])])
    out['allresults_1_mass'] = dict(out['allresults_1'], masses=[(4, {'mx': 1, 'my': 1})])
    # --- Tests/test_beam_allresults.py:128-158 portal-like frame with inclined member, pinned feet
    s1, s2 = rect(100, 100), rect(200, 200)
    secs = [s1, s1, s2, s1, s1, s1, s1, s1]
    out['allresults_2'] = dict(
        nodes=[(0, 0), (0, 2000), (2000, 4000), (5000, 4000), (5000, 3000), (5000, 2000), (5000, 1000),
               (5000, 0), (5000, -1000)],
        beams=[dict(i=k + 1, j=k + 2, E=STEEL_E, A=secs[k][0], I=secs[k][1], rho=STEEL_RHO) for k in range(8)],
        supports={1: ['ux', 'uy'], 9: ['ux', 'uy']},
        load_cases=[dict(nodal=[(3, {'FX': 1000, 'FY': 500}), (5, {'FX': -1500})],
                         beam=[(1, 'uniform perpendicular force', dict(value=2.0))])])
    out['allresults_2_mass'] = dict(out['allresults_2'], masses=[(4, {'mx': 15})])
    # --- Tests/test_beam_allresults.py:289-336 clamped beam with axial beam loads
    out['allresults_4'] = dict(
        nodes=[(500 * x, 0) for x in range(7)], beams=cb, supports={7: ['ux', 'uy', 'rotz']},
        load_cases=[dict(nodal=[(1, {'FY': -1000, 'FX': -1000})],
                         beam=[(0, 'uniform axial force', dict(value=1)),
                               (2, 'concentrated axial force', dict(value=-1000, position=0.4)),
                               (4, 'concentrated axial force', dict(value=1000, position=0.99))])])
    # --- Tests/test_internal_beamloads.py:26-42, 74-85, 117-128, 159-170 clamped-clamped element
    cc = dict(nodes=[(0, 0), (100, 0)], beams=[dict(i=1, j=2, E=STEEL_E, A=A, I=I, rho=STEEL_RHO)],
              supports={1: ['ux', 'uy', 'rotz'], 2: ['ux', 'uy', 'rotz']})
    out['clamped_clamped_loads'] = dict(cc, load_cases=[
        dict(beam=[(0, 'uniform perpendicular force', dict(value=-1.0))]),
        dict(beam=[(0, 'concentrated perpendicular force', dict(value=-30, position=0.3))]),
        dict(beam=[(0, 'concentrated moment', dict(value=-200, position=0.3))])])
    # --- Tests/test_beam_linear_elastic.py:251-262 style: 3-element structure with non-collinear end
    out['three_elements'] = dict(
        nodes=[(0, 0), (100, 0), (200, 0), (200, 100)],
        beams=[dict(i=k + 1, j=k + 2, E=STEEL_E, A=A, I=I, rho=STEEL_RHO) for k in range(3)],
        supports={1: ['ux', 'uy', 'rotz']},
        load_cases=[dict(nodal=[(4, {'FX': 10.0})]), dict(nodal=[(4, {'FY': -5.0, 'MZ': 100.})]),
                    dict(nodal=[(3, {'FY': 1.0})],
                         beam=[(1, 'concentrated perpendicular force', dict(value=-3.0, position=0.6))])])
    # --- Tests/test_beam_modal.py:29-59: 30-element clamped rod (SI variant kept by the test), both masses
    Ar, Ir = rect(0.003, 0.02)
    rod = [dict(i=k + 1, j=k + 2, E=2.1e11, A=Ar, I=Ir, rho=7850) for k in range(30)]
    out['rod30_consistent'] = dict(nodes=[(k * 0.45 / 30, 0) for k in range(31)], beams=rod,
                                   supports={1: ['ux', 'uy', 'rotz']})
    out['rod30_lumped'] = dict(nodes=[(k * 0.45 / 30, 0) for k in range(31)],
                               beams=[dict(bb, lumped=True) for bb in rod], supports={1: ['ux', 'uy', 'rotz']})
    return out


def p20_cases(n=4):
    from beamfe2_b200 import sweep
    P = sweep.p20_draw_params(n)
    return {('p20_v%d' % k): sweep.p20_model_dict(P[k]) for k in range(n)}


POST_CASES = ('p20_v0', 'p20_v1', 'allresults_1', 'allresults_2', 'chopra_105')
# cases with CONCENTRATED beam loads (Loads.py:347-365, 421-432, 489-502): 11 positions, so that xi = 0.3 -- the load
# position of clamped_clamped_loads / cantilever_example -- is sampled exactly and the reference's own comparisons at
# xi == position are pinned; n_free = 0 for clamped_clamped_loads (no modal part)
POST_CASES_CONC = ('clamped_clamped_loads', 'vertical_beam', 'cantilever_example')


def post_golden(models):
    """Internal-action queries (HermitianBeam2D.internal_action) at 5 positions of every beam and the
    participation factors of Examples/modal_masses.py:40-49, computed by the REAL reference."""
    import warnings
    out = {}
    for name in POST_CASES + POST_CASES_CONC:
        pos = [0.0, 0.25, 0.5, 0.75, 1.0] if name in POST_CASES else [k / 10 for k in range(11)]
        model = models[name]
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            s = rh.build_structure(model)
            acts = []
            for lc in model['load_cases']:
                rh.apply_load_case(s, lc)
                s.solver['linear static'].solve()
                res = s.results['linear static']
                per_beam = []
                for b in s.beams:
                    d = res.element_displacements(local=True, beam=b, asvector=True)
                    per_beam.append([[float(b.internal_action(action=a, disp=d, pos=p)) for p in pos]
                                     for a in ('axial', 'shear', 'moment')])
                acts.append(per_beam)
            out[name + '/internal_actions'] = np.array(acts)
            if name in POST_CASES_CONC:
                continue
            s.solver['modal'].solve()
            m = np.asarray(s.M_with_masses)
            gam = []
            for k in range(min(6, len(s.results['modal'].displacement_results))):
                fi = np.asarray(s.results['modal'].modeshape(k))
                row = []
                for comp in (0, 1):
                    unit = np.zeros((len(m), 1))
                    unit[comp::3] = 1.0
                    row.append(float((fi.T @ m @ unit) / (fi.T @ m @ fi)))
                gam.append(row)
            out[name + '/gamma'] = np.array(gam)
    return out


def main():
    gold = os.path.join(ROOT, 'tests', 'golden')
    os.makedirs(gold, exist_ok=True)
    models = cases()
    models.update(p20_cases())
    if '--post-only' in sys.argv:
        np.savez_compressed(os.path.join(gold, 'ref_post.npz'), **post_golden(models))
        print('wrote ref_post.npz')
        return
    arrays = {}
    for name, model in models.items():
        n_free = 3 * len(model['nodes']) - len(orc.positions_to_eliminate(model['supports']))
        ref = rh.run_reference(model, modal=n_free > 0)
        for k, v in ref.items():
            arrays['%s/%s' % (name, k)] = np.asarray(v)
        print('reference case', name, 'ok')

    def enc(o):
        if isinstance(o, (np.floating, np.integer)):
            return o.item()
        raise TypeError(type(o))
    with open(os.path.join(gold, 'ref_cases.json'), 'w') as f:
        json.dump(models, f, default=enc, indent=0)
    np.savez_compressed(os.path.join(gold, 'ref_cases.npz'), **arrays)
    np.savez_compressed(os.path.join(gold, 'ref_post.npz'), **post_golden(models))
    # tests/golden/p20_truth.npz (40-digit solutions) is made by oracle/make_truth.py
    print('wrote', gold)


if __name__ == '__main__':
    main()
