"""Pins the CPU oracle (oracle/beamfe2_oracle.py) to the reference:
   (1) every golden case produced by the REAL reference (tests/golden, made by oracle/make_golden.py),
   (2) the reference's own known-answer tests, restated here with their file:line,
   (3) live against the reference checkout when it is present (build container only; CPU test)."""
import math

import numpy as np
import pytest

from oracle import beamfe2_oracle as orc
from oracle import ref_harness as rh
from tests import util

CASES = sorted(util.golden_models().keys())


def run_oracle(model, static='penalty'):
    pm = util.pack(model)
    n_free = 3 * len(model['nodes']) - len(pm['fixed'])
    has_lc = pm['f_nodal'].shape[0] > 0
    return pm, orc.solve_batch(pm['conn'], pm['fixed'], pm['coords'], pm['props'],
                               nodal_mass=pm['nodal_mass'] if n_free > 0 else None,
                               f_nodal=pm['f_nodal'] if has_lc else None,
                               r_local=pm['r_local'] if has_lc else None,
                               lumped=pm['lumped'], static=static)


@pytest.mark.parametrize('case', CASES)
def test_oracle_matches_reference_golden(case):
    model = util.golden_models()[case]
    g = util.golden(case)
    pm, o = run_oracle(model)
    # INT: DOF map is bit-exact
    assert np.array_equal(np.asarray(pm['fixed'], dtype=np.int64), g['positions_to_eliminate'])
    assert int(g['sumdof']) == 3 * len(model['nodes'])
    # element + assembly
    L, alpha = orc.length_direction(pm['coords'], pm['conn'])
    assert np.array_equal(L, g['length'])
    assert np.array_equal(alpha, g['direction'])
    Kg, Mg = orc.element_km_global(pm['coords'], pm['conn'], pm['props'], pm['lumped'])
    assert util.relmax(Kg, g['Ke_global']) < 1e-15
    assert util.relmax(Mg, g['Me_global']) < 1e-15
    assert util.relmax(o['K'], g['K']) < 1e-15
    if 'f' in g:
        assert util.relmax(o['f'], g['f']) < 1e-15
        # the penalty inverse is ill-conditioned by construction (1e41): compare on the free DOFs
        keep = orc.free_positions(len(model['nodes']), pm['fixed'])
        if len(keep):
            assert util.relmax(o['u'][:, keep], g['u'][:, keep]) < 1e-9
        assert util.relmax(o['endforces'], g['endforces']) < 1e-9
        assert util.relmax(o['reactions'], g['reactions']) < 1e-9
        # and the condensed solve (north_star target) agrees with the reference on the free DOFs
        _, o2 = run_oracle(model, static='condensed')
        if len(keep):
            assert util.relmax(o2['u'][:, keep], g['u'][:, keep]) < 1e-9
    if 'circular_freq' in g and 'lumped' not in case:
        assert util.relmax(o['M'], g['M_with_masses']) < 1e-15
        lam = o['lam']
        w = np.sqrt(lam[lam > 0])
        assert w.shape == g['circular_freq'].shape
        assert np.abs(w / g['circular_freq'] - 1).max() < 1e-9
        assert util.mode_err(o['phi'], g['shapes']).max() < 1e-6


def test_unit_beam_known_answer():
    """Tests/test_beam_linear_elastic.py:37-46 -- integer stiffness matrix of the unit beam."""
    k = orc.ke_local(np.array(1.), np.array(1.), np.array(1.), np.array(1.))
    expected = np.array([[1, 0, 0, -1, 0, 0], [0, 12, 6, 0, -12, 6], [0, 6, 4, 0, -6, 2],
                         [-1, 0, 0, 1, 0, 0], [0, -12, -6, 0, 12, -6], [0, 6, 2, 0, -6, 4]], dtype=float)
    assert np.array_equal(k, expected)
    assert np.array_equal(k, k.T)                                   # :48-51
    with pytest.raises(np.linalg.LinAlgError):                       # :53-55 not positive definite
        np.linalg.cholesky(k)


def test_cantilever_closed_forms():
    """Tests/test_beam_linear_elastic.py:72-115 -- tip FY: PL^3/3EI, PL^2/2EI; tip MZ: ML^2/2EI, ML/EI."""
    E, A, I, L = 2.1e5, 30.0, 22.5, 100.0
    coords = np.array([[0., 0.], [L, 0.]])
    conn = np.array([[0, 1]])
    props = np.array([[E, A, I, 1.0]])
    f = np.zeros((3, 6))
    Px, Py, Mz = 137.0, -42.0, 991.0
    f[0, 3], f[1, 4], f[2, 5] = Px, Py, Mz
    o = orc.solve_batch(conn, [0, 1, 2], coords, props, f_nodal=f)
    assert math.isclose(o['u'][0, 3], Px * L / (E * A), rel_tol=1e-12)
    assert math.isclose(o['u'][1, 4], Py * L ** 3 / (3 * E * I), rel_tol=1e-12)
    assert math.isclose(o['u'][1, 5], Py * L ** 2 / (2 * E * I), rel_tol=1e-12)
    assert math.isclose(o['u'][2, 4], Mz * L ** 2 / (2 * E * I), rel_tol=1e-12)
    assert math.isclose(o['u'][2, 5], Mz * L / (E * I), rel_tol=1e-12)


def test_fixed_end_vectors():
    """Tests/test_internal_beamloads.py:74-85, 117-128, 159-170 -- reactions of the clamped-clamped beam
    are minus the local fixed-end vectors."""
    L = np.array(100.0)
    r = orc.beam_load_reactions_local('uniform perpendicular force', -1.0, L)
    assert np.allclose(-r, [0, 50, 833.3333333333334, 0, 50, -833.3333333333334], rtol=1e-14)
    r = orc.beam_load_reactions_local('concentrated perpendicular force', -30, L, 0.3)
    assert np.allclose(-r, [0, 23.52, 441, 0, 6.48, -189], rtol=1e-12)
    r = orc.beam_load_reactions_local('concentrated moment', 500, L, 0.3)
    assert np.allclose(-r, [0, 6.3, -35., 0, -6.3, 165.], rtol=1e-12)
    g = util.golden('clamped_clamped_loads')
    assert np.allclose(g['reactions'][0], [0, 50, 833.33333333, 0, 50, -833.33333333], atol=1e-5)
    assert np.allclose(g['reactions'][1], [0, 23.52, 441, 0, 6.48, -189], atol=1e-5)


def test_reference_golden_frequencies():
    """Tests/test_beam_allresults.py:97-113, 206-224 -- the frequency lists the reference asserts."""
    exp1 = [7.613434040661384, 11.09323127793537, 15.326172330326486, 32.693985190216615, 41.63737293315384,
            54.555488726848836, 84.23566717176273, 110.95431530916404, 138.32201256527927, 432.24810222320696,
            1326.485590168174, 2310.0531736149446, 3428.9430318147342, 4633.931252454926, 5560.428871743147]
    exp2 = [1.8098134635564367, 11.886156983577484, 15.868886208391775, 46.8696497539431, 56.91143213363919,
            82.10962203501121, 101.00378682424322, 146.20319107194706, 172.68133458649052, 191.3400650417073,
            278.6082063770441, 314.0478965189491, 347.15683024183244, 429.28119919168194, 606.6507091264467,
            626.4477915106148, 705.2507665620494, 873.2857713706983, 922.718433097769, 1147.3740125153104,
            1231.6447012112453, 1861.2153655737682, 2520.6433978302216]
    for case, exp, head in (('allresults_1', exp1, None), ('allresults_2', exp2, None),
                            ('allresults_1_mass', [0.10469612329002663, 9.231045221012806, 10.311990761936627,
                                                   13.586565471302825, 32.34979734510168], 5),
                            ('allresults_2_mass', [0.531783218671432, 11.802244251990954, 15.313592559544537,
                                                   46.307812940599206, 56.707579685571524], 5)):
        _, o = run_oracle(util.golden_models()[case])
        fr = np.sqrt(o['lam']) / (2 * math.pi)
        fr = fr[:head] if head else fr
        assert np.allclose(fr, exp, atol=1e-5), case


def test_reference_golden_static_vectors():
    """Tests/test_beam_allresults.py:43-95, 161-204 -- reactions and displacements the reference asserts."""
    _, o = run_oracle(util.golden_models()['allresults_1'])
    exp_r = [0, 481.971154, 0, 0, 0, 0, 0, 608.173077, 0, 0, 0, 0, 0, 4848.557692, 0, 0, 0, 0, 1000., 10061.298077,
             -1739182.692308]
    assert np.allclose(o['reactions'][0], exp_r, atol=1e-5)
    exp_u = [-0.47619, 0, -12.591575, -0.396825, -4170.694275, 0.158985, -0.31746, 0, 11.955637, -0.238095,
             5898.326211, 12.432591, -0.15873, 0, -35.230973, -0.079365, -12671.067359, 6.052011, 0, 0, 0]
    assert np.allclose(o['u'][0], exp_u, atol=1e-3)
    _, o = run_oracle(util.golden_models()['allresults_2'])
    assert np.allclose(o['reactions'][0][[0, 1, 24, 25]], [2848.39394, -669.678788, 1651.60606, -3830.321212], atol=1e-5)
    _, o = run_oracle(util.golden_models()['allresults_4'])          # :311-349
    exp = np.zeros(21)
    exp[18:] = [500., 1000., -3000000.]
    assert np.allclose(o['reactions'][0], exp, atol=1e-5)
    ef0 = o['endforces'][0, 0]
    # internal_action queries :338-349 (axial 1000 at pos 0; baseline only -> +load curve on the host)
    assert math.isclose(orc.internal_action_baseline(ef0, 'axial', 0.0), 1000, abs_tol=1e-8)
    assert math.isclose(orc.internal_action_baseline(ef0, 'shear', 0.5), 1000, abs_tol=1e-8)


def test_chopra_and_cantilever_configs():
    """SURVEY.md section 8c golden values of the named configs (measured by running the examples)."""
    _, o = run_oracle(util.golden_models()['chopra_102'])
    w = np.sqrt(o['lam'])
    assert np.allclose(w, [5.916030257609508, 12.756376506226633, 30.532685775787705, 30.9719326649463,
                           647.8729257674084, 1688.6956098408298], rtol=1e-10)
    _, o = run_oracle(util.golden_models()['chopra_103'])
    assert np.allclose(np.sqrt(o['lam']), [3.972912555204209, 10.038411980942874, 19.091566031493347,
                                           32.50479596290722, 165.81366539291142, 426.6985880635858], rtol=1e-10)
    _, o = run_oracle(util.golden_models()['cantilever_example'])
    assert np.allclose(o['f'][0], [100, -1000, -33333.333333333336, 100, -1000, 33333.333333333336], rtol=1e-15)
    assert np.allclose(o['u'][0, 3:], [0.003174603174603175, -423.2804232804231, -2.8218694885361533], rtol=1e-10)
    assert np.allclose(o['reactions'][0, :3], [-200, 2000, 200000], rtol=1e-9)
    assert np.allclose(np.sqrt(o['lam']), [395.5998291935669, 3897.7207300923574, 44792.515298335195], rtol=1e-10)


def test_truth_fixture_consistent_with_scipy():
    """The 40-digit P20 truth agrees with scipy / numpy within their own FP64 error bounds."""
    import torch
    from beamfe2_b200 import sweep
    t = util.p20_truth()
    P = t['params']
    assert P.shape[0] >= 64 and np.array_equal(P[:8], sweep.p20_draw_params(8))
    ratio = t['lam'][:, -1] / t['lam'][:, 0]
    assert ratio.min() < 5e5 and ratio.max() > 6e7             # picked by condition, not the first few draws
    pk = {k: v.numpy() for k, v in sweep.p20_pack(torch.from_numpy(P)).items()}
    conn, sup = sweep.p20_topology()
    o = orc.solve_batch(conn, orc.positions_to_eliminate(sup), pk['coords'], pk['props'], pk['nodal_mass'],
                        pk['f_nodal'], pk['r_local'], n_modes=57)
    eps = np.finfo(float).eps
    bound = 1e-10 + 64 * eps * o['lam'][:, -1:] / o['lam']          # LAPACK: |dlam| <= c eps ||C||
    assert (np.abs(o['lam'] / t['lam'] - 1) < bound).all()
    assert util.relmax(o['u'], t['u']) < 1e-10
    for b in range(P.shape[0]):
        e = util.mode_err_clustered(o['phi'][b, :10], t['phi'][b], t['lam'][b])
        assert np.nanmax(e) < 1e-7, (b, e)
    # one extended-precision Rayleigh-quotient step turns scipy's eigenvectors into eigenvalues that agree with the
    # 40-digit truth to the last bits -- the reference the large-batch GPU tests use for ALL 57 eigenvalues
    keep = orc.free_positions(21, orc.positions_to_eliminate(sup))
    rq = util.refined_eigenvalues(o['Kc'], o['Mc'], o['phi'][:, :, keep])
    assert np.abs(rq / t['lam'] - 1).max() < 1e-13


@pytest.mark.skipif(not rh.reference_available(), reason='reference checkout not present')
def test_oracle_live_against_reference():
    """Build container only: re-run the real reference on fresh random P20 variants."""
    from beamfe2_b200 import sweep
    P = sweep.p20_draw_params(3, shard=7)
    for row in P:
        model = sweep.p20_model_dict(row)
        ref = rh.run_reference(model)
        pm, o = run_oracle(model)
        assert util.relmax(o['K'], ref['K']) < 1e-15
        assert util.relmax(o['f'], ref['f']) < 1e-15
        keep = orc.free_positions(21, pm['fixed'])
        assert util.relmax(o['u'][:, keep], ref['u'][:, keep]) < 1e-9
        assert util.relmax(o['endforces'], ref['endforces']) < 1e-9
        assert np.abs(np.sqrt(o['lam']) / ref['circular_freq'] - 1).max() < 1e-9
