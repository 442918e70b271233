"""Drop-in Structure / Solver / Results API: host bookkeeping on the CPU, and (gpu) the reference's own
test scenarios re-run through the GPU solvers with the reference's assertions and tolerances
(Tests/__init__.py: ATOL = 1e-12)."""
import math

import numpy as np
import pytest

from beamfe2_b200 import HermitianBeam_2D as HB
from beamfe2_b200 import Node, Structure, Loads
from oracle import ref_harness as rh
from tests import util

ATOL = 1e-12


def build(model):
    nodes = [Node.Node(ID=k + 1, coords=c) for k, c in enumerate(model['nodes'])]
    beams = [HB.HermitianBeam2D(ID=k + 1, i=nodes[b['i'] - 1], j=nodes[b['j'] - 1], E=b['E'], I=b['I'], A=b['A'],
                                rho=b['rho']) for k, b in enumerate(model['beams'])]
    for bm, b in zip(beams, model['beams']):
        if b.get('lumped'):
            bm.mass_matrix = 'lumped'
    s = Structure.Structure(beams=beams, supports=model['supports'])
    for nid, mass in model.get('masses', []):
        s.add_nodal_mass(nodeID=nid, mass=dict(mass))
    return s


def apply(s, lc):
    s.clear_loads()
    for nid, dyn in lc.get('nodal', []):
        s.add_nodal_load(nodeID=nid, dynam=dict(dyn))
    for bi, kind, kw in lc.get('beam', []):
        s.add_internal_loads(beam=s.beams[bi], loadtype=kind, **kw)


# ------------------------------------------------------------------------------------ CPU
@pytest.mark.parametrize('case', ['allresults_1_mass', 'allresults_2', 'allresults_4', 'chopra_105', 'p20_v1',
                                  'vertical_beam', 'clamped_clamped_loads'])
def test_host_bookkeeping_matches_reference(case):
    """Load vector, positions and packed tensors equal what the real reference produced (golden)."""
    model = util.golden_models()[case]
    g = util.golden(case)
    s = build(model)
    assert s.sumdof == int(g['sumdof'])
    assert s.positions_to_eliminate == g['positions_to_eliminate'].tolist()
    pm = rh.pack_model(model)
    for c, lc in enumerate(model.get('load_cases', [])):
        apply(s, lc)
        assert np.array_equal(np.asarray(s.load_vector).ravel(), g['f'][c])         # bit-exact
        plan, pk = s.pack()
        # nodal part = _load_vector minus the rotated beam-load part (the reference's right-hand side whatever call
        # filled it): equal to the directly applied loads up to the rounding of that subtraction
        scale = max(1.0, np.abs(g['f'][c]).max())
        assert np.abs(pk['f_nodal'][0] - pm['f_nodal'][c]).max() <= 1e-12 * scale
        assert np.array_equal(pk['r_local'][0], pm['r_local'][c])
    plan, pk = s.pack()
    for k in ('coords', 'props', 'nodal_mass'):
        assert np.array_equal(pk[k], pm[k])
    assert np.array_equal(plan.fixed, g['positions_to_eliminate'].astype(np.int32))


def test_api_surface_and_error_behaviour():
    n1, n2 = Node.Node.from_dict({'ID': 1, 'coords': (0, 0)}), Node.Node(ID=2, coords=(3.0, 4.0))
    with pytest.raises(Exception):
        Node.Node.from_dict({'ID': 1})
    b = HB.HermitianBeam2D.from_dict({'ID': 1, 'i': n1, 'j': n2, 'E': 10., 'I': 2., 'A': 3., 'rho': 1.})
    with pytest.raises(Exception):
        HB.HermitianBeam2D.from_dict({'ID': 1, 'i': n1, 'j': n2})
    assert b.l == 5.0 and b.EA == 30. and b.EI == 20. and math.isclose(b.direction, math.atan2(4, 3))
    assert b.mass_matrix == 'consistent' and b.nodes == (n1, n2)
    s = Structure.Structure(beams=[b], supports={1: ['ux', 'uy', 'rotz']})
    assert sorted(s.solver) == ['buckling', 'linear static', 'modal'] and not s.results['modal'].solved
    assert s.position_in_matrix(nodeID=2, DOF='rotz') == 5 and s.position_in_matrix(nodeID=2, dynam='FY') == 4
    assert s.nodenr_dof_from_position(5) == (2, 'MZ')
    assert s.solver['linear static'].solve() is False            # no loads: Solver.py:79-80
    with pytest.raises(Exception):
        s.add_nodal_load(nodeID=7, dynam={'FX': 1})
    s.add_nodal_load(nodeID=2, dynam={'FX': 1.0})
    s.add_nodal_load(nodeID=2, dynam={'FX': 2.0, 'MZ': 1.0})
    assert np.asarray(s.load_vector).ravel().tolist() == [0, 0, 0, 3.0, 0, 1.0]
    s.add_nodal_load(nodeID=2, dynam={'FY': 5.0}, clear=True)   # clear wipes nodal AND beam loads
    assert np.asarray(s.load_vector).ravel().tolist() == [0, 0, 0, 0, 5.0, 0]
    s.add_nodal_mass(nodeID=2, mass={'mx': 2.0})
    assert np.asarray(s.mass_vector).ravel().tolist() == [0, 0, 0, 2.0, 0, 0]
    s.set_mass_matrix_type('lumped')
    assert b.mass_matrix == 'lumped'
    m = np.matrix(np.arange(36.).reshape(6, 6))
    assert s.condense(m).shape == (3, 3) and s.condense(m)[0, 0] == 21.
    # beam loads: local fixed-end vectors and their global rotation (Loads.py:100-112)
    s.add_internal_loads(beam=b, loadtype='uniform perpendicular force', value=-2.0)
    r = b.internal_loads[-1].reactions_asvector
    assert np.allclose(r, [[0, -5.0, -50 / 12., 0, -5.0, 50 / 12.]])
    glob = b.internal_loads[-1].reactions[n2]
    assert math.isclose(glob['FX'], 0.8 * 5.0) and math.isclose(glob['FY'], -0.6 * 5.0)
    assert b.internal_action_points[0] == 0 and b.internal_action_points[-1] == 1
    with pytest.raises(AssertionError):
        b.add_internal_load(loadtype='nonsense', value=1)
    assert set(Loads.BEAM_LOAD_TYPES) == set(rh_beam_load_types())


def rh_beam_load_types():
    from oracle import beamfe2_oracle as orc
    return orc.BEAM_LOAD_KINDS


def test_helpers_transfer_matrix():
    from beamfe2_b200 import helpers
    from oracle import beamfe2_oracle as orc
    for al in (0.0, 0.3, -2.1, math.atan(0.5)):
        assert np.allclose(helpers.transfer_matrix(al), orc.transfer_matrix(al), atol=0)
        assert np.allclose(helpers.transfer_matrix(al, blocks=1), orc.rotation3(al), atol=0)
        assert np.allclose(helpers.transfer_matrix(math.degrees(al), asdegree=True), orc.transfer_matrix(al), atol=1e-15)
    assert helpers.transfer_matrix(0.1, blocks=1, blocksize=2).shape == (2, 2)
    assert helpers.np_matrix_tolist(np.matrix([[1, 2], [3, 4]])) == [1, 2, 3, 4]


# ------------------------------------------------------------------------------------ GPU
gpu = pytest.mark.gpu


@gpu
def test_helpers_compile_global_matrix():
    from beamfe2_b200 import helpers
    for case in ('allresults_2', 'chopra_105'):
        g = util.golden(case)
        s = build(util.golden_models()[case])
        K = helpers.compile_global_matrix(s.beams, stiffness=True)
        M = helpers.compile_global_matrix(s.beams, mass=True)
        assert util.relmax(np.tril(K), np.tril(g['K'])) < 1e-14
        assert util.relmax(np.tril(M), np.tril(g['M'])) < 1e-14


@gpu
def test_unit_beam_matrices():
    """Tests/test_beam_linear_elastic.py:37-70."""
    n1, n2 = Node.Node(ID=1, coords=(0, 0)), Node.Node(ID=2, coords=(1, 0))
    b = HB.HermitianBeam2D(ID=1, i=n1, j=n2, E=1., I=1., A=1., rho=1.)
    exp = np.array([[1, 0, 0, -1, 0, 0], [0, 12, 6, 0, -12, 6], [0, 6, 4, 0, -6, 2], [-1, 0, 0, 1, 0, 0],
                    [0, -12, -6, 0, 12, -6], [0, 6, 2, 0, -6, 4]], dtype=float)
    assert np.allclose(b.Ke, exp, atol=ATOL)
    assert np.allclose(b.Ke, b.Ke.T, atol=ATOL)
    with pytest.raises(np.linalg.LinAlgError):
        np.linalg.cholesky(b.Ke)
    s = Structure.Structure(beams=[b], supports={1: ['ux', 'uy', 'rotz']})
    assert s.stiffness_matrix_is_symmetric and s.node_numbering_is_ok and s.stiffness_matrix_is_ok
    assert np.allclose(s.K, exp, atol=ATOL)
    assert s.K_with_BC[0, 0] > 1e40
    free = Structure.Structure(beams=[b], supports={})
    assert not free.stiffness_matrix_is_nonsingular


@gpu
def test_cantilever_tip_loads_closed_form():
    """Tests/test_beam_linear_elastic.py:72-115, 162-222 -- one element, also rotated by atan(0.5)."""
    E, A, I = 2.1e5, 30.0, 22.5
    for (x, y) in ((100.0, 0.0), (200.0, 100.0)):
        n1, n2 = Node.Node(ID=1, coords=(0, 0)), Node.Node(ID=2, coords=(x, y))
        b = HB.HermitianBeam2D(ID=1, i=n1, j=n2, E=E, I=I, A=A, rho=7.85e-9)
        s = Structure.Structure(beams=[b], supports={1: ['ux', 'uy', 'rotz']})
        L, al = b.l, b.direction
        P = 123.0
        # axial load along the member, transverse load, end moment
        s.add_nodal_load(nodeID=2, dynam={'FX': P * math.cos(al), 'FY': P * math.sin(al)}, clear=True)
        assert s.solver['linear static'].solve() is True
        d = s.results['linear static'].element_displacements(local=True, beam=b, asvector=True)
        assert math.isclose(d[3, 0], P * L / (E * A), rel_tol=1e-10) and abs(d[4, 0]) < 1e-9 * abs(d[3, 0]) + 1e-12
        s.add_nodal_load(nodeID=2, dynam={'FX': -P * math.sin(al), 'FY': P * math.cos(al)}, clear=True)
        s.solver['linear static'].solve()
        d = s.results['linear static'].element_displacements(local=True, beam=b, asvector=True)
        assert math.isclose(d[4, 0], P * L ** 3 / (3 * E * I), rel_tol=1e-10)
        assert math.isclose(d[5, 0], P * L ** 2 / (2 * E * I), rel_tol=1e-10)
        s.add_nodal_load(nodeID=2, dynam={'MZ': P}, clear=True)
        s.solver['linear static'].solve()
        d = s.results['linear static'].element_displacements(local=True, beam=b, asvector=True)
        assert math.isclose(d[4, 0], P * L ** 2 / (2 * E * I), rel_tol=1e-10)
        assert math.isclose(d[5, 0], P * L / (E * I), rel_tol=1e-10)
        g = s.results['linear static'].global_displacements()
        assert sorted(g) == ['rotz', 'ux', 'uy'] and g['ux'].shape == (2, 1)


@gpu
@pytest.mark.parametrize('case', ['allresults_1', 'allresults_2', 'allresults_4', 'chopra_105', 'three_elements',
                                  'clamped_clamped_loads', 'cantilever_example', 'p20_v2'])
def test_static_through_dropin(case):
    model = util.golden_models()[case]
    g = util.golden(case)
    s = build(model)
    fixed = np.array(s.positions_to_eliminate, dtype=int)
    keep = np.setdiff1d(np.arange(s.sumdof), fixed)
    for c, lc in enumerate(model['load_cases']):
        apply(s, lc)
        assert s.solver['linear static'].solve() is True
        res = s.results['linear static']
        assert res.solved
        u = np.asarray(res.displacement_results[0]).ravel()
        assert res.displacement_results[0].shape == (s.sumdof, 1)
        if len(keep):
            assert util.relmax(u[keep], g['u'][c][keep]) < 1e-9
        r = np.asarray(res.reaction_forces_asvector).ravel()
        assert util.relmax(r, g['reactions'][c]) < 1e-9
        for e, b in enumerate(s.beams):
            d = res.element_displacements(local=True, beam=b, asvector=True)
            q = np.asarray(b.nodal_reactions_asvector(disps=d)).ravel()
            assert np.abs(q - g['endforces'][c][e]).max() <= 1e-9 * np.abs(g['endforces'][c]).max()
            assert np.abs(res.element_end_forces[e] - g['endforces'][c][e]).max() <= 1e-9 * np.abs(g['endforces'][c]).max()
    rf = s.results['linear static'].reaction_forces
    assert sorted(rf) == sorted(model['supports'])


@gpu
def test_allresults_asserted_values():
    """Tests/test_beam_allresults.py:43-113, 311-349 with the reference's own tolerances."""
    m = util.golden_models()['allresults_1']
    s = build(m)
    apply(s, m['load_cases'][0])
    s.solver['linear static'].solve()
    exp_r = [0, 481.971154, 0, 0, 0, 0, 0, 608.173077, 0, 0, 0, 0, 0, 4848.557692, 0, 0, 0, 0, 1000., 10061.298077,
             -1739182.692308]
    assert np.allclose(np.asarray(s.results['linear static'].reaction_forces_asvector).ravel(), exp_r, atol=1e-5)
    s.solver['modal'].solve()
    exp_f = [7.613434040661384, 11.09323127793537, 15.326172330326486, 32.693985190216615, 41.63737293315384,
             54.555488726848836, 84.23566717176273, 110.95431530916404, 138.32201256527927, 432.24810222320696,
             1326.485590168174, 2310.0531736149446, 3428.9430318147342, 4633.931252454926, 5560.428871743147]
    assert np.allclose(s.results['modal'].frequencies, exp_f, atol=1e-5)
    assert np.allclose(s.results['modal'].periods, [1 / f for f in exp_f], rtol=1e-6)
    s.add_nodal_mass(nodeID=4, mass={'mx': 1, 'my': 1})
    s.solver['modal'].solve()
    assert np.allclose(s.results['modal'].frequencies[0:5], [0.10469612329002663, 9.231045221012806, 10.311990761936627,
                                                             13.586565471302825, 32.34979734510168], atol=1e-5)
    m = util.golden_models()['allresults_4']
    s = build(m)
    apply(s, m['load_cases'][0])
    s.solver['linear static'].solve()
    beam = s.beams[0]
    disp = s.results['linear static'].element_displacements(local=True, beam=beam, asvector=True)
    assert abs(beam.internal_action(disp=disp, action='axial', pos=0.) - 1000) < 1e-6
    assert abs(beam.internal_action(disp=disp, action='axial', pos=0.5) - 750) < 1e-6
    assert abs(beam.internal_action(disp=disp, action='axial', pos=1.) - 500) < 1e-6
    assert abs(beam.internal_action(disp=disp, action='moment', pos=0.)) < 1e-3
    assert abs(beam.internal_action(disp=disp, action='shear', pos=0.5) - 1000) < 1e-6
    with pytest.raises(AssertionError):
        beam.internal_action(action='axial', pos=1.2, disp=disp)
    with pytest.raises(AssertionError):
        beam.internal_action(action='woos', pos=0.2, disp=disp)
    assert s.results['linear static'].reaction_forces == {7: {'FX': 500.0, 'FY': 1000.0, 'MZ': -3000000.0}}


@gpu
@pytest.mark.parametrize('case', ['chopra_102', 'chopra_103', 'cantilever_example', 'allresults_2_mass'])
def test_modal_through_dropin(case):
    model = util.golden_models()[case]
    g = util.golden(case)
    s = build(model)
    assert s.solver['modal'].solve() is True
    res = s.results['modal']
    w = np.array(res.circular_frequencies)
    assert w.shape == g['circular_freq'].shape
    assert np.abs(w / g['circular_freq'] - 1).max() < 1e-9
    assert len(res.displacement_results) == len(g['shapes']) and res.modeshape(0).shape == (s.sumdof, 1)
    shapes = np.stack([np.asarray(x).ravel() for x in res.displacement_results])
    assert util.mode_err(shapes[:3], g['shapes'][:3]).max() < 1e-8
    M = np.asarray(s.M_with_masses)
    assert util.relmax(M, g['M_with_masses']) < 1e-14
    assert np.abs(shapes @ M @ shapes.T - np.eye(len(shapes))).max() < 1e-9


@gpu
def test_modal_returns_false_without_mass():
    n1, n2 = Node.Node(ID=1, coords=(0, 0)), Node.Node(ID=2, coords=(1, 0))
    b = HB.HermitianBeam2D(ID=1, i=n1, j=n2, E=1., I=1., A=1., rho=0.)
    s = Structure.Structure(beams=[b], supports={1: ['ux', 'uy', 'rotz']})
    assert s.solver['modal'].solve() is False and not s.results['modal'].solved


@gpu
def test_solve_structures_batch_matches_single_solves():
    """Many drop-in Structures of one topology in one launch == solving them one by one."""
    from beamfe2_b200 import Solver, sweep
    P = sweep.p20_draw_params(6)
    many, single = [], []
    for row in P:
        for lst in (many, single):
            s = build(sweep.p20_model_dict(row))
            apply(s, sweep.p20_model_dict(row)['load_cases'][2])
            lst.append(s)
    flags = Solver.solve_structures(many)
    assert flags == [(True, True)] * 6
    for a, b in zip(many, single):
        assert b.solver['linear static'].solve() and b.solver['modal'].solve()
        # static: the batched call runs inside the fused static+modal kernel, the single solve in the static-only
        # kernel (closed-form beam stiffness): same algorithm, different rounding -> the north-star tolerance
        assert util.relmax(a.results['linear static'].displacement_results[0],
                           b.results['linear static'].displacement_results[0]) < 1e-10
        assert util.relmax(a.results['linear static'].reaction_forces_asvector,
                           b.results['linear static'].reaction_forces_asvector) < 1e-10
        assert a.results['modal'].circular_frequencies == b.results['modal'].circular_frequencies
        assert np.array_equal(a.results['modal'].modeshape(3), b.results['modal'].modeshape(3))
    other = build(util.golden_models()['chopra_102'])
    with pytest.raises(ValueError):
        Solver.solve_structures([many[0], other])


def test_pack_renumbers_wide_bands():
    """A frame numbered column by column has a band as wide as the structure; Structure.pack() lets the plan
    renumber the free DOFs (RCM) -- host-side only, no GPU needed."""
    nodes = [Node.Node(ID=k + 1, coords=(1000.0 * (k % 6), 1000.0 * (k // 6))) for k in range(24)]
    beams, bid = [], 1
    for k in range(24):
        if k % 6 < 5:
            beams.append(HB.HermitianBeam2D(ID=bid, i=nodes[k], j=nodes[k + 1], E=2.1e5, I=1e6, A=1e3, rho=7.85e-9)); bid += 1
        if k + 6 < 24:
            beams.append(HB.HermitianBeam2D(ID=bid, i=nodes[k], j=nodes[k + 6], E=2.1e5, I=1e6, A=1e3, rho=7.85e-9)); bid += 1
    s = Structure.Structure(beams=beams, supports={1: ['ux', 'uy', 'rotz'], 6: ['ux', 'uy', 'rotz']})
    plan, pk = s.pack()
    from beamfe2_b200.plan import Plan
    plain = Plan(24, plan.conn, s.positions_to_eliminate)
    assert plan.hbw <= plain.hbw and sorted(plan.pos_of_free.tolist()) == sorted(plain.pos_of_free.tolist())


def test_reactions_subtract_only_recorded_nodal_loads():
    """Results.py:119-123: the reference subtracts structure.nodal_loads (what add_nodal_load recorded); a load put
    straight into the load vector with add_single_dynam_to_node (Examples/cantilever_beam_hangman.py) is solved for
    but not subtracted.  The kernel subtracts the whole nodal right-hand side; the result object puts the
    unrecorded part back (host logic, no GPU needed: the kernel output is emulated)."""
    from beamfe2_b200 import Results
    n1, n2 = Node.Node(ID=1, coords=(0.0, 0.0)), Node.Node(ID=2, coords=(100.0, 0.0))
    b = HB.HermitianBeam2D(ID=1, i=n1, j=n2, E=2.1e5, I=22.5, A=30.0, rho=7.85e-9)
    s = Structure.Structure(beams=[b], supports={1: ['ux', 'uy', 'rotz']})
    s.add_nodal_load(nodeID=2, dynam={'FX': 7.0}, clear=True)               # recorded
    s.add_single_dynam_to_node(nodeID=2, dynam={'FY': 100.0})               # load vector only
    _, pk = s.pack()
    assert pk['f_nodal'][0].tolist() == [0, 0, 0, 7.0, 100.0, 0]
    end_forces_global = np.array([-7.0, -100.0, -1e4, 7.0, 100.0, 0.0])     # equilibrium of the cantilever
    res = Results.LinearStaticResult(structure=s, displacements=[np.matrix(np.zeros((6, 1)))],
                                     reactions=end_forces_global - pk['f_nodal'][0], f_nodal=pk['f_nodal'][0])
    res.solved = True
    assert np.allclose(np.asarray(res.reaction_forces_asvector).ravel(), [-7.0, -100.0, -1e4, 0.0, 100.0, 0.0])
    assert res.reaction_forces == {1: {'FX': -7.0, 'FY': -100.0, 'MZ': -1e4}}


def test_shape_functions_and_load_classes_match_the_reference():
    """Host-side helpers a user script may touch: Hermite shape functions / deflected shape
    (HermitianBeam_2D.py:176-222, 307-334) and the named beam-load classes (Loads.py:163-502) -- against the real
    reference when it is installed."""
    n1, n2 = Node.Node(ID=1, coords=(10.0, -5.0)), Node.Node(ID=2, coords=(40.0, 35.0))
    b = HB.HermitianBeam2D(ID=1, i=n1, j=n2, E=2.1e5, I=22.5, A=30.0, rho=7.85e-9)
    L = b.l
    for x in (0.0, 0.3, 1.0):
        assert np.allclose(b.N(x, L)[:, [0, 3]].sum(), 1.0) and math.isclose(b.N2(x, L) + b.N5(x, L), 1.0)
    assert b.N3(0.0, L) == b.N6(1.0, L) == 0.0
    h = 1e-6                                                    # end slopes: N3'(0) = N6'(1) = 1 (per unit of x L)
    assert math.isclose(b.N3(h, L) / (h * L), 1.0, rel_tol=1e-5) and math.isclose(-b.N6(1 - h, L) / (h * L), 1.0, rel_tol=1e-5)
    ld = Loads.ConcentratedMoment(value=3.0, position=0.25, beam=b)
    assert isinstance(ld, Loads.BeamLoad) and ld.loadtype == 'concentrated moment' and ld.position == 0.25
    b.add_internal_load(loadtype='uniform axial force', value=2.0)
    assert isinstance(b.internal_loads[-1], Loads.UniformAxialForce) and Structure.UniformAxialForce is Loads.UniformAxialForce
    if not rh.reference_available():
        return
    ref = rh.import_reference()
    rb = ref.HermitianBeam_2D.HermitianBeam2D(ID=1, i=ref.Node.Node(ID=1, coords=(10.0, -5.0)),
                                              j=ref.Node.Node(ID=2, coords=(40.0, 35.0)), E=2.1e5, I=22.5, A=30.0,
                                              rho=7.85e-9)
    for x in (0.0, 0.17, 0.5, 0.99, 1.0):
        assert np.allclose(b.N(x=x, L=L), rb.N(x=x, L=L), rtol=1e-13, atol=1e-13)
    d = np.matrix([[0.1], [0.2], [0.003], [-0.05], [0.4], [-0.002]])
    for local in (True, False):
        mine, theirs = b.deflected_shape(local=local, scale=3.0, disps=d), rb.deflected_shape(local=local, scale=3.0, disps=d)
        assert len(mine) == len(theirs)
        for p, q in zip(mine, theirs):
            assert np.allclose(np.asarray(p, dtype=float).ravel(), np.asarray(q, dtype=float).ravel(), rtol=1e-12, atol=1e-12)
