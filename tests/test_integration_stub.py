"""CPU half of INTEGRATION.md option B: integration/_bfe2.pack_structure reads a REAL reference Structure (the
unmodified package under baseline/_ref or the read-only checkout) and must produce exactly the arrays the C ABI
expects -- checked against the oracle's independent packing of the same model and against the reference's own load
vector.  (The GPU half, solve_structure + the reference's Tests/, is tests/test_reference_suite.py.)"""
import numpy as np
import pytest

from oracle import beamfe2_oracle as orc
from oracle import ref_harness as rh
from tests import util

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason='reference not installed (baseline/_ref)')


@pytest.mark.parametrize('case', ['allresults_1_mass', 'allresults_2', 'chopra_105', 'p20_v1', 'vertical_beam',
                                  'clamped_clamped_loads', 'cantilever_example', 'rod30_lumped'])
def test_pack_structure_of_a_reference_structure(case):
    from integration import _bfe2
    model = util.golden_models()[case]
    s = rh.build_structure(model)
    lcs = model.get('load_cases') or [dict()]
    for lc in lcs:
        rh.apply_load_case(s, lc)
        pk = _bfe2.pack_structure(s)
        ref = rh.pack_model(dict(model, load_cases=[lc]))
        for key in ('conn', 'fixed', 'coords', 'props', 'nodal_mass'):
            assert np.array_equal(pk[key], ref[key]), key
        assert np.array_equal(pk['lumped'].astype(bool), ref['lumped'])
        scale = max(1.0, np.abs(ref['r_local']).max(), np.abs(ref['f_nodal']).max())
        assert np.abs(pk['r_local'] - ref['r_local'][0]).max() <= 1e-12 * scale
        assert np.abs(pk['f_nodal'] - ref['f_nodal'][0]).max() <= 1e-12 * scale
        # nodal part + rotated beam part = the reference's own right-hand side (Solver.py:83)
        if s._load_vector is not None:
            f = orc.load_vector(pk['f_nodal'][None], pk['r_local'][None], pk['coords'], pk['conn'])[0]
            assert np.abs(f - np.asarray(s._load_vector).ravel()).max() <= 1e-12 * scale
