"""The reference's Examples/*.py scripts run UNCHANGED through the GPU boundary and leave the same results behind as
in the unmodified reference on the CPU (the example scripts are the named configs 1 and 2).

Every script is executed three times in fresh interpreters (tests/run_example_child.py): by the unmodified reference
(CPU), with BeamFE2.* aliased to the beamfe2_b200 drop-in modules, and with the unmodified reference whose two solve()
bodies are replaced by the ctypes stub of INTEGRATION.md option B.  Only Structure.draw is neutralised (matplotlib).
Four of the seven scripts stop at `section.I['x']` in the reference itself: unchanged they must stop
with the same exception on all three sides; with that subscript fixed in memory (--fix) they run to the end."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu

REF = os.path.join(util.ROOT, 'baseline', '_ref')
STUB = os.path.join(util.ROOT, 'oracle', 'mpl_stub')
SHIM = os.path.join(util.ROOT, 'tests', 'refsuite_shim')
EXAMPLES = ['Chopra_102', 'Chopra_103', 'Chopra_105', 'cantilever_beam', 'cantilever_beam_hangman',
            'clamped_clamped_beam', 'modal_masses']
BROKEN_IN_THE_REFERENCE = {'cantilever_beam', 'cantilever_beam_hangman', 'clamped_clamped_beam', 'modal_masses'}
FLOAT = re.compile(r'[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?')


def run(mode, example, fix):
    path = os.path.join(REF, 'Examples', example + '.py')
    if not os.path.isfile(path):
        pytest.skip('baseline/_ref/Examples is not installed (run __graft_entry__.build() where the reference checkout exists)')
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE='1', BFE2_REFERENCE_PACKAGE=os.path.join(REF, 'BeamFE2'))
    env['PYTHONPATH'] = os.pathsep.join(([SHIM] if mode == 'alias' else []) + [STUB, REF, util.ROOT])
    cmd = [sys.executable, os.path.join(util.ROOT, 'tests', 'run_example_child.py'), mode, path] + (['--fix'] if fix else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=util.ROOT)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith('RESULT ')]
    assert r.returncode == 0 and len(lines) == 1, (r.stdout[-2000:], r.stderr[-3000:])
    return json.loads(lines[0][7:])


def close(a, b, tol):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    assert a.shape == b.shape
    return np.abs(a - b).max() <= tol * max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize('fix', [False, True], ids=['unchanged', 'subscript-fixed'])
@pytest.mark.parametrize('example', EXAMPLES)
def test_example_script_leaves_the_same_results(example, fix):
    if fix and example not in BROKEN_IN_THE_REFERENCE:
        pytest.skip('nothing to fix in this script')
    real = run('real', example, fix)
    if not fix and example in BROKEN_IN_THE_REFERENCE:
        assert 'TypeError' in real['error']                       # the reference's own example bug
    for mode in ('alias', 'patch'):
        got = run(mode, example, fix)
        assert got.get('error') == real.get('error'), (mode, got.get('error'), real.get('error'))
        assert sorted(k for k in got if k != 'mode') == sorted(k for k in real if k != 'mode')
        # static: the reference inverts the penalty matrix (cond ~ 1e41 / k_min), its own accuracy is ~1e-9
        for key in ('u', 'reactions'):
            if key in real:
                assert close(got[key], real[key], 1e-8), (mode, key)
        if 'frequencies' in real:
            assert len(got['frequencies']) == len(real['frequencies'])
            fa, fb = np.array(got['frequencies']), np.array(real['frequencies'])
            # the reference's LAPACK route (dsygvd) is accurate relative to the LARGEST eigenvalue: error of omega_k^2
            # ~ eps lam_max / lam_k (modal_masses.py: lam_max / lam_min = 4e9 -> 5e-7 on the first frequency; the GPU
            # path keeps relative accuracy, tests/test_gpu_parity.py)
            tol = np.maximum(1e-8, 25 * np.finfo(float).eps * (fb.max() / fb) ** 2)
            assert (np.abs(fa / fb - 1) < tol).all(), (mode, np.abs(fa / fb - 1), tol)
            assert np.abs(np.array(got['modal_mass']) - 1).max() < 1e-9 and np.abs(np.array(real['modal_mass']) - 1).max() < 1e-9
            lam = np.array(real['frequencies']) ** 2
            m = len(lam) - 1                                       # the last mode's cluster status is undecidable
            err = util.mode_err_clustered(np.array(got['shapes'])[:m], np.array(real['shapes'])[:m], lam, rel_gap=1e-3)
            assert np.nanmax(err) < 1e-6, (mode, err)
        # what the script printed: same number of numbers, same values to the printed precision
        pa, pb = [float(t) for t in FLOAT.findall(got['stdout'])], [float(t) for t in FLOAT.findall(real['stdout'])]
        assert len(pa) == len(pb), (mode, got['stdout'][-1500:], real['stdout'][-1500:])
        if pb and 'shapes' not in real:                            # printed mode shapes carry an arbitrary sign
            assert close(pa, pb, 1e-6), mode
