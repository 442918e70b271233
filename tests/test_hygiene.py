"""Repository hygiene checks the grading contract spells out: the product package never touches the oracle
or the reference, every header symbol is bound, nothing reads /root/reference at GPU-test time."""
import os
import re

from tests import util

ROOT = util.ROOT


def _py_files(sub):
    for d, _, files in os.walk(os.path.join(ROOT, sub)):
        for f in files:
            if f.endswith('.py'):
                yield os.path.join(d, f)


def test_product_never_imports_oracle_or_reference():
    bad = []
    for path in _py_files('beamfe2_b200'):
        src = open(path).read()
        if re.search(r'^\s*(from|import)\s+(oracle|BeamFE2)\b', src, flags=re.M) or '/root/reference' in src:
            bad.append(path)
    assert not bad, bad
    for path in [os.path.join(ROOT, 'beamfe2_b200', 'csrc', f) for f in os.listdir(os.path.join(ROOT, 'beamfe2_b200', 'csrc'))]:
        if path.endswith(('.cu', '.cuh')):
            assert 'oracle' not in open(path).read()


def test_gpu_side_files_do_not_read_the_reference_checkout():
    for path in list(_py_files('tests')) + [os.path.join(ROOT, 'bench.py'), os.path.join(ROOT, '__graft_entry__.py')]:
        src = open(path).read()
        if os.path.basename(path) in ('test_hygiene.py',):
            continue
        assert '/root/reference' not in src, path      # only oracle/ref_harness.py knows that path


def test_no_cpu_fallback_in_the_loader():
    src = open(os.path.join(ROOT, 'beamfe2_b200', '_lib.py')).read()
    assert 'no CPU fallback' in src and 'raise Bfe2Error' in src
    from beamfe2_b200 import _lib
    assert os.path.basename(_lib.LIB_PATH) == 'libbfe2.so' or os.environ.get('BFE2_LIB')


def test_every_product_module_imports_on_a_cpu_only_machine():
    """Importing the package (ctypes loader, drop-in classes, subspace-iteration paths) must not need a GPU;
    only calling a solver does."""
    import importlib
    for name in ('beamfe2_b200', 'beamfe2_b200.batch', 'beamfe2_b200.beam', 'beamfe2_b200.frame_modes',
                 'beamfe2_b200.dist', 'beamfe2_b200.sweep', 'beamfe2_b200.Solver', 'beamfe2_b200.Structure'):
        importlib.import_module(name)
    from beamfe2_b200.beam import LargeBeam, SubspaceIteration
    from beamfe2_b200.frame_modes import FrameModes
    assert issubclass(LargeBeam, SubspaceIteration) and issubclass(FrameModes, SubspaceIteration)
