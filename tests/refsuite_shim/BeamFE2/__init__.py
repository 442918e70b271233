"""Module-alias shim used by tests/test_reference_suite.py: `BeamFE2.{Node, HermitianBeam_2D, Structure, Solver,
Results, Loads, helpers}` resolve to the beamfe2_b200 drop-in modules, while the reference's own scalar
pre-processing modules (BeamSections, Material -- out of scope of the accelerated path) are
loaded unchanged from the reference install named by $BFE2_REFERENCE_PACKAGE.  With this directory first on
sys.path the reference's Tests/*.py run unmodified against the GPU path."""
import importlib
import importlib.util
import os
import sys

import numpy as _np

_np.set_printoptions(precision=6, suppress=True, linewidth=250)    # importing the reference does this (BeamFE2/__init__.py:2)
_REF = os.environ['BFE2_REFERENCE_PACKAGE']
for _n in ('Node', 'HermitianBeam_2D', 'Structure', 'Solver', 'Results', 'Loads', 'helpers'):
    _m = importlib.import_module('beamfe2_b200.' + _n)
    sys.modules[__name__ + '.' + _n] = _m
    globals()[_n] = _m
for _n in ('BeamSections', 'Material'):
    _s = importlib.util.spec_from_file_location(__name__ + '.' + _n, os.path.join(_REF, _n + '.py'))
    _m = importlib.util.module_from_spec(_s)
    sys.modules[_s.name] = _m
    _s.loader.exec_module(_m)
    globals()[_n] = _m
