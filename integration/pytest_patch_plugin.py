"""pytest plugin (`-p integration.pytest_patch_plugin`): import the UNMODIFIED reference package that is on
sys.path and replace the bodies of its two solvers by the libbfe2 versions of integration/solver_patch.py
(INTEGRATION.md option B) before any test module is collected."""


def pytest_configure(config):
    import BeamFE2.Results
    import BeamFE2.Solver
    from integration import solver_patch
    solver_patch.apply(BeamFE2.Solver, BeamFE2.Results)
