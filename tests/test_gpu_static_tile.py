"""The opt-in one-lane-per-structure kernel of the static-only path (beamfe2_b200/csrc/bfe2_tile.cu; switched on with
BFE2_STATIC_TILE=1, read once per process -- hence the child process)."""
import os
import subprocess
import sys

import pytest

from tests import util

pytestmark = pytest.mark.gpu


def test_tile_kernel_matches_the_oracle_and_the_group_kernel():
    env = dict(os.environ, BFE2_STATIC_TILE='1', BFE2_STATIC_TILE_MIN='1', PYTHONPATH=util.ROOT)
    r = subprocess.run([sys.executable, '-m', 'tests.static_tile_child'], cwd=util.ROOT, env=env, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith('OK'), (r.stdout[-2000:], r.stderr[-4000:])


def test_tile_kernel_is_off_by_default():
    import torch
    from beamfe2_b200 import _lib, batch, sweep
    from beamfe2_b200.plan import Plan
    conn, sup = sweep.p20_topology()
    plan = Plan.from_supports(21, conn, sup)
    pk = sweep.p20_pack(torch.from_numpy(sweep.p20_draw_params(9000)).cuda())
    n0 = _lib.lib().bfe2_static_tile_launches()
    out = batch.solve_fused(plan, pk['coords'], pk['props'], None, pk['f_nodal'][:, :1].contiguous(),
                            pk['r_local'][:, :1].contiguous(), modal=False)
    assert (out['info'] == 0).all() and _lib.lib().bfe2_static_tile_launches() == n0
