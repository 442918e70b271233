"""The scripts under examples/ run to the end on a B200 (each checks its own numbers against closed forms or prints
them; a non-zero exit or a traceback fails the test)."""
import glob
import os
import subprocess
import sys

import pytest

from tests import util

pytestmark = pytest.mark.gpu
SCRIPTS = sorted(os.path.basename(p) for p in glob.glob(os.path.join(util.ROOT, 'examples', '*.py')))


@pytest.mark.parametrize('script', SCRIPTS)
def test_example_runs(script):
    env = dict(os.environ, PYTHONPATH=util.ROOT)
    r = subprocess.run([sys.executable, os.path.join(util.ROOT, 'examples', script)], cwd=util.ROOT, env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and 'Traceback' not in r.stderr, (r.stdout[-1500:], r.stderr[-3000:])
    assert r.stdout.strip(), 'the example printed nothing'
