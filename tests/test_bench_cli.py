"""bench.py command-line contract that can be checked without a GPU: the reference arm prints one JSON line with
the agreed keys; the product arm refuses to run without CUDA (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

from tests import util

BENCH = os.path.join(util.ROOT, 'bench.py')


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, BENCH, '--impl', 'reference', '--steps', '1', '--warmup', '0', '--cpu-sample', '8', '--ref-sample', '16'],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'structures/s' and d['higher_is_better'] is True
    assert d['value'] > 0 and d['dtype'] == 'f64' and d['scaling'] == 'weak' and d['vs_baseline'] is None
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    # all four CPU arms of SURVEY 8(d): the real reference (A) and the port (B), one core and all cores
    arms = d['cpu_baseline']['arms']
    assert sorted((a['kind'], a['cores'] == 1) for a in arms) == [('port', False), ('port', True), ('reference', False),
                                                                   ('reference', True)] or d['cpu_baseline']['cores'] == 1
    for a in arms:
        assert a.get('value', 0) > 0 or 'unavailable' in a
    ref = [a for a in arms if a['kind'] == 'reference' and 'value' in a]
    port1 = [a for a in arms if a['kind'] == 'port' and a['cores'] == 1][0]
    if ref:                                   # the vectorised port is the stronger baseline (what the headline uses)
        assert port1['value'] > ref[0]['value']
    assert 'workload' in d['config'] and 'model' not in d['config']


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2')
    r = subprocess.run([sys.executable, BENCH, '--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ''


@pytest.mark.skipif(torch.cuda.is_available(), reason='has a GPU')
def test_product_arm_fails_loudly_without_cuda():
    r = subprocess.run([sys.executable, BENCH, '--steps', '1', '--warmup', '0'], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0
    assert 'no CPU fallback' in (r.stderr + r.stdout)
