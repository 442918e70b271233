set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python bench.py --no-cpu > gpurun_out/r2_b20.json 2>/dev/null
python -c "
import json
d=json.load(open('gpurun_out/r2_b20.json')); print('bench', round(d['value']), d['ms_per_step'], 'sweeps', d['mean_sweeps'], 'e2e', round(d['e2e']['value']), d['e2e']['matches_device_run'], d['failed_structures'])"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout 900 python tools/soak.py --shards 12 2>&1 | tail -2
timeout 600 python tools/soak_midsize.py 2>&1 | tail -4
