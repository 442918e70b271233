set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python bench.py --no-cpu > gpurun_out/r2_b18_default.json 2>/dev/null
BFE2_FORM_F=1 timeout 300 python bench.py --no-cpu > gpurun_out/r2_b18_formf.json 2>/dev/null
python -c "
import json
for f in ('r2_b18_default','r2_b18_formf'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, round(d['value']), d['ms_per_step'], 'sweeps', d['mean_sweeps'], 'e2e ok', d['e2e']['matches_device_run'], d['failed_structures'])"
BFE2_FORM_F=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "truth or scipy_batch or golden or large_batch" 2>&1 | tail -6
BFE2_FORM_F=1 timeout 900 python tools/soak.py --shards 8 2>&1 | tail -2
