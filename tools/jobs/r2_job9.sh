set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_t9.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_t9.log
tail -25 gpurun_out/r2_t9.log
timeout 600 python bench.py --no-cpu > gpurun_out/r2_bench1_c.json 2> gpurun_out/r2_bench1_c.err; tail -3 gpurun_out/r2_bench1_c.err
timeout 600 python bench.py --no-cpu --engine two_ended > gpurun_out/r2_bench1_te_final.json 2> gpurun_out/r2_bench1_te_final.err
python -c "
import json
for f in ('r2_bench1_c','r2_bench1_te_final'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['matches_device_run'])"
