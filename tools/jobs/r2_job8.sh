set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/r2_t8.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_t8.log
tail -8 gpurun_out/r2_t8.log
timeout 600 python bench.py --no-cpu > gpurun_out/r2_bench1_te3.json 2> gpurun_out/r2_bench1_te3.err; tail -3 gpurun_out/r2_bench1_te3.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench1_te3.json')); print('TE value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"
timeout 300 python bench.py --no-cpu --batch 20000 --steps 2 --warmup 3 > gpurun_out/plain_te.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_solve_te -s 3 -c 1 -o gpurun_out/r2_te_v3 python bench.py --no-cpu --batch 20000 --steps 2 --warmup 3 > gpurun_out/ncu_te.log 2>&1
tail -3 gpurun_out/ncu_te.log
