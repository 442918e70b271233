set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( time timeout 900 $TR --nproc-per-node 2 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_2gpu_full.json 2> gpurun_out/r2_bench_2gpu_full.err ) 2>&1 | tail -4
grep '^{' gpurun_out/r2_bench_2gpu_full.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['cpu_baseline']['value'], [(a['kind'],a['cores'],round(a.get('value',0))) for a in d['cpu_baseline']['arms']])"
( time timeout 900 $TR --nproc-per-node 2 --master-port 29542 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_ref_2gpu.json 2> gpurun_out/r2_ref_2gpu.err ) 2>&1 | tail -4
grep -c '^{' gpurun_out/r2_ref_2gpu.json
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_t16.log 2>&1; tail -4 gpurun_out/r2_t16.log
