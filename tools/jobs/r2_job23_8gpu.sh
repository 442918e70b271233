# final 8-GPU bench of the round (kernel with the eigenvalue refinement), launched the way the driver does
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 8 > gpurun_out/r2_bench_final_8gpu.log 2>gpurun_out/r2_bench_final_8gpu.err
grep '^{' gpurun_out/r2_bench_final_8gpu.log > gpurun_out/r2_bench_final_8gpu.json
python -c "import json; d=json.load(open('gpurun_out/r2_bench_final_8gpu.json')); print('N=8 device', round(d['value']/1e6,2), 'M/s  e2e', round(d['e2e']['value']/1e6,2), 'M/s  ms', d['ms_per_step'], d['clocks'])"
