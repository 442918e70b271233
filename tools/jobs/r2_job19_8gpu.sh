set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8"
for ch in 3125 12500 50000; do
  timeout 200 $TR --master-port $((29600+ch%97)) tools/bench_copy.py --in-bytes 1720 --out-bytes 9336 --chunk $ch > gpurun_out/r2_copy8_ch$ch.json 2>/dev/null
  grep '^{' gpurun_out/r2_copy8_ch$ch.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('copy chunk', d['chunk'], round(d['aggregate_GBps'],1), 'GB/s', round(d['structures_per_s_ceiling']/1e6,2), 'M/s')"
done
for ch in 3125 12500 25000; do
  timeout 300 $TR --master-port $((29700+ch%97)) bench.py --gpus 8 --no-cpu --steps 5 --chunk $ch > gpurun_out/r2_bench8_ch$ch.json 2>/dev/null
  grep '^{' gpurun_out/r2_bench8_ch$ch.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench chunk', d['e2e']['chunk'], 'e2e', round(d['e2e']['value']/1e6,2), 'M/s', 'device', round(d['value']/1e6,2))"
done
