set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1; free -g >> gpurun_out/r2_topo.txt; nproc >> gpurun_out/r2_topo.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 1 2 4 8; do
  timeout 300 $TR --nproc-per-node $n --master-port $((29500+n)) tools/bench_copy.py > gpurun_out/r2_copy_${n}gpu_dense.json 2> gpurun_out/r2_copy_${n}gpu_dense.err
  timeout 300 $TR --nproc-per-node $n --master-port $((29510+n)) tools/bench_copy.py --in-bytes 1720 --out-bytes 9336 > gpurun_out/r2_copy_${n}gpu_compact.json 2> gpurun_out/r2_copy_${n}gpu_compact.err
done
for n in 2 4 8; do
  timeout 600 $TR --nproc-per-node $n --master-port $((29520+n)) bench.py --gpus $n --no-cpu > gpurun_out/r2_bench_${n}gpu.json 2> gpurun_out/r2_bench_${n}gpu.err
done
timeout 900 $TR --nproc-per-node 8 --master-port 29540 bench.py --gpus 8 --batch 1250000 --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_config5_8gpu.json 2> gpurun_out/r2_config5_8gpu.err
tail -3 gpurun_out/r2_config5_8gpu.err
cat gpurun_out/r2_copy_*gpu_*.json
