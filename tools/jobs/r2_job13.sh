set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/bench_frames.py --batch 256 > gpurun_out/r2_frames_150.json 2> gpurun_out/r2_frames_150.err; tail -2 gpurun_out/r2_frames_150.err; cat gpurun_out/r2_frames_150.json
timeout 600 python tools/bench_frames.py --storeys 6 --bays 8 --batch 128 > gpurun_out/r2_frames_big.json 2> gpurun_out/r2_frames_big.err; tail -2 gpurun_out/r2_frames_big.err; cat gpurun_out/r2_frames_big.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 1 --steps 3 --warmup 3 --cpu-sample 2048 --ref-sample 64 > gpurun_out/r2_bench_torchrun1.json 2> gpurun_out/r2_bench_torchrun1.err; tail -c 600 gpurun_out/r2_bench_torchrun1.json
