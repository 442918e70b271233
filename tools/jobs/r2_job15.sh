set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_random_frames.py -m gpu -q 2>&1 | tail -5
timeout 600 python tools/bench_sizes.py jacobi > gpurun_out/r2_sizes_regs2b.txt 2>&1; tail -7 gpurun_out/r2_sizes_regs2b.txt
