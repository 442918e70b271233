set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_random_frames.py tests/test_dropin.py tests/test_gpu_buckling.py -m gpu -q 2>&1 | tail -15
timeout 600 python tools/bench_sizes.py jacobi > gpurun_out/r2_sizes_regs2.txt 2>&1; cat gpurun_out/r2_sizes_regs2.txt
BFE2_MIDSIZE_SMEM=1 timeout 600 python tools/bench_sizes.py jacobi > gpurun_out/r2_sizes_smem.txt 2>&1; tail -7 gpurun_out/r2_sizes_smem.txt
