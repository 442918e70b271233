set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_buckling.py -m gpu -q 2>&1 | tail -5
timeout 300 python tools/all_kernels_smoke.py > gpurun_out/r2_all_kernels.log 2>&1; tail -3 gpurun_out/r2_all_kernels.log
timeout 600 python tools/bench_sizes.py jacobi > gpurun_out/r2_sizes_jacobi.txt 2>&1
timeout 600 python tools/bench_sizes.py two_ended > gpurun_out/r2_sizes_te.txt 2>&1
cat gpurun_out/r2_sizes_jacobi.txt gpurun_out/r2_sizes_te.txt
timeout 600 python tools/bench_kernels.py > gpurun_out/r2_kernels.json 2> gpurun_out/r2_kernels.err; tail -2 gpurun_out/r2_kernels.err
timeout 300 python tools/bench_beam.py --n 100000 > gpurun_out/r2_beam_1e5_final.json 2> gpurun_out/r2_beam_1e5_final.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
