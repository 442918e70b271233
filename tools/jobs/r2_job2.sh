set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t2.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_t2.log
tail -15 gpurun_out/r2_t2.log
timeout 600 python bench.py > gpurun_out/r2_bench1_a.json 2> gpurun_out/r2_bench1_a.err
timeout 120 python tools/bench_copy.py > gpurun_out/r2_copy_1gpu_dense.json 2>&1
timeout 120 python tools/bench_copy.py --copies-in 5 --copies-out 7 > gpurun_out/r2_copy_1gpu_dense_12copies.json 2>&1
timeout 120 python tools/bench_copy.py --in-bytes 1720 --out-bytes 9336 > gpurun_out/r2_copy_1gpu_compact.json 2>&1
cat gpurun_out/r2_bench1_a.json
