set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t3.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_t3.log
tail -8 gpurun_out/r2_t3.log
timeout 600 python bench.py --no-cpu > gpurun_out/r2_bench1_b.json 2> gpurun_out/r2_bench1_b.err
cat gpurun_out/r2_bench1_b.json
