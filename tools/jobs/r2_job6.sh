set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_t6.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_t6.log
tail -30 gpurun_out/r2_t6.log
timeout 600 python bench.py --no-cpu > gpurun_out/r2_bench1_te.json 2> gpurun_out/r2_bench1_te.err; tail -3 gpurun_out/r2_bench1_te.err
cat gpurun_out/r2_bench1_te.json
