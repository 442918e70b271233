set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_reference_suite.py -m gpu -q > gpurun_out/r2_t4_refsuite.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_t4_refsuite.log
tail -60 gpurun_out/r2_t4_refsuite.log
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_reference_suite.py > gpurun_out/r2_t4.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_t4.log
tail -8 gpurun_out/r2_t4.log
