set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_random_frames.py tests/test_gpu_beam.py -m gpu -q -x 2>&1 | tail -25
