set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python tools/soak.py --shards 24 > gpurun_out/r2_soak_jacobi.txt 2>&1; tail -2 gpurun_out/r2_soak_jacobi.txt
timeout 900 python tools/soak.py --shards 6 --engine two_ended > gpurun_out/r2_soak_te.txt 2>&1; tail -2 gpurun_out/r2_soak_te.txt
timeout 600 python tools/soak_midsize.py > gpurun_out/r2_soak_midsize.txt 2>&1; tail -4 gpurun_out/r2_soak_midsize.txt
