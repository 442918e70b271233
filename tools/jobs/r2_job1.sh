set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t1.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_t1.log
timeout 300 python tools/bench_beam.py --n 100000 > gpurun_out/r2_beam_1e5.json 2> gpurun_out/r2_beam_1e5.err
timeout 300 python tools/bench_beam.py --n 10000 > gpurun_out/r2_beam_1e4.json 2> gpurun_out/r2_beam_1e4.err
timeout 300 python tools/bench_beam.py --n 100000 --precision fp64 --no-cpu > gpurun_out/r2_beam_1e5_fp64.json 2>&1
for p in 256 1024 4096; do timeout 300 python tools/bench_beam.py --n 100000 --no-cpu --partitions $p > gpurun_out/r2_beam_1e5_P$p.json 2>&1; done
timeout 120 python tools/bench_copy.py > gpurun_out/r2_copy_1gpu_dense.json 2>&1
timeout 120 python tools/bench_copy.py --copies-in 5 --copies-out 7 > gpurun_out/r2_copy_1gpu_dense_12copies.json 2>&1
timeout 300 python tools/bench_beam.py --n 100000 --no-cpu > gpurun_out/plain_beam.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_gram_partial|k_update' -s 2 -c 4 -o gpurun_out/r2_dmma python tools/bench_beam.py --n 100000 --no-cpu > gpurun_out/ncu_dmma.log 2>&1
tail -5 gpurun_out/r2_t1.log
