set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_t10.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_t10.log
tail -12 gpurun_out/r2_t10.log
timeout 300 python tools/bench_beam.py --n 100000 > gpurun_out/r2_beam_1e5_final.json 2> gpurun_out/r2_beam_1e5_final.err
timeout 900 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; tail -2 gpurun_out/r2_bench_1gpu.err
timeout 300 python bench.py --no-cpu --steps 3 --warmup 3 > gpurun_out/plain_b.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --no-cpu --steps 3 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_solve -s 3 -c 1 -o gpurun_out/r2_fused_final python bench.py --no-cpu --steps 3 --warmup 3 > gpurun_out/ncu_fused.log 2>&1
tail -2 gpurun_out/ncu_fused.log
