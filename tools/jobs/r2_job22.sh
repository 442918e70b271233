set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_t22.log 2>&1; tail -5 gpurun_out/r2_t22.log
timeout 900 python tools/soak.py --shards 24 > gpurun_out/r2_soak_jacobi.txt 2>&1; tail -1 gpurun_out/r2_soak_jacobi.txt
timeout 600 python tools/soak_midsize.py > gpurun_out/r2_soak_midsize.txt 2>&1; tail -4 gpurun_out/r2_soak_midsize.txt
timeout 900 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err
timeout 600 python tools/bench_kernels.py > gpurun_out/r2_kernels.json 2>/dev/null
timeout 600 python tools/bench_sizes.py jacobi > gpurun_out/r2_sizes_jacobi.txt 2>&1
timeout 300 python bench.py --no-cpu --steps 3 --warmup 3 > gpurun_out/plain_b.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --no-cpu --steps 3 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_solve -s 3 -c 1 -o gpurun_out/r2_fused_final python bench.py --no-cpu --steps 3 --warmup 3 > gpurun_out/ncu_fused.log 2>&1
tail -2 gpurun_out/ncu_fused.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
