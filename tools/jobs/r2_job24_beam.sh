# config 4 after the two-level partitioning: bench records at n = 1e4 / 1e5 and a partition sweep
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/bench_beam.py --n 100000 > gpurun_out/r2_config4_beam_1e5.json 2>gpurun_out/beam_1e5.err
python tools/bench_beam.py --n 10000 > gpurun_out/r2_config4_beam_1e4.json 2>gpurun_out/beam_1e4.err
for P in 316 -1000 0 -4000 -8000 -16000; do
  python tools/bench_beam.py --n 100000 --no-cpu --partitions $P 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('P request %6d -> level-1 partitions %5d  assemble+factor %6.2f ms  static solve %6.3f ms  tip rel err %.1e  modal %.3f s (%d it)' % ($P, d['partitions'], d['gpu_assemble_factor_ms'], d['gpu_static_solve_ms'], d['gpu_static']['rel_err_tip'], d['gpu_modal']['seconds'], d['gpu_modal']['iterations']))"
done > gpurun_out/r2_config4_two_level.txt
cat gpurun_out/r2_config4_two_level.txt
timeout 300 python tools/profile_beam_modal.py 2>&1 | grep -E "partitions|modal" > gpurun_out/r2_config4_modal_pieces.txt
cat gpurun_out/r2_config4_modal_pieces.txt
