#!/bin/bash
# ncu capture of the static-only tile kernel (one load case), full set + source page
set -x
mkdir -p gpurun_out
timeout 300 python tools/static_tile_probe.py 100000 1 > gpurun_out/tile_probe.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_static_tile -c 1 -o gpurun_out/tile_full -f python tools/static_tile_probe.py 100000 1 > gpurun_out/tile_ncu.log 2>&1
ls -la gpurun_out/
