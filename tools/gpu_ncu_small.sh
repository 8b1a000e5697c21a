#!/bin/bash
TAG=${1:-ncusmall}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_drm|k_stencil3_shell' -s 12 -c 4 -o $O/small_n320 -f \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $O/ncu.log 2>&1
tail -3 $O/ncu.log | cut -c1-300
