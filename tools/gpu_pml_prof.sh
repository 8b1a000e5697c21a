#!/bin/bash
TAG=${1:-pmlprof}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 300 --csv --log-file $O/launches_c3.csv \
    python tools/bench_configs.py c3 --scale 0.5 --steps 40 > $O/c3.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 300 --csv --log-file $O/launches_c2.csv \
    python tools/bench_configs.py c2 --steps 40 > $O/c2.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 100 --csv --log-file $O/launches_j2.csv \
    python tools/bench_configs.py j2 --steps 20 > $O/j2.log 2>&1
tail -2 $O/c3.log
