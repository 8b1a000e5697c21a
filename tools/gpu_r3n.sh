#!/bin/bash
# round-2 pass n (1 GPU): evidence captures -- launch list of the final step at 320^3, ncu --set full of k_nbr_nodes, overlap A/B
TAG=${1:-r3n}
O=gpurun_out/$TAG
mkdir -p $O
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 70 --csv --log-file $O/launches_n320.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-verify > $O/ncu_launch.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_nbr_nodes -s 4 -c 1 -o $O/nbr_nodes_hexgen160 -f \
    python tools/bench_configs.py hexgen --steps 4 > $O/ncu_nbr.log 2>&1
SVLGPU_NO_OVERLAP=1 timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-verify > $O/bench_n320_nooverlap.json 2> $O/bench_nooverlap.err
python -c "
import json; d=json.load(open('$O/bench_n320_nooverlap.json')); print('NO_OVERLAP %.4g el/s'%d['value'], d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'])"
ls -la $O
