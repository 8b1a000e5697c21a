#!/bin/bash
# One GPU pass: parity tests, contract bench, ncu launch list + full capture of the dominant kernel, secondary configs.
# Usage (from repo root, on the GPU box): bash tools/gpu_pass.sh <tag>
TAG=${1:-pass}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,memory.total --format=csv > $O/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 600 python bench.py --steps 100 --warmup 5 > $O/bench_n320.json 2> $O/bench_n320.err
timeout 300 python bench.py --n 200 --steps 100 --warmup 5 --no-cpu-baseline > $O/bench_n200.json 2> $O/bench_n200.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 80 --csv --log-file $O/launches_n320.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $O/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stencil3_v4 -s 6 -c 1 -o $O/stencil_v4_n320 -f \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $O/ncu_full.log 2>&1
timeout 900 python tools/bench_configs.py > $O/configs.jsonl 2> $O/configs.err
tail -3 $O/pytest_gpu.log; cat $O/bench_n320.json
