#!/bin/bash
# round-2 pass c (1 GPU): full GPU suite (reactions, support motion, separable kernel) + contract bench with parity check
TAG=${1:-r3c}
O=gpurun_out/$TAG
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
tail -15 $O/pytest_gpu.log
timeout 900 python bench.py --steps 100 --warmup 5 > $O/bench_n320.json 2> $O/bench_n320.err; tail -2 $O/bench_n320.err
python - <<PY
import json
try:
    d=json.load(open("$O/bench_n320.json")); r=d["roofline"]
    print("%.4g el/s"%d["value"], "ms/step %.4f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "kernel ms %.4f"%r["avg_launch_ms"], "frac %.3f"%r["frac"], "fp64 frac %.3f"%r["fp64_fraction"], d["kernel_ms"], d.get("parity_check"), d["clocks"])
except Exception as e: print("failed", e)
PY
