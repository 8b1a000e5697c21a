#!/bin/bash
# A/B of the dominant-class stencil kernel variants on the contract workload (one box, back to back).
TAG=${1:-var}
O=gpurun_out/$TAG
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
run() { name=$1; shift; env "$@" timeout 400 python bench.py --steps 60 --warmup 5 --no-cpu-baseline > $O/bench_$name.json 2> $O/bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("$O/bench_$name.json")); print("$name", "%.4g el/s"%d["value"], "ms/step %.4f"%d["ms_per_step"], "stencil ms %.4f"%d["roofline"]["avg_launch_ms"], "e2e %.4g"%d["e2e"]["value"])
except Exception as e: print("$name failed", e)
PY
}
run v3 SVLGPU_STENCIL_V=3
run v4nobar SVLGPU_X=1
run v4bar SVLGPU_STENCIL_BAR=1
run v4r6nobar SVLGPU_STENCIL_R=6
run v4r6bar SVLGPU_STENCIL_R=6 SVLGPU_STENCIL_BAR=1
