#!/bin/bash
# A/B of scheduling options on the contract workload
TAG=${1:-ab}
O=gpurun_out/$TAG
mkdir -p $O
run() { name=$1; shift; env "$@" timeout 400 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > $O/bench_$name.json 2> $O/bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("$O/bench_$name.json")); print("$name", "%.4g el/s"%d["value"], "ms/step %.4f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "e2e ms %.4f"%d["e2e"]["ms_per_step"], {k:round(v,4) for k,v in d["kernel_ms"].items() if v})
except Exception as e: print("$name failed", e)
PY
}
run prio SVLGPU_X=1
run noprio SVLGPU_NO_PRIO=1
run prio_lowreg SVLGPU_SHELL_LOWREG=1
