#!/bin/bash
# round-2 pass s (1 GPU): last sanity run -- host-call path tests + the contract bench with its default flags
TAG=${1:-r3s}
O=gpurun_out/$TAG
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "step_host or drm or kat or reaction" > $O/pytest_sel.log 2>&1; echo "pytest exit $?" >> $O/pytest_sel.log
tail -3 $O/pytest_sel.log
timeout 400 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > $O/bench_n320.json 2> $O/bench_n320.err; tail -2 $O/bench_n320.err
python - <<PY
import json
d=json.load(open("$O/bench_n320.json")); r=d["roofline"]
print("%.4g el/s"%d["value"], "ms/step %.4f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "kernel ms %.4f"%r["avg_launch_ms"], "frac %.3f"%r["frac"], d["kernel_ms"], (d.get("parity_check") or {}).get("max_rel_err_full_state_vs_oracle"), d["clocks"])
PY
