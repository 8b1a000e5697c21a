#!/bin/bash
# round-2 pass r (1 GPU): DRM force + application in one kernel (k_drm_pw_apply) -- parity subset, contract bench, A/B
TAG=${1:-r3r}
O=gpurun_out/$TAG
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "drm or golden or reaction or kat or step_host or graph" > $O/pytest_sel.log 2>&1; echo "pytest exit $?" >> $O/pytest_sel.log
tail -4 $O/pytest_sel.log
timeout 400 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > $O/bench_n320.json 2> $O/bench_n320.err; tail -2 $O/bench_n320.err
SVLGPU_DRM_NO_INLINE=1 timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-verify > $O/bench_n320_noinline.json 2> $O/bench_noinline.err
python - <<PY
import json
for f in ("bench_n320","bench_n320_noinline"):
    try:
        d=json.load(open("$O/%s.json"%f)); r=d["roofline"]
        print(f, "%.4g el/s"%d["value"], "ms/step %.4f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "kernel ms %.4f"%r["avg_launch_ms"], "frac %.3f"%r["frac"], "floor %.3f"%r["step_floor"]["frac"], d["kernel_ms"], (d.get("parity_check") or {}).get("max_rel_err_full_state_vs_oracle"), d["clocks"], "launches/step", d["gpu_launches"]/d["steps"])
    except Exception as e: print(f, "failed", e)
PY
