#!/bin/bash
# round-2 pass b (1 GPU): separable dominant-class kernel -- parity, A/B bench against v4, ncu
TAG=${1:-r3b}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants or separable or box_block or kat or internal_force or golden or drm" > $O/pytest_sel.log 2>&1; echo "pytest exit $?" >> $O/pytest_sel.log
tail -5 $O/pytest_sel.log
timeout 600 python bench.py --steps 100 --warmup 5 > $O/bench_n320.json 2> $O/bench_n320.err; tail -2 $O/bench_n320.err
SVLGPU_NO_SEP=1 timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-verify > $O/bench_n320_v4.json 2> $O/bench_n320_v4.err
python - <<PY
import json
for f in ("bench_n320","bench_n320_v4"):
    try:
        d=json.load(open("$O/%s.json"%f)); r=d["roofline"]
        print(f, "%.4g el/s"%d["value"], "ms/step %.4f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "kernel ms %.4f"%r["avg_launch_ms"], "frac %.3f"%r["frac"], "fp64 peak", r["fp64_peak_measured_TFLOPs"], "fp64 frac", r["fp64_fraction"], "copy", r["copy_GBs_measured_here"], d["kernel_ms"], d.get("parity_check"))
    except Exception as e: print(f, "failed", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stencil3_sep -s 6 -c 1 -o $O/stencil_sep_n320 -f \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-verify > $O/ncu_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 80 --csv --log-file $O/launches_n320.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-verify > $O/ncu_launch.log 2>&1
ls -la $O
