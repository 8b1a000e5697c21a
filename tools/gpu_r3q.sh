#!/bin/bash
# round-2 pass q (1 GPU): z-chunk length of the separable stencil kernel vs wave quantisation (2800 CTAs = 6.3 waves of 444 today)
TAG=${1:-r3q}
O=gpurun_out/$TAG
mkdir -p $O
for KZ in 29 16 36; do
  SVLGPU_STENCIL_KZ=$KZ timeout 200 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-verify > $O/bench_kz$KZ.json 2> $O/bench_kz$KZ.err
done
python - <<PY
import json
for kz in (29,16,36):
    try:
        d=json.load(open("$O/bench_kz%d.json"%kz)); r=d["roofline"]
        print("KZ", kz, "%.4g el/s"%d["value"], "ms/step %.4f"%d["ms_per_step"], "kernel ms %.4f"%r["avg_launch_ms"], "frac %.3f"%r["frac"])
    except Exception as e: print(kz, "failed", e)
PY
