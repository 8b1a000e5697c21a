#!/bin/bash
# round-2 pass o (1 GPU): full suite + contract bench with the fused DRM kernel (A/B against the three-kernel form), e2e without the
# side-stream fork, neighbour-list kernel unroll A/B
TAG=${1:-r3o}
O=gpurun_out/$TAG
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
tail -6 $O/pytest_gpu.log
timeout 600 python bench.py --steps 100 --warmup 5 > $O/bench_n320.json 2> $O/bench_n320.err; tail -2 $O/bench_n320.err
SVLGPU_DRM_NO_FUSE=1 timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-verify > $O/bench_n320_nofuse.json 2> $O/bench_nofuse.err
python - <<PY
import json
for f in ("bench_n320","bench_n320_nofuse"):
    try:
        d=json.load(open("$O/%s.json"%f)); r=d["roofline"]
        print(f, "%.4g el/s"%d["value"], "ms/step %.4f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "kernel ms %.4f"%r["avg_launch_ms"], "frac %.3f"%r["frac"], "floor %.3f"%r["step_floor"]["frac"], d["kernel_ms"], (d.get("parity_check") or {}).get("max_rel_err_full_state_vs_oracle"), d["clocks"], "launches/step", d["gpu_launches"]/d["steps"])
    except Exception as e: print(f, "failed", e)
PY
timeout 300 python tools/bench_configs.py hexgen hexshuf --steps 20 > $O/configs_nbr_unr1.jsonl 2> $O/configs_unr1.err
SVLGPU_NBR_UNROLL=3 timeout 300 python tools/bench_configs.py hexgen hexshuf --steps 20 > $O/configs_nbr_unr3.jsonl 2> $O/configs_unr3.err
python - <<PY
import json
for f in ("configs_nbr_unr1","configs_nbr_unr3"):
    for l in open("$O/%s.jsonl"%f):
        d=json.loads(l); print(f, d["config"][:60], "| ms %.4f"%d["ms_per_step"], "el/s %.3g"%d["element_updates_per_s"])
PY
