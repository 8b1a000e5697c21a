#!/bin/bash
# round-2 pass p (gpurun --gpus 2): final checks -- full GPU suite (incl. the 2-rank test), multi-rank parity with the analytic plane wave,
# contract bench at N = 1 (default flags) and N = 2
TAG=${1:-r3p}
O=gpurun_out/$TAG
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
timeout 600 python bench.py --steps 100 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; tail -2 $O/bench_n1.err
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; tail -2 $O/bench_n2.err
python - <<PY
import json
for f in ("bench_n1","bench_n2"):
    try:
        d=json.load(open("$O/%s.json"%f)); r=d["roofline"]
        print(f, "%.4g el/s"%d["value"], "ms/step %.4f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "kernel ms %.4f"%r["avg_launch_ms"], "frac %.3f"%r["frac"], "floor %.3f"%r["step_floor"]["frac"], d["kernel_ms"], (d.get("parity_check") or {}).get("max_rel_err_full_state_vs_oracle"), d.get("replicas"), (d.get("strong") or {}).get("ms_per_step"), d["clocks"])
    except Exception as e: print(f, "failed", e)
PY
