#!/bin/bash
# round-2 pass m (1 GPU): full suite + contract bench + PML 120^3 after the extrapolated starting guess / shell slot masks
TAG=${1:-r3m}
O=gpurun_out/$TAG
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
tail -6 $O/pytest_gpu.log
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > $O/bench_n320.json 2> $O/bench_n320.err; tail -2 $O/bench_n320.err
python - <<PY
import json
try:
    d=json.load(open("$O/bench_n320.json")); r=d["roofline"]
    print("%.4g el/s"%d["value"], "ms/step %.4f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "kernel ms %.4f"%r["avg_launch_ms"], "frac %.3f"%r["frac"], d["kernel_ms"], d["parity_check"]["max_rel_err_full_state_vs_oracle"], d["clocks"])
except Exception as e: print("failed", e)
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29516 tools/bench_pml_multi.py --size 120 --steps 20 > $O/bench_pml120.json 2> $O/bench_pml120.err
SVLGPU_PML_NO_EXTRAP=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29517 tools/bench_pml_multi.py --size 120 --steps 20 > $O/bench_pml120_noextrap.json 2> $O/bench_pml120_noextrap.err
grep -h config $O/bench_pml120.json $O/bench_pml120_noextrap.json | cut -c1-600
