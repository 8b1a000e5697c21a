#!/bin/bash
# round-2 pass h (gpurun --gpus N): parity at N ranks, configs[2] (200^3 + PML3D) at N ranks, contract bench at N (weak + strong sub-record)
N=${1:-8}; TAG=${2:-r3h}; WHAT=${3:-all}
O=gpurun_out/$TAG
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$WHAT" = "all" ]; then
timeout -k 10 400 $TR --master-port 29514 tests/multigpu_check.py > $O/check_n$N.log 2>&1; echo "check exit $?" >> $O/check_n$N.log
grep -E "multigpu|exit|rror" $O/check_n$N.log | tail -30
fi
timeout -k 10 700 $TR --master-port 29515 tools/bench_pml_multi.py --size ${PML_N:-200} --steps 20 > $O/bench_pml_n$N.json 2> $O/bench_pml_n$N.err
cat $O/bench_pml_n$N.json; tail -2 $O/bench_pml_n$N.err
if [ "$WHAT" = "all" ]; then
timeout -k 10 900 $TR --master-port 29512 bench.py --gpus $N --steps 100 --warmup 5 > $O/bench_n$N.json 2> $O/bench_n$N.err; tail -2 $O/bench_n$N.err
python - <<PY
import json
try:
    d=json.load(open("$O/bench_n$N.json")); r=d["roofline"]
    print("N=$N weak %.4g el/s"%d["value"], "ms/step %.4f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "kernel ms %.4f"%r["avg_launch_ms"], d["kernel_ms"], d.get("replicas"), d.get("strong"), d["clocks"])
except Exception as e: print("failed", e)
PY
fi
