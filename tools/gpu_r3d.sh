#!/bin/bash
# round-2 pass d (gpurun --gpus 2): multi-rank parity incl. reactions / supports, N=1 bench after the DRM rework, weak + strong at N=2,
# configs[2] (hex8 + PML3D) at N = 1 and 2
N=${1:-2}; TAG=${2:-r3d}
O=gpurun_out/$TAG
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout -k 10 420 $TR --master-port 29514 tests/multigpu_check.py > $O/check_n$N.log 2>&1; echo "check exit $?" >> $O/check_n$N.log
grep -E "multigpu|exit|rror" $O/check_n$N.log | tail -30
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err; tail -2 $O/bench_n1.err
timeout -k 10 900 $TR --master-port 29512 bench.py --gpus $N --steps 100 --warmup 5 > $O/bench_n$N.json 2> $O/bench_n$N.err; tail -2 $O/bench_n$N.err
python - <<PY
import json
for f in ("bench_n1","bench_n$N"):
    try:
        d=json.load(open("$O/%s.json"%f)); r=d["roofline"]
        print(f, "%.4g el/s"%d["value"], "ms/step %.4f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "kernel ms %.4f"%r["avg_launch_ms"], "frac %.3f"%r["frac"], d["kernel_ms"], d.get("parity_check"), d.get("replicas"), d.get("strong"), d["clocks"])
    except Exception as e: print(f, "failed", e)
PY
PML=${PML_N:-120}
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29516 tools/bench_pml_multi.py --size $PML --steps 20 > $O/bench_pml_n1.json 2> $O/bench_pml_n1.err
timeout -k 10 600 $TR --master-port 29515 tools/bench_pml_multi.py --size $PML --steps 20 > $O/bench_pml_n$N.json 2> $O/bench_pml_n$N.err
cat $O/bench_pml_n1.json $O/bench_pml_n$N.json; tail -2 $O/bench_pml_n$N.err
