#!/bin/bash
# multi-GPU pass (run with gpurun --gpus N): parity of the interface exchange + contract bench at N GPUs
N=${1:-8}; TAG=${2:-multi}
O=gpurun_out/$TAG
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/multigpu_check.py > $O/check_n$N.log 2>&1; echo "check exit $?" >> $O/check_n$N.log
grep -E "multigpu|exit" $O/check_n$N.log | tail -20
# the C++ host driver over per-rank JSON files in the reference's schema (SeismoVLAB_gpu.exe -np N)
timeout 600 python tests/multigpu_host_check.py $N > $O/check_host_n$N.log 2>&1; echo "check(host) exit $?" >> $O/check_host_n$N.log
grep -E "multigpu host|exit" $O/check_host_n$N.log | tail -8
timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps 100 --warmup 5 > $O/bench_weak_n$N.json 2> $O/bench_weak_n$N.err
timeout 600 $TR --master-port 29513 bench.py --gpus $N --steps 200 --warmup 10 --scaling strong > $O/bench_strong_n$N.json 2> $O/bench_strong_n$N.err
python - <<PY
import json
for k in ("weak","strong"):
    try:
        d=json.load(open("$O/bench_%s_n$N.json"%k)); print(k, "N=$N %.4g el/s"%d["value"], "ms/step %.4f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], d["config"]["partition_grid"], "if nodes", d["config"]["interface_nodes_rank0"])
    except Exception as e: print(k, "failed", e)
PY
tail -3 $O/bench_weak_n$N.err
# BASELINE configs[2] across the GPUs (only meaningful once the pml lines above are OK)
if grep -q "check exit 0" $O/check_n$N.log; then
  timeout 900 $TR --master-port 29515 tools/bench_pml_multi.py --size ${PML_N:-200} --steps 20 > $O/bench_pml_n$N.json 2> $O/bench_pml_n$N.err
  cat $O/bench_pml_n$N.json
fi
