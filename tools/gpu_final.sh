#!/bin/bash
# 2-GPU round-end check: smoke, multi-GPU parity, weak bench at N=2, Newmark secondary bench (GPU 0)
TAG=${1:-final}
O=gpurun_out/$TAG
mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log; tail -3 $O/smoke.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/multigpu_check.py > $O/check_n2.log 2>&1; echo "check exit $?" >> $O/check_n2.log
grep -E "multigpu|exit" $O/check_n2.log | tail -9
timeout 600 $TR --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 5 > $O/bench_weak_n2.json 2> $O/bench_weak_n2.err
timeout 600 python tools/bench_configs.py newmark > $O/newmark.jsonl 2> $O/newmark.err
python - <<PY
import json
d=json.load(open("$O/bench_weak_n2.json")); print("weak N=2 %.4g el/s"%d["value"], "ms/step %.4f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"])
for l in open("$O/newmark.jsonl"):
    d=json.loads(l); print(d["config"][:70], "ms %.4f"%d["ms_per_step"], "el/s %.4g"%d["element_updates_per_s"], "cg it/step", d["pml_iterations_per_step"])
PY
tail -2 $O/newmark.err
# A/B: synchronous interface pass (the round-1 default before the comm-stream variant)
SVLGPU_HALO_SYNC=1 timeout 600 $TR --master-port 29514 bench.py --gpus 2 --steps 100 --warmup 5 > $O/bench_weak_n2_sync.json 2> $O/bench_weak_n2_sync.err
python - <<PY
import json
d=json.load(open("$O/bench_weak_n2_sync.json")); print("weak N=2 SYNC halo pass %.4g el/s"%d["value"], "ms/step %.4f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"])
PY
