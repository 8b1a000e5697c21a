#!/bin/bash
# quick pass: parity tests + contract bench + chosen secondary configs
TAG=${1:-quick}; shift
O=gpurun_out/$TAG
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > $O/bench_n320.json 2> $O/bench_n320.err
python - <<PY
import json
d=json.load(open("$O/bench_n320.json")); print("n320 %.4g el/s"%d["value"], "ms/step %.4f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], {k:round(v,4) for k,v in d["kernel_ms"].items()})
PY
timeout 900 python tools/bench_configs.py "$@" > $O/configs.jsonl 2> $O/configs.err
python - <<PY
import json
for l in open("$O/configs.jsonl"):
    d=json.loads(l); print(d["config"][:60], "ms %.4f"%d["ms_per_step"], "el/s %.4g"%d["element_updates_per_s"], "pml it", d["pml_iterations_per_step"], {k:round(v,4) for k,v in d["kernel_ms"].items() if v})
PY
tail -3 $O/configs.err
