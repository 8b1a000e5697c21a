#!/bin/bash
TAG=${1:-pml}
O=gpurun_out/$TAG
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -k "pml or golden or host_driver" > $O/pytest_pml.log 2>&1; tail -3 $O/pytest_pml.log
timeout 900 python tools/bench_configs.py c2 c3 > $O/configs.jsonl 2> $O/configs.err
python - <<PY
import json
for l in open("$O/configs.jsonl"):
    d=json.loads(l); print(d["config"][:60], "ms %.4f"%d["ms_per_step"], "el/s %.4g"%d["element_updates_per_s"], "pml it", d["pml_iterations_per_step"], {k:round(v,4) for k,v in d["kernel_ms"].items() if v})
PY
tail -3 $O/configs.err
