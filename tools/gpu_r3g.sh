#!/bin/bash
# round-2 pass g (1 GPU): full suite after the renumbering pre-pass / finite check, generic-mesh configs again
TAG=${1:-r3g}
O=gpurun_out/$TAG
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
tail -8 $O/pytest_gpu.log
timeout 900 python tools/bench_configs.py hexshuf hexgen hexjit quad4gen --steps 20 > $O/configs_generic.jsonl 2> $O/configs_generic.err
python - <<PY
import json
for l in open("$O/configs_generic.jsonl"):
    d=json.loads(l); print(d["config"][:90], "| ms %.4f"%d["ms_per_step"], "el/s %.3g"%d["element_updates_per_s"], "frac %.3f"%d["frac_of_hbm_peak"], "nbr", d["nbr_nodes"], d["nbr_classes"], "gen", d["generic_elements"], "plan %.1fs"%d["plan_s"], {k:round(v,3) for k,v in d["kernel_ms"].items() if v})
PY
tail -3 $O/configs_generic.err
