#!/bin/bash
# round-2 pass f (1 GPU): full GPU suite (incl. the reference executable with the GPU integrator linked in), generic-mesh configs,
# 128^3 parity check of the timed kernel variant
TAG=${1:-r3f}
O=gpurun_out/$TAG
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
tail -8 $O/pytest_gpu.log
timeout 900 python tools/bench_configs.py hexshuf hexgen quad4gen c1 --steps 20 > $O/configs_generic.jsonl 2> $O/configs_generic.err
python - <<PY
import json
for l in open("$O/configs_generic.jsonl"):
    d=json.loads(l); print(d["config"][:90], "| ms %.4f"%d["ms_per_step"], "el/s %.3g"%d["element_updates_per_s"], "frac %.3f"%d["frac_of_hbm_peak"], "nbr", d["nbr_nodes"], d["nbr_classes"], "gen", d["generic_elements"], "plan %.1fs"%d["plan_s"], {k:round(v,3) for k,v in d["kernel_ms"].items() if v})
PY
tail -3 $O/configs_generic.err
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --verify-n 128 > $O/bench_n320_verify128.json 2> $O/bench_verify.err
python -c "
import json; d=json.load(open('$O/bench_n320_verify128.json')); print(d['parity_check']); print('%.4g el/s'%d['value'], d['ms_per_step'], d['clocks'])"
