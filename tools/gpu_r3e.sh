#!/bin/bash
# round-2 pass e (1 GPU): neighbour-list node classes (parity + bench), small-partition step anatomy (160^3: graph on / off, launch list)
TAG=${1:-r3e}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "unstructured or drm or golden or reaction or newmark or host_driver" > $O/pytest_sel.log 2>&1; echo "pytest exit $?" >> $O/pytest_sel.log
tail -6 $O/pytest_sel.log
timeout 900 python tools/bench_configs.py hexshuf hexshuf_gp hexjit hexgen --steps 20 > $O/configs_generic.jsonl 2> $O/configs_generic.err
python - <<PY
import json
for l in open("$O/configs_generic.jsonl"):
    d=json.loads(l); print(d["config"][:90], "| ms %.3f"%d["ms_per_step"], "el/s %.3g"%d["element_updates_per_s"], "frac %.3f"%d["frac_of_hbm_peak"], "nbr", d["nbr_nodes"], d["nbr_classes"], "gen", d["generic_elements"], "plan %.1fs"%d["plan_s"], {k:round(v,3) for k,v in d["kernel_ms"].items() if v})
PY
tail -3 $O/configs_generic.err
SVLGPU_GRAPH=1 timeout 300 python bench.py --n 160 --steps 400 --warmup 20 --no-cpu-baseline --no-verify > $O/bench_n160_graph1.json 2> $O/bench_n160_graph1.err
timeout 300 python bench.py --n 160 --steps 400 --warmup 20 --no-cpu-baseline --no-verify > $O/bench_n160.json 2> $O/bench_n160.err
python - <<PY
import json
for f in ("bench_n160","bench_n160_graph1"):
    try:
        d=json.load(open("$O/%s.json"%f)); r=d["roofline"]
        print(f, "%.4g el/s"%d["value"], "ms/step %.4f"%d["ms_per_step"], "e2e ms %.4f"%d["e2e"]["ms_per_step"], "launches/step", d["gpu_launches"]/d["steps"], d["kernel_ms"])
    except Exception as e: print(f, "failed", e)
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 60 --csv --log-file $O/launches_n160.csv \
    python bench.py --n 160 --steps 10 --warmup 3 --no-cpu-baseline --no-verify > $O/ncu_launch160.log 2>&1
