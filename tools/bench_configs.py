#!/usr/bin/env python
"""Secondary measurements: every kernel family of the path on a BASELINE.json-config-like workload, each against
its own SURVEY.md 8(d) figure.  Not the contract bench (that is bench.py); prints one JSON line per config.

  python tools/bench_configs.py [c1 c2 c3 c5 quad4 j2small ...] [--steps K]
"""
import argparse
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from svl_b200 import model as M  # noqa: E402
from svl_b200.capi import DeviceModel  # noqa: E402

SOIL = [1.3e7, 0.3, 2000.0]
J2 = [2.9e7, 2.0e7, 2000.0, 1.0e7, 1.0, 1.0e4]        # fixture F07
PEAK = 6650.0


def peak():
    """the same MEASURED_PEAKS.json reading as bench.py (any key layout), fallback 6650 GB/s"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("svl_bench", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(b)
        return b.measured_peaks()[0]
    except Exception:
        return PEAK


def run(name, m, bytes_per_elem, steps, warm=5, note="", options=None):
    t0 = time.perf_counter()
    d = DeviceModel(m, max_rows=warm + 2 * steps + 8, options=options)
    t_plan = time.perf_counter() - t0
    d.step(1, 1 + warm, True)
    d.step(1 + warm, 1 + warm + steps, True)
    c = d.counters()
    ms = c["last_step_ms"] / steps
    d.set_kernel_timing(True)
    d.step(1 + warm + steps, 1 + warm + steps + min(steps, 10), True)
    nts = min(steps, 10)
    kt = {}
    for i, k in enumerate(("stencil_dom", "gauss_elements", "gather_nodes", "point_loads", "stencil_shell", "drm",
                           "pml_elem_products", "pml_gathers", "pml_vector_updates")):
        avg, n = d.kernel_time(i)
        kt[k] = avg * n / nts                      # ms per step spent in this kernel family
    c2 = d.counters()
    rate = m.n_elem / (ms * 1e-3)
    U = d.get_state(0)
    line = {"config": name, "elements": m.n_elem, "dof": m.n_total, "ms_per_step": ms, "element_updates_per_s": rate,
            "algorithmic_bytes_per_element_update": bytes_per_elem,
            "algorithmic_GBs": rate * bytes_per_elem / 1e9, "frac_of_hbm_peak": rate * bytes_per_elem / 1e9 / peak(),
            "kernel_ms": kt, "block_nodes": c["n_block_nodes"], "generic_elements": c["n_generic_elements"],
            "nbr_nodes": c2["n_nbr_nodes"], "nbr_classes": c2["n_nbr_classes"],
            "pml_elements": c2["n_pml_elements"], "pml_unknowns": c2["n_pml_unknowns"],
            "pml_iterations_per_step": (c2["pml_iterations"] / c2["pml_solves"]) if c2["pml_solves"] else 0,
            "launches_per_step": c2["launches_per_step"], "plan_s": t_plan, "finite": bool(np.isfinite(U).all()),
            "peak_abs_u": float(np.abs(U).max()), "note": note}
    print(json.dumps(line), flush=True)
    d.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="*", default=["c1", "quad4", "j2", "c2", "c3"])
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--scale", type=float, default=1.0, help="linear scale of the config sizes")
    a = ap.parse_args()
    S = a.scale
    for w in a.which:
        if w == "c1":     # configs[0]: 20^3 soil column, the reference's own CPU-runnable case
            m = M.make_box_model((20, 20, 20), 1.0, mat=(M.ELASTIC3DLINEAR, SOIL), nt=4000)
            run("configs[0] 20^3 lin3DHexa8 (launch-latency bound)", m, 140.0, 10 * a.steps)
        elif w == "quad4":  # configs[1] without the PML: 2000 x 1000 quad4 lattice
            ne = (int(2000 * S), int(1000 * S))
            m = M.make_area_model(ne, 1.0, nt=4000)
            run(f"{ne[0]}x{ne[1]} lin2DQuad4 + Elastic2DPlaneStrain lattice", m, 92.0, 5 * a.steps)
        elif w == "quad4gen":
            ne = (int(2000 * S), int(1000 * S))
            m = M.make_area_model(ne, 1.0, nt=4000)
            m.blocks = []
            run(f"{ne[0]}x{ne[1]} lin2DQuad4 Gauss-point path (no lattice)", m, 92.0, a.steps)
        elif w == "j2":     # configs[4]: 100 x 100 x 400 Plastic3DJ2
            ne = (int(100 * S), int(100 * S), int(400 * S))
            m = M.make_box_model(ne, 1.0, mat=(M.PLASTIC3DJ2, J2), nt=4000, load_dir=(3.0e6, 0.0, 1.0e6))
            # base shear through point loads on the whole free surface so that Gauss points yield
            N1 = (ne[0] + 1) * (ne[1] + 1)
            top = np.arange(N1 * ne[2], N1 * (ne[2] + 1), dtype=np.int32)
            m.point_loads = [M.PointLoad(top, np.array([4.0e4, 0.0, 0.0]), np.ones(1))]
            run(f"configs[4] {ne[0]}x{ne[1]}x{ne[2]} lin3DHexa8 + Plastic3DJ2", m, 1804.0, a.steps)
        elif w == "hexgen":
            ne = (int(160 * S),) * 3
            m = M.make_box_model(ne, 1.0, mat=(M.ELASTIC3DLINEAR, SOIL), nt=4000)
            m.blocks = []
            run(f"{ne[0]}^3 lin3DHexa8 elastic, Gauss-point path (no lattice)", m, 140.0, a.steps)
        elif w == "hexshuf":
            # unstructured numbering: the same cells with node ids and element order permuted at random, two materials in
            # layers -> no lattice block; pre-summed rows + neighbour lists (k_nbr_nodes) where the row repeats
            ne = (int(160 * S),) * 3
            mats = [(M.ELASTIC3DLINEAR, SOIL), (M.ELASTIC3DLINEAR, [5.0e7, 0.25, 2200.0])]
            m = M.make_box_model(ne, 1.0, nt=4000, layers=mats)
            m.dt *= 0.5
            m = M.shuffle_numbering(m, 20260117)
            run(f"{ne[0]}^3 lin3DHexa8, 2 materials, random node / element numbering (neighbour-list node classes)", m, 140.0, a.steps)
        elif w == "hexshuf_gp":
            ne = (int(160 * S),) * 3
            mats = [(M.ELASTIC3DLINEAR, SOIL), (M.ELASTIC3DLINEAR, [5.0e7, 0.25, 2200.0])]
            m = M.make_box_model(ne, 1.0, nt=4000, layers=mats)
            m.dt *= 0.5
            m = M.shuffle_numbering(m, 20260117)
            run(f"{ne[0]}^3 lin3DHexa8, 2 materials, random numbering, Gauss-point path (nbr_classes = 0)", m, 140.0, a.steps,
                options={"nbr_classes": 0.0})
        elif w == "hexjit":
            # geometry jitter +-0.1 h with default_rng(20260117): every element its own class -> Gauss-point path
            ne = (int(160 * S),) * 3
            m = M.make_box_model(ne, 1.0, mat=(M.ELASTIC3DLINEAR, SOIL), nt=4000, jitter=0.1, seed=20260117)
            m.dt *= 0.5
            m.blocks = []
            run(f"{ne[0]}^3 lin3DHexa8, jittered geometry (+-0.1 h): Gauss-point path from the coordinates", m, 140.0, a.steps)
        elif w == "newmark":   # the implicit NewmarkBeta + Linear step at 4x the explicit time step
            n = int(160 * S)
            m = M.make_box_model((n, n, n), 1.0, mat=(M.ELASTIC3DLINEAR, SOIL), nt=4000)
            m.dt *= 4.0
            run(f"NewmarkBeta + Linear, {n}^3 lin3DHexa8 lattice, dt = 2 h/Vp (matrix-free CG, rtol 1e-13)", m, 140.0,
                max(4, a.steps // 4), note="one step = 1 + CG-iterations applications of the K operator; pml_iterations_per_step counts CG iterations",
                options={"integrator": 1.0})
        elif w == "c2":     # configs[1]: quad4 half-space + PML2DQuad4 layer, left / right / bottom
            ne = (int(2000 * S), int(1000 * S))
            m = M.make_pml_model(ne, 10, 1.0, soil=(M.ELASTIC2DPLANESTRAIN, SOIL), nt=4000)
            m.dt *= 0.5
            run(f"configs[1] {ne[0]}x{ne[1]} lin2DQuad4 + 10-cell PML2DQuad4 layer", m, 92.0, max(4, a.steps // 4),
                note="dt = 0.25 h/Vp; every step solves the non-diagonal PML block")
        elif w == "c3":     # configs[2]: hex8 half-space + PML3DHexa8 layer on 5 faces
            n = int(200 * S)
            m = M.make_pml_model((n, n, n), 10, 1.0, soil=(M.ELASTIC3DLINEAR, SOIL), nt=4000)
            m.dt *= 0.5
            run(f"configs[2] {n}^3 lin3DHexa8 + 10-cell PML3DHexa8 layer", m, 140.0, max(2, a.steps // 10),
                note="dt = 0.25 h/Vp; every step solves the non-diagonal PML block")


if __name__ == "__main__":
    main()
