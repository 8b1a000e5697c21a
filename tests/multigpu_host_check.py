#!/usr/bin/env python
"""Several-rank run of the C++ host driver (the reference's `mpirun -np N SeismoVLAB.exe -dir ... -file ...`):

  python tests/multigpu_host_check.py [N]          (needs N GPUs; default 2)

For each case: the global model is written as N per-rank JSON files in the reference's schema (global tags / dof numbers,
svl_b200.model.write_reference_partitions), `SeismoVLAB_gpu.exe -np N` runs them (one forked process per GPU, NCCL id
through the partition directory), and the per-rank NODE recorder files `<resp>.<rank>.out` are compared with the
single-domain oracle (first green run on 2 B200s: profiles/r3a_multigpu_host_check_n2.log)."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from oracle_lib import Oracle  # noqa: E402
from svl_b200 import model as M, partition as P  # noqa: E402

EXE = os.path.join(ROOT, "svl_b200", "SeismoVLAB_gpu.exe")
NE = {"kat444": (4, 4, 4), "drm_box": (6, 6, 5), "quad4_area": (8, 6), "lysmer_column": (3, 3, 6), "pml2d": None, "pml3d": None}


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    ok = True
    o = Oracle()
    for name, ne in NE.items():
        m = cases.CASES[name]()
        if ne is None:
            grid = P.proc_grid(n) if m.ndim == 3 else ((n, 1) if n <= 2 else (2, n // 2))
            ep = P.centroid_epart(m, grid)
        else:
            grid = P.proc_grid(n) if len(ne) == 3 else ((1, n) if n <= 2 else (2, n // 2))
            ep = P.block_epart(ne, grid)
        tmp = tempfile.mkdtemp(prefix="svlmh_")
        part = M.write_reference_partitions(m, ep, n, tmp, "Case", "Run", ndps=17)
        r = subprocess.run([EXE, "-np", str(n), "-dir", part, "-file", "Case.1.$.json"], capture_output=True, text=True, timeout=600)
        if r.returncode != 0:
            print(f"[multigpu host N={n}] {name:16s} driver exit {r.returncode}: {r.stdout[-400:]} {r.stderr[-400:]} FAIL", flush=True)
            ok = False
            continue
        ref, _ = o.run(m)
        # stitch the per-rank recorder files back into the global recorder order (columns: all dofs of each node)
        width = [int(m.node_ndof[q]) for q in m.rec_nodes]
        off = np.concatenate([[0], np.cumsum(width)]).astype(int)
        out = np.full_like(ref, np.nan)
        spread = 0.0
        for rk in range(n):
            fn = os.path.join(tmp, "Solution", "Run", f"disp.{rk}.out")
            if not os.path.exists(fn):
                continue
            with open(fn) as f:
                nn = int(f.readline().split()[0])
                tags = [int(f.readline().split()[0]) for _ in range(nn)]
            data = M.read_node_recorder(fn)
            c = 0
            for t in tags:
                i = list(m.rec_nodes).index(t - 1)
                blk = data[:, c:c + width[i]]
                seen = ~np.isnan(out[:, off[i]:off[i + 1]])
                if seen.all():
                    spread = max(spread, float(np.abs(out[:, off[i]:off[i + 1]] - blk).max()))
                out[:, off[i]:off[i + 1]] = blk
                c += width[i]
        err = cases.rel_err(out, ref) if not np.isnan(out).any() else float("nan")
        good = err < cases.TOL[name] and spread == 0.0
        ok &= bool(good)
        print(f"[multigpu host N={n}] {name:16s} grid={grid} rec err={err:.2e} replica spread={spread:.1e} {'OK' if good else 'FAIL'}", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
