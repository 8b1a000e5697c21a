#!/usr/bin/env python
"""Writes the per-rank partition files of a synthetic BASELINE-like model (reference schema, big tables in binary sidecars)
so that the C++ host driver can run it the way the reference is run:

  python tools/gen_rank_files.py c3 --n 200 --np 8 --out /tmp/c3          # 200^3 hex8 + 10-cell PML3DHexa8 layer, 2x2x2
  svl_b200/SeismoVLAB_gpu.exe -np 8 -dir /tmp/c3/Partition -file 'C3.1.$.bin.json'

configs: c1 (n^3 soil column), c3 (n^3 half-space + PML layer on 5 faces), c5 (J2 column n x n x 4n).  Recorded nodes: 16 on
the free surface.  Nothing here touches a GPU."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from svl_b200 import model as M, partition as P  # noqa: E402

SOIL = [1.3e7, 0.3, 2000.0]
J2 = [2.9e7, 2.0e7, 2000.0, 1.0e7, 1.0, 1.0e4]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["c1", "c3", "c5"])
    ap.add_argument("--n", type=int, default=40)
    ap.add_argument("--np", type=int, default=2)
    ap.add_argument("--nt", type=int, default=200)
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    t0 = time.perf_counter()
    n = a.n
    if a.config == "c1":
        m = M.make_box_model((n, n, n), 1.0, mat=(M.ELASTIC3DLINEAR, SOIL), nt=a.nt)
        ep = P.block_epart((n, n, n), P.proc_grid(a.np))
        N1, top = n + 1, n
    elif a.config == "c5":
        m = M.make_box_model((n, n, 4 * n), 1.0, mat=(M.PLASTIC3DJ2, J2), nt=a.nt, load_dir=(3.0e6, 0.0, 1.0e6))
        ep = P.block_epart((n, n, 4 * n), (1, 1, a.np))
        N1, top = n + 1, 4 * n
    else:
        m = M.make_pml_model((n, n, n), 10, 1.0, soil=(M.ELASTIC3DLINEAR, SOIL), nt=a.nt)
        m.dt *= 0.5                                                   # 0.25 h / Vp: the CentralDifference + PML pair
        # PML elements weighted 250x: the block solve is what the step costs (9.0 ms vs 0.5 ms at 200^3), so the ranks must
        # share the PML shell evenly -- 5.40 vs 6.08 ms per step on 2 GPUs against the plain geometric split
        ep = P.weighted_epart(m, P.proc_grid(a.np))
        N1, top = n + 1, n
    ix = np.linspace(0, N1 - 1, 4).astype(int)
    m.rec_nodes = np.array([int(i + N1 * j + N1 * N1 * top) for j in ix for i in ix], dtype=np.int32)
    t1 = time.perf_counter()
    name = a.config.upper()
    if a.np == 1:
        part = M.write_reference_json(m, a.out, name, "Run", binary=True)
    else:
        part = M.write_reference_partitions(m, ep, a.np, a.out, name, "Run", binary=True)
    t2 = time.perf_counter()
    size = sum(os.path.getsize(os.path.join(part, f)) for f in os.listdir(part))
    print(f"{name}: {m.n_elem} elements, {m.n_total} dofs, {len(m.constraints)} ties; model {t1 - t0:.1f} s, files {t2 - t1:.1f} s, "
          f"{size / 1e6:.0f} MB in {part}")
    print(f"run: svl_b200/SeismoVLAB_gpu.exe {'-np %d ' % a.np if a.np > 1 else ''}-dir {part} -file '{name}.1.$.bin.json'")


if __name__ == "__main__":
    main()
