#!/usr/bin/env python
"""BASELINE.json configs[2] on N GPUs: n^3 lin3DHexa8 half-space + 10-cell PML3DHexa8 layer on 5 faces, split 2x2x2 (or
proc_grid(N)) by element centroid, one rank per GPU under torchrun.  Not the contract bench (bench.py); prints one JSON
line on rank 0.  Strong scaling: the global mesh is fixed, every rank keeps its partition.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \
      tools/bench_pml_multi.py --size 200 --steps 20

Parity of the multi-rank block solve: tests/multigpu_check.py (pml2d / pml3d cases).
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from svl_b200 import capi, model as M, partition as P  # noqa: E402

SOIL = [1.3e7, 0.3, 2000.0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", dest="n", type=int, default=200)   # (not --n: torchrun abbreviates it)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--part", default="weighted", choices=["weighted", "centroid"],
                    help="weighted: recursive bisection with PML elements weighted 250x (balances the block solve); centroid: plain geometric split")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    t0 = time.perf_counter()
    m = M.make_pml_model((a.n, a.n, a.n), 10, 1.0, soil=(M.ELASTIC3DLINEAR, SOIL), nt=4000)
    m.dt *= 0.5                                               # dt = 0.25 h / Vp as in tools/bench_configs.py c3
    grid = P.proc_grid(world)
    epart = (P.weighted_epart if a.part == "weighted" else P.centroid_epart)(m, grid) if world > 1 else None
    s = P.split_model(m, epart, world, ranks=(rank,))[rank] if world > 1 else m
    t_build = time.perf_counter() - t0
    comm = (rank, world, bytes(uid.cpu().numpy())) if world > 1 else None
    d = capi.DeviceModel(s, device=local, max_rows=a.warmup + a.steps + 8, comm=comm)
    d.step(1, 1 + a.warmup, True)
    dist.barrier(); torch.cuda.synchronize()
    d.step(1 + a.warmup, 1 + a.warmup + a.steps, True)
    torch.cuda.synchronize()
    c = d.counters()
    ms = torch.tensor([c["last_step_ms"] / a.steps], device="cuda", dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)                 # device-timed step loop, max over ranks
    U = d.get_state(0)
    fin = torch.tensor([1 if np.isfinite(U).all() else 0], device="cuda")
    dist.all_reduce(fin, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"config": f"configs[2] {a.n}^3 lin3DHexa8 + 10-cell PML3DHexa8 layer", "n_gpus": world,
                          "partition_grid": grid, "partitioner": a.part if world > 1 else None,
                          "pml_elements_per_rank": [int((np.isin(m.elem_kind, (3, 4)) & (epart == r)).sum()) for r in range(world)] if world > 1 else None,
                          "elements": int(m.n_elem), "dof": int(m.n_total),
                          "ms_per_step": float(ms.item()), "element_updates_per_s": m.n_elem / (float(ms.item()) * 1e-3),
                          "pml_iterations_per_step": (c["pml_iterations"] / c["pml_solves"]) if c["pml_solves"] else 0,
                          "pml_unknowns_rank0": c["n_pml_unknowns"], "launches_per_step_rank0": c["launches_per_step"],
                          "build_s": t_build, "finite": bool(fin.item()), "scaling": "strong"}), flush=True)
    d.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
