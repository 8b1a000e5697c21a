#!/usr/bin/env python
"""Multi-GPU parity check, run under torchrun (one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tests/multigpu_check.py

Every rank builds the same global case, keeps its partition (svl_b200.partition.split_model), runs it on its GPU
with the NCCL interface exchange inside svlgpu_step, and rank 0 compares the stitched final displacement field and
the recorder histories with the single-domain oracle.  Prints one line per case and exits non-zero on failure."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from svl_b200 import capi, partition as P  # noqa: E402

NE = {"kat444": (4, 4, 4), "drm_box": (6, 6, 5), "quad4_area": (8, 6), "j2_column": (2, 2, 6),
      "hex8_layered_rayleigh": (4, 3, 6), "hex8_distorted": (3, 4, 5), "drm_area": (8, 6),
      "lysmer_column": (3, 3, 6),            # ZeroLength1D dashpots follow the rank of their soil node
      # soil box + PML layer (EQUAL ties, 9- / 5-dof PML nodes on the cuts): split by element centroid; the block solve
      # exchanges the shared unknowns and all-reduces its dot products
      "pml2d": None, "pml3d": None,
      # REACTION recorders on restrained nodes of the cut (partial F_int - F_ext summed over the ranks) and a moving support
      "reaction_box": (4, 3, 6), "support_column": (3, 3, 6),
      # the same with every recorded node inside rank 0's block: the other ranks record nothing but must still join the
      # reaction pass's interface exchange (option reaction_collective) -- this hung an 8-rank run before the option existed
      "reaction_box@left": (4, 3, 6),
      # the plane wave evaluated on the device (k_drm_pw_fused: DRM rows on the cut go into the exchanged partial forces),
      # the oracle reads the tabulated field of the same wave
      "drm_box@analytic": (6, 6, 5), "drm_area@analytic": (8, 6)}
RUNS = [(name, ne, "CENTRALDIFFERENCE") for name, ne in NE.items()]
# NewmarkBeta + Linear across ranks (interface sums inside the K operator, all-reduced dot products).  Both the PML and the
# Newmark cases were first seen green on 2 B200s in profiles/r3a_multigpu_check_pml_newmark_n2.log.
RUNS += [(name, NE[name], "NEWMARK") for name in cases.NEWMARK_CASES if NE.get(name) is not None]


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8))
    ok = True
    for name, ne, integrator in RUNS:
        # a fresh communicator per model keeps the check independent of call order
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        newmark = integrator == "NEWMARK"
        base = name.split("@")[0]
        reac = base in cases.REACTION_CASE_FUNCS
        m = cases.newmark_case(name) if newmark else (cases.REACTION_CASE_FUNCS[base]() if reac else cases.CASES[base]())
        m_dev = m
        if name.endswith("@analytic"):
            import copy
            m = cases.CASES[base]()
            m_dev = copy.copy(m); m_dev.drm = copy.copy(m.drm); m_dev.drm.field = None
        if name.endswith("@left"):
            m.rec_nodes = np.array([0, 1, 5, 6, 10], dtype=np.int32)       # x <= 1 of the 5 x 4 x 7 node lattice, all on the fixed base
        if ne is None:
            grid = P.proc_grid(world) if m.ndim == 3 else ((world, 1) if world <= 2 else (2, world // 2))
            subs = P.split_model(m_dev, P.centroid_epart(m, grid), world)
        elif len(ne) == 3:
            grid = P.proc_grid(world)
            if name == "j2_column" and world <= 6:
                grid = (1, 1, world)
            if reac:                                  # cut THROUGH the restrained base so that reactions need the exchange
                grid = {2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}.get(world, grid)
        else:
            grid = (1, world) if world <= 2 else (2, world // 2)
        if ne is not None:
            subs = P.split_model(m_dev, P.block_epart(ne, grid), world)
        s = subs[rank]
        d = capi.DeviceModel(s, device=local, comm=(rank, world, bytes(uid.cpu().numpy())), fields=(0, 3) if reac else (0,),
                             options={"integrator": 1.0} if newmark else None)
        d.step(1, m.nt, True)
        U = d.get_state(0)
        rec = d.read_recorder(0) if len(s.rec_nodes) else np.zeros((m.nt - 1, 0))
        recR = d.read_recorder(1) if reac and len(s.rec_nodes) else np.zeros((m.nt - 1, rec.shape[1]))
        nd = m.ndim
        gd = np.concatenate([np.arange(m.node_ptr[n], m.node_ptr[n + 1]) for n in s.global_nodes])    # PML nodes: 9 / 5 dofs
        rec_w = [int(s.node_ndof[n]) for n in s.rec_nodes]
        gathered = [None] * world
        dist.all_gather_object(gathered, (gd, U, s.rec_global, rec, d.counters(), rec_w, recR))
        d.close()
        if rank == 0:
            from oracle_lib import Oracle
            ref, Uref = Oracle().run(m, integrator=integrator)
            Ug = np.full(m.n_total, np.nan)
            spread = 0.0
            for gd_r, U_r, *_rest in gathered:
                seen = ~np.isnan(Ug[gd_r])
                if seen.any():
                    spread = max(spread, np.abs(Ug[gd_r][seen] - U_r[seen]).max())   # replicas must agree bit for bit
                Ug[gd_r] = U_r
            err_u = np.abs(Ug - Uref).max() / np.abs(Uref).max()
            # recorder columns back in the global recorder order
            cols = {}
            colsR = {}
            for _, _, rg, rc, _, rw, rR in gathered:
                off = np.concatenate([[0], np.cumsum(rw)]).astype(int)
                for i, n in enumerate(rg):
                    cols[int(n)] = rc[:, off[i]:off[i + 1]]
                    colsR[int(n)] = rR[:, off[i]:off[i + 1]]
            out = np.concatenate([cols[int(n)] for n in m.rec_nodes], axis=1)
            err_r = cases.rel_err(out, ref)
            if reac:                                  # reaction rows against the single-domain oracle
                refR, _ = Oracle().run(m, field=3)
                outR = np.concatenate([colsR[int(n)] for n in m.rec_nodes], axis=1)
                err_r = max(err_r, cases.rel_err(outR, refR))
            tol = cases.TOL_NEWMARK if newmark else (1e-9 if name.endswith("@analytic") else cases.TOL[base])
            good = err_u < tol and err_r < tol and spread == 0.0
            ok &= good
            c = gathered[0][4]
            print(f"[multigpu world={world}] {(name + ' newmark') if newmark else name:24s} grid={grid} max rel err U={err_u:.2e} rec={err_r:.2e} "
                  f"replica spread={spread:.1e} block_nodes(r0)={c['n_block_nodes']} generic(r0)={c['n_generic_elements']} "
                  f"{'OK' if good else 'FAIL'}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
