#!/usr/bin/env python
"""bench.py -- element-updates/s of the FP64 explicit step (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W          our arm (CUDA kernels through the C ABI)
  python bench.py --impl reference --steps K --warmup W  the reference's own CPU code (oracle/_ref)

Workload at N=1 (config.workload): BASELINE.json configs[3]-like 3-D elastic half-space,
n^3 lin3DHexa8 + Elastic3DLinear (default n=320: 321^3 nodes ~ 10^8 DOF, the mesh the north_star
target is quoted on; it fits one B200), lumped mass, CentralDifference, bottom fixed, vertically
incident SV Ricker plane wave injected through a one-element DRM layer 5 cells inside the boundary
(evaluated on the device), one host-fed Ricker point load, 16 recorded nodes.
A "step" is one CentralDifference step of the whole mesh.  Prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HEX8_BYTES = 140.0      # algorithmic bytes per lin3DHexa8 element-update
HEX8_FLOPS_STENCIL = 504.0   # SURVEY 8(d) "minimum-known formulation" (27 x 3x3 node stencil + update)
# FP64 instructions the dominant kernel EXECUTES per node (SASS of the built kernel, profiles/r3b_sass_k_stencil3_sep.md):
#   k_stencil3_sep<4,4>: 48 DFMA + 27.4 DADD + 3 DMUL (x pass on R+2 rows per R nodes, y pass, z scatter, update)
#   k_stencil3_v4<4,4,SYM>: 153 DFMA + 9 (update)
KERNELS = {"sep": {"name": "k_stencil3_sep", "fp64_instr": 78.375, "flop": 2 * 48 + 27.375 + 3, "bytes": 72},
           "v4": {"name": "k_stencil3_v4", "fp64_instr": 162.0, "flop": 2 * 153 + 15, "bytes": 73}}
MAT = [1.3e7, 0.3, 2000.0]   # fixture J05 soil


def pick_hbm_peak(doc):
    """HBM GB/s out of a MEASURED_PEAKS.json document whose exact schema this repo does not control: every numeric leaf whose
    key path mentions hbm / bandwidth / copy / dram is a candidate; the dominant kernel is timed inside a long step, so a
    'sustained' figure wins over a 'burst' one; values below 100 are taken as TB/s.  Returns (GB/s, key path) or None."""
    leaves = []

    def walk(x, path):
        if isinstance(x, dict):
            for k, v in x.items():
                walk(v, path + [str(k)])
        elif isinstance(x, (list, tuple)):
            for i, v in enumerate(x):
                walk(v, path + [str(i)])
        elif isinstance(x, (int, float)) and not isinstance(x, bool):
            leaves.append((".".join(path), float(x)))

    walk(doc, [])
    cand = [(k, v) for k, v in leaves if v > 0 and any(t in k.lower() for t in ("hbm", "bandwidth", "copy", "dram"))
            and not any(t in k.lower() for t in ("flop", "bf16", "fp16", "tf"))]
    if not cand:
        return None
    # keys that name HBM outright win over generic "*bandwidth*" ones; then sustained > unlabelled > burst
    cand.sort(key=lambda kv: (0 if "hbm" in kv[0].lower() else 1,
                              0 if "sustain" in kv[0].lower() else 1 if "burst" not in kv[0].lower() else 2))
    for k, v in cand:
        if v < 100.0:
            v *= 1000.0                               # TB/s
        if 1000.0 <= v <= 20000.0:
            return v, k
    return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            got = pick_hbm_peak(json.load(f))
        if got:
            return got[0], f"measured (MEASURED_PEAKS.json {got[1]})"
    except (OSError, KeyError, TypeError, ValueError):
        pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.active = index, [], False, False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                if self.active:                       # only samples taken DURING the timed regions count
                    self.samples.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if len(s) >= 6 and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) >= 6 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples if len(s) >= 6 for i in range(4) if s[2 + i] == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def build_workload(n, nt, rank=0, world=1, strong=False):
    """Rank `rank`'s block of the global box: world == 1 -> the whole n^3 mesh; world > 1 -> weak scaling
    (every rank n^3 elements, global mesh = proc_grid(world) * n) or strong scaling (global n^3 split)."""
    from svl_b200 import model as M
    from svl_b200 import partition as P
    grid = P.proc_grid(world)
    if strong:
        if any(n % g for g in grid):
            raise SystemExit(f"--scaling strong needs n divisible by the process grid {grid}")
        nl = tuple(n // g for g in grid)
    else:
        nl = (n, n, n)
    G = [nl[a] * grid[a] for a in range(3)]                    # global elements per axis
    if world == 1:
        m = M.make_box_model(nl, 1.0, mat=(M.ELASTIC3DLINEAR, MAT), nt=nt, fix="bottom")
        m.halos = {}
        rpos = (0, 0, 0)
    else:
        m = P.local_box(nl, grid, rank, 1.0, mat=(M.ELASTIC3DLINEAR, MAT), nt=nt)
        rpos = m.grid_pos
    mu = MAT[0] / (2 * (1 + MAT[1]))
    vs = math.sqrt(mu / MAT[2])
    f0 = vs / (10.0 * 1.0) / 4.0          # >= 10 cells per S wavelength with margin
    pw = dict(dir=[0.0, 0.0, 1.0], pol=[1.0, 0.0, 0.0], xref=[0.0, 0.0, 0.0], c=vs, f0=f0, t0=1.2 / f0, amp=1e-3)
    if min(G) >= 16:
        M.add_drm_box(m, x0=[G[0] / 2, G[1] / 2, G[2]], xl=[G[0] / 2 - 5.5, G[1] / 2 - 5.5, G[2] - 5.5], planewave=pw)
        if len(m.drm.elems) == 0:
            m.drm = None
    # host-fed Ricker point load at the centre of the free surface: handed to the lowest rank that holds the node
    gl = (G[0] // 2, G[1] // 2, G[2])
    holders = []
    for r in range(world):
        rx, ry, rz = r % grid[0], (r // grid[0]) % grid[1], r // (grid[0] * grid[1])
        if all(rp * nl[a] <= gl[a] <= (rp + 1) * nl[a] for a, rp in enumerate((rx, ry, rz))):
            holders.append(r)
    N1 = [c + 1 for c in nl]
    if rank == holders[0]:
        li = [gl[a] - rpos[a] * nl[a] for a in range(3)]
        node = li[0] + N1[0] * li[1] + N1[0] * N1[1] * li[2]
        m.point_loads = [M.PointLoad(np.array([node], dtype=np.int32), np.array([0.0, 0.0, 1.0]),
                                     1e4 * M.ricker(nt, m.dt, f0, 1.2 / f0))]
    else:
        m.point_loads = []
    ix = np.linspace(0, N1[0] - 1, 4).astype(int); iy = np.linspace(0, N1[1] - 1, 4).astype(int)
    rec = [int(i + N1[0] * j + N1[0] * N1[1] * nl[2]) for j in iy for i in ix]      # 16 nodes of the block's top layer
    m.rec_nodes = np.array(rec, dtype=np.int32)
    m.global_elems_per_axis = G
    return m


def cpu_baseline_port(budget_s=12.0):
    """The CPU restatement (oracle/, OpenMP over elements) on a bounded sample of the same workload."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Oracle
    from svl_b200 import model as M
    o = Oracle()
    cores = os.cpu_count() or 1
    n, nt = 80, 12
    m = M.make_box_model((n, n, n), 1.0, mat=(M.ELASTIC3DLINEAR, MAT), nt=nt)
    t0 = time.perf_counter(); o.run(m, nt=2, nthreads=cores); t_setup = time.perf_counter() - t0
    t0 = time.perf_counter(); o.run(m, nt=nt, nthreads=cores); t_full = time.perf_counter() - t0
    dt_steps = max(t_full - t_setup, 1e-9)
    rate = m.n_elem * (nt - 2) / dt_steps
    return {"value": rate, "unit": "element-updates/s", "cores": cores, "kind": "port",
            "sample": f"{n}^3 lin3DHexa8 box, {nt - 2} CentralDifference steps (setup subtracted), oracle/svl_oracle.c "
                      f"with OpenMP over elements on {cores} threads"}


def run_reference_arm(args, emit):
    """Times the UNMODIFIED reference executable (oracle/_ref/SeismoVLAB.exe, built from the reference's
    own sources against oracle/shim) on the host: same element / material / integrator, a bounded
    sample of the mesh.  Falls back to the oracle port when the executable did not travel."""
    from svl_b200 import model as M
    exe = os.path.join(ROOT, "oracle", "_ref", "SeismoVLAB.exe")
    K, W = args.steps, args.warmup
    if not os.path.exists(exe):
        cb = cpu_baseline_port()
        line = {"metric": "element-updates/sec (FP64 explicit step)", "value": cb["value"], "unit": "element-updates/s",
                "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": None, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
                "config": {"workload": "oracle port (reference executable absent on this box)"},
                "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "element-updates/s",
                                            "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return
    n = args.ref_n
    S = args.ref_steps
    P = os.cpu_count() or 1
    import shutil
    tmp = tempfile.mkdtemp(prefix="svlref_")

    def prepare(nt):
        """one partition directory per process: the reference has no threads, its parallel
        path is one MPI rank per partition; MPI/MUMPS are not in this image, so the P ranks run as P
        independent single-partition processes (interface coupling dropped: an upper bound of its throughput)"""
        m = build_workload(n, nt)
        if m.drm is not None:
            from svl_b200.model import add_drm_box
            G = m.global_elems_per_axis
            add_drm_box(m, x0=[G[0] / 2, G[1] / 2, G[2]], xl=[G[0] / 2 - 5.5, G[1] / 2 - 5.5, G[2] - 5.5],
                        planewave=m.drm.planewave, tabulate_nt=nt)
        base = os.path.join(tmp, f"nt{nt}_0")
        M.write_reference_json(m, base, "Bench", "Bench")
        dirs = [base]
        for q in range(1, P):
            dq = os.path.join(tmp, f"nt{nt}_{q}")
            if os.path.exists(dq):
                shutil.rmtree(dq)
            shutil.copytree(base, dq)
            dirs.append(dq)
        return dirs, m.n_elem

    def run_all(dirs):
        t0 = time.perf_counter()
        procs = [subprocess.Popen([exe, "-dir", os.path.join(dq, "Partition"), "-file", "Bench.1.$.json"],
                                  stdout=subprocess.DEVNULL) for dq in dirs]
        for pr in procs:
            if pr.wait() != 0:
                raise SystemExit("reference executable failed")
        return time.perf_counter() - t0

    dirs_lo, nelem = prepare(2)                  # parse + Initialize + 1 step
    dirs_hi, _ = prepare(2 + S)                  # ... + S more steps
    times, his = [], []
    for it in range(W + K):
        t_lo = run_all(dirs_lo)
        t_hi = run_all(dirs_hi)
        if it >= W:
            times.append(t_hi - t_lo)
            his.append(t_hi)
    shutil.rmtree(tmp, ignore_errors=True)
    # the step time is a difference of two process run times: samples where start-up jitter swallowed it (difference below 2 % of
    # the run) are dropped; if none is left, the whole run time stands in (an upper bound of the step time = a lower bound of the rate)
    good = [t for t, h in zip(times, his) if t > 0.02 * h]
    times = good if good else his
    tot = sum(times)
    rate = P * nelem * S * len(times) / tot
    cb = {"value": rate, "unit": "element-updates/s", "cores": P, "kind": "reference",
          "sample": f"{P} concurrent single-rank processes of the unmodified reference executable (one per host thread; "
                    f"no MPI/MUMPS in this image, so ranks do not exchange interface data: upper bound), each a {n}^3 "
                    f"lin3DHexa8 box (+DRM layer), {S} CentralDifference steps per bench step, timed as the difference "
                    f"between runs with nt={2 + S} and nt=2 (parse/Initialize cancel); Eigen replaced by oracle/shim"}
    line = {"metric": "element-updates/sec (FP64 explicit step)", "value": rate, "unit": "element-updates/s",
            "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": 1e3 * tot / len(times), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": f"3-D elastic half-space, lin3DHexa8 + Elastic3DLinear, lumped CentralDifference, DRM SV "
                                   f"plane-wave layer, 1 point load, 16 recorded nodes (BASELINE configs[3]-like) -- reference "
                                   f"CPU path on a bounded sample: {P} x {n}^3 elements, {S} steps per bench step"},
            "cpu_baseline": cb,
            "e2e": {"value": rate, "unit": "element-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def ncu_traffic(n, kern):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture of this very command (profiles/ncu_traffic.json); None when no capture matches."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            j = json.load(f)
        return j.get(f"{kern}@n{n}", {}).get("dram_bytes_per_launch")
    except OSError:
        return None


def parity_check(n, steps, device, emit_note=None):
    """bench.py --verify (on by default at N = 1): the SAME kernel variant and tiling rules as the timed run on an
    n^3 box with the DRM layer and the point load, every free dof started with a random velocity (so every
    node's stencil row matters from step 2 on), `steps` CentralDifference steps, FULL final state against the CPU oracle
    (oracle/svl_oracle.c, OpenMP over elements).  Tolerance 1e-10 (north_star)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Oracle
    from svl_b200 import model as M
    from svl_b200.capi import DeviceModel
    nt = steps + 1
    m = build_workload(n, nt)
    md = m
    if m.drm is not None:
        # the oracle reads a tabulated field, the device evaluates the plane wave itself (the kernels of the timed run);
        # the pulse is moved into the checked window (t0 = 8 dt, period 16 dt) so that the DRM forces are not negligible
        import copy
        G = m.global_elems_per_axis
        pw = dict(m.drm.planewave, t0=8.0 * m.dt, f0=1.0 / (16.0 * m.dt), c=m.drm.planewave["c"] * 50.0)   # and reaches all of it
        M.add_drm_box(m, x0=[G[0] / 2, G[1] / 2, G[2]], xl=[G[0] / 2 - 5.5, G[1] / 2 - 5.5, G[2] - 5.5], planewave=pw, tabulate_nt=nt)
        md = copy.copy(m)
        md.drm = copy.copy(m.drm)
        md.drm.field = None
    # random initial VELOCITY (U0 = 0): the reference's first step uses the stored (zero) stresses whatever U0 is
    V0 = np.random.default_rng(42).uniform(-1.0, 1.0, m.n_total)
    V0[np.asarray(m.totaldof)[np.asarray(m.freedof_flat) < 0]] = 0.0
    t0 = time.perf_counter()
    d = DeviceModel(md, device=device, max_rows=nt + 2, V0=V0)
    d.step(1, nt, True)
    U = d.get_state(0)
    c = d.counters()
    d.close()
    t_dev = time.perf_counter() - t0
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    _, Uref = Oracle().run(m, nt=nt, nthreads=cores, V0=V0)
    t_cpu = time.perf_counter() - t0
    err = float(np.abs(U - Uref).max() / np.abs(Uref).max())
    return {"mesh": f"{n}^3 lin3DHexa8 + DRM layer (plane wave evaluated on the device, tabulated for the oracle) + point load, "
                    f"random V0, {steps} steps", "dof": int(m.n_total),
            "block_nodes": int(c["n_block_nodes"]), "generic_elements": int(c["n_generic_elements"]),
            "max_rel_err_full_state_vs_oracle": err, "tol": 1e-10, "ok": bool(err < 1e-10),
            "oracle_s": t_cpu, "device_s_incl_plan": t_dev, "oracle_threads": cores}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--n", type=int, default=int(os.environ.get("SVL_BENCH_N", "320")), help="elements per side")
    ap.add_argument("--ref-n", type=int, default=16)
    ap.add_argument("--ref-steps", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-verify", action="store_true", help="skip the parity check of the timed kernel variant against the oracle")
    ap.add_argument("--verify-n", type=int, default=64, help="mesh of the parity check (the oracle's set-up costs ~80 us per "
                    "element: 64^3 ~ 20 s; profiles/ holds a 128^3 run)")
    ap.add_argument("--verify-steps", type=int, default=12)
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the strong-scaling sub-record")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = n^3 elements per GPU (default), strong = the n^3 mesh split over the GPUs")
    args = ap.parse_args()
    # the contract is ONE JSON line on stdout: libraries that print there (NCCL's version banner) go to stderr
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            run_reference_arm(args, emit)
        return

    from svl_b200 import capi
    from svl_b200.capi import DeviceModel
    K, W = args.steps, max(args.warmup, 3)
    nt = 4 * (W + K) + 40
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def allmax(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def barrier():
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    variant = "v4" if os.environ.get("SVLGPU_NO_SEP") else "sep"
    KV = KERNELS[variant]
    peak, peak_src = measured_peaks()
    fp64_peak, copy_peak = capi.measure_peaks(local_rank) if rank == 0 else (None, None)

    def replica_check(m, d):
        """every rank hashes the final displacements of the nodes it shares with each peer (the lists have the same
        order on both sides); rank 0 compares the pairs: replicas must agree bit for bit"""
        if dist is None:
            return None
        import hashlib
        mine = {}
        for peer, nodes in sorted(m.halos.items()):
            dofs = np.concatenate([np.asarray(m.totaldof)[m.node_ptr[q]:m.node_ptr[q + 1]] for q in nodes]).astype(np.int32)
            mine[int(peer)] = hashlib.sha256(d.get_state(0, dofs).tobytes()).hexdigest()
        allh = [None] * world
        dist.all_gather_object(allh, mine)
        pairs = bad = 0
        for r in range(world):
            for p_, hsh in allh[r].items():
                if p_ > r:
                    pairs += 1
                    bad += int(allh[p_].get(r) != hsh)
        return {"interface_pairs": pairs, "pairs_differing": bad, "ok": bad == 0,
                "what": "sha256 of U at the nodes shared by each pair of ranks after the timed steps"}

    def run_case(strong, K, W, with_e2e=True, with_kernel_timing=True, sample_clocks=False):
        t0 = time.perf_counter()
        m = build_workload(args.n, nt, rank, world, strong=strong)
        t_model = time.perf_counter() - t0
        t0 = time.perf_counter()
        d = DeviceModel(m, device=local_rank, max_rows=nt + 4, comm=comm_for())
        t_plan = time.perf_counter() - t0
        c = d.counters()
        n_elem_total = int(allsum(m.n_elem))
        n_dof_total = int(allsum(m.n_total))          # interface dofs counted once per replica
        amp = m.point_loads[0].series if m.point_loads else None
        # ---- device-resident throughput: K steps in one C-ABI call, CUDA events on the launching stream,
        #      barrier + synchronize on both sides, max over ranks
        k = 1
        d.step(k, k + W, True); k += W
        barrier()
        sampler.active = sample_clocks
        w0 = time.perf_counter()
        d.step(k, k + K, True); k += K
        barrier()
        wall = time.perf_counter() - w0

        ms = allmax(d.counters()["last_step_ms"])
        out = {"m": m, "value": n_elem_total * K / (ms * 1e-3), "ms_per_step": ms / K, "elements": n_elem_total,
               "dof": n_dof_total, "launches": allsum(d.counters()["launches_per_step"] * K), "c": c,
               "t_model": t_model, "t_plan": t_plan, "wall": wall}
        if with_kernel_timing:                        # per-launch CUDA-event durations of the kernels (separate pass)
            d.set_kernel_timing(True)
            d.step(k, k + min(K, 20), True); k += min(K, 20)
            out["kt"] = {w: d.kernel_time(w) for w in range(6)}
            d.set_kernel_timing(False)
        if with_e2e:                                  # end to end through the per-step C-ABI call with HOST buffers
            row = np.zeros(3 * len(m.rec_nodes))
            for _ in range(3):
                d.step_host(k, [amp[k]] if amp is not None else [], rec=0, row=row); k += 1
            barrier()
            e0 = time.perf_counter()
            for _ in range(K):
                d.step_host(k, [amp[k]] if amp is not None else [], rec=0, row=row); k += 1
            barrier()
            e2e_s = allmax(time.perf_counter() - e0)
            if sample_clocks and dist is None:
                # nvidia-smi answers in ~0.1 s and the timed regions are shorter than that: keep the SAME kernels running
                # (untimed) until the sampler has seen the GPU under this load a few times
                t_end = time.perf_counter() + 4.0
                while len(sampler.samples) < 6 and time.perf_counter() < t_end:
                    d.step(k, k + 50, True); k += 50
            elif sample_clocks:                       # several ranks: the same fixed number of extra steps on every rank
                d.step(k, k + 800, True); k += 800
            sampler.active = False
            if not np.all(np.isfinite(row)):
                raise SystemExit("non-finite response")
            out["e2e"] = {"value": n_elem_total * K / e2e_s, "unit": "element-updates/s",
                          "h2d_bytes_per_step": int(allsum(8 * len(m.point_loads))),
                          "d2h_bytes_per_step": int(allsum(row.nbytes)), "ms_per_step": 1e3 * e2e_s / K,
                          "note": "one svlgpu_step_host call per step and rank with HOST buffers: the step's load amplitudes "
                                  "are read from pinned host memory and the recorder row is written to pinned host memory by "
                                  "the step's own kernels (zero-copy over the bus, no separate copy-engine submissions), all "
                                  "kernels (+ NCCL interface exchange), stream sync; the state vectors stay resident in HBM "
                                  "exactly as the reference keeps U,V,A resident in host RAM between steps"}
        U = d.get_state(0, np.asarray(m.rec_dofs(), np.int32))
        if not np.all(np.isfinite(U)):
            raise SystemExit("non-finite state")
        out["replicas"] = replica_check(m, d)
        d.close()
        return out

    uid_count = [0]

    def comm_for():
        """a fresh NCCL communicator per model (weak case, strong case)"""
        if dist is None:
            return None
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        uid_count[0] += 1
        return (rank, world, bytes(uid.cpu().numpy()))

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    strong_main = args.scaling == "strong"
    R = run_case(strong_main, K, W, sample_clocks=True)
    sampler.active = False
    if rank == 0:
        sampler.stop_flag = True; sampler.join(timeout=2)
    m, c, kt = R["m"], R["c"], R["kt"]
    st_ms, st_n = kt[0]
    nodes = c["n_block_nodes"]
    # the dominant kernel advances the interior class; shell classes are a separate (timed) kernel
    real = KV["bytes"] * nodes / (st_ms * 1e-3) / 1e9 if st_n else None
    contract = HEX8_BYTES * nodes / (st_ms * 1e-3) / 1e9 if st_n else None
    fp64_exec = KV["flop"] * nodes / (st_ms * 1e-3) / 1e12 if st_n else None
    fp64_slots = 2.0 * KV["fp64_instr"] * nodes / (st_ms * 1e-3) / 1e12 if st_n else None
    roof = {"bound": "hbm", "kernel": f"{KV['name']} (block-stencil force + CentralDifference update, rank 0)",
            "achieved": real, "peak": peak, "unit": "GB/s", "frac": (real / peak) if real else None,
            "traffic": ncu_traffic(args.n if world == 1 else None, KV["name"]),
            "peak_source": peak_src, "avg_launch_ms": st_ms, "launches_timed": st_n,
            "bytes_per_node": KV["bytes"],
            "what": "achieved = compulsory bytes of the kernel (U_n and U_{n-1} read, U_{n+1} written: 72 B per node; the class "
                    "table replaces connectivity, coordinates and masses) x lattice nodes / CUDA-event launch time; ncu's "
                    "dram__bytes per launch is `traffic`",
            "contract_equivalent": {"bytes_per_element_update": HEX8_BYTES, "GBs": contract,
                                    "frac": (contract / peak) if contract else None,
                                    "note": "SURVEY 8(d) 140 B per element-update figure; exceeds 1 because the kernel does "
                                            "not move connectivity / coordinates / mass at all"},
            "copy_GBs_measured_here": copy_peak, "frac_of_copy_measured_here": (real / copy_peak) if real and copy_peak else None,
            "frac_of_8TBs_nominal": (real / 8000.0) if real else None,
            "fp64_peak_measured_TFLOPs": fp64_peak,
            "fp64_executed_TFLOPs": fp64_exec,
            "fp64_fraction": (fp64_slots / fp64_peak) if fp64_slots and fp64_peak else None,
            "fp64_what": "fp64_fraction = FP64 instructions executed per node (SASS count: %.1f, every one a pipe slot of an "
                         "FMA) x 2 flop / time / measured DFMA peak (svlgpu_measure_peaks); fp64_executed counts DADD / DMUL "
                         "as 1 flop" % KV["fp64_instr"],
            "step_floor": {"ms": KV["bytes"] * nodes / (peak * 1e9) * 1e3, "frac": KV["bytes"] * nodes / (peak * 1e9) * 1e3 / R["ms_per_step"],
                           "note": "whole step (DRM layer, shell classes, loads, recorder included) against the time to move "
                                   "the dominant kernel's compulsory bytes at peak"}}

    cb = None
    pc = None
    if rank == 0 and world == 1:
        if not args.no_cpu_baseline:
            cb = cpu_baseline_port()
        if not args.no_verify:
            pc = parity_check(args.verify_n, args.verify_steps, local_rank)
    strong = None
    if world > 1 and not strong_main and not args.no_strong:
        # strong scaling of the SAME n^3 mesh (north_star: the 10^8-DOF mesh on 8 GPUs) next to the weak-scaling value
        S = run_case(True, K, W, with_e2e=False, with_kernel_timing=False)
        strong = {"value": S["value"], "unit": "element-updates/s", "ms_per_step": S["ms_per_step"], "elements": S["elements"],
                  "dof": S["dof"], "mesh": "x".join(str(g) for g in S["m"].global_elems_per_axis), "replicas": S["replicas"],
                  "note": "the N = 1 mesh split over the GPUs; efficiency = value / (N x the N = 1 line's value)"}

    G = m.global_elems_per_axis
    line = {"metric": "element-updates/sec (FP64 explicit step)", "value": R["value"], "unit": "element-updates/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": R["ms_per_step"], "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"3-D elastic half-space, {G[0]}x{G[1]}x{G[2]} lin3DHexa8 + Elastic3DLinear "
                                   f"({R['dof']} DOF), lumped CentralDifference, DRM SV plane-wave layer, "
                                   f"1 host-fed point load, 16 recorded nodes per partition (BASELINE configs[3]-like; "
                                   f"{world} block partition(s), one per GPU, NCCL interface-force exchange)",
                       "elements": R["elements"], "dof": R["dof"], "dt": m.dt,
                       "partition_grid": list(__import__("svl_b200.partition", fromlist=["x"]).proc_grid(world)),
                       "interface_nodes_rank0": int(sum(len(v) for v in m.halos.values())),
                       "l2": "state vectors (3 x %.0f MB per GPU) exceed the 126 MB L2" % (m.n_total * 8 / 1e6),
                       "block_nodes_rank0": c["n_block_nodes"], "generic_elements_rank0": c["n_generic_elements"],
                       "node_classes_rank0": c["n_node_classes"], "model_build_s": R["t_model"], "plan_upload_s": R["t_plan"],
                       "wall_s_timed_region": R["wall"], "dominant_kernel": KV["name"]},
            "clocks": sampler.summary() if rank == 0 else None, "e2e": R["e2e"], "gpu_launches": int(R["launches"]),
            "roofline": roof,
            "kernel_ms": {"stencil_dom": kt[0][0], "stencil_shell_gather": kt[4][0], "gauss_elements": kt[1][0],
                          "gather_nodes": kt[2][0], "point_loads": kt[3][0], "drm": kt[5][0]},
            "cpu_baseline": cb, "parity_check": pc, "replicas": R["replicas"], "strong": strong}
    if rank == 0:
        emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
