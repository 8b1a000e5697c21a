"""Element-based domain decomposition, one partition per GPU.

Mirrors what the reference pre-processor does when it writes one JSON file per rank
(01-Pre_Process/Core/SeismoVLAB.py:300-420, Partition.py:87-201): elements are assigned to
partitions (METIS `mpmetis` there; a geometric block split for the synthetic structured meshes here, or
any `epart` array read from a METIS `*.epart` file), interface nodes are duplicated in every partition
that touches them (:354-358), point loads / nodal masses / recorders are handed to exactly one partition
(:163-180, :118-122).  The only extra product of this module is the per-neighbour list of shared
interface nodes that `svlgpu_add_halo` needs (both sides in ascending global node order).
"""
from __future__ import annotations

import copy
from typing import Dict, List

import numpy as np

from .model import ELEM_NODES, DRMLoad, Model, PointLoad, make_box_model, add_drm_box


def proc_grid(nparts: int):
    """(px, py, pz) with px*py*pz = nparts, as cubic as possible, z split first (1,1,2), (1,2,2), (2,2,2)."""
    g = [1, 1, 1]
    ax = 2
    n = nparts
    f = 2
    while n > 1:
        while n % f:
            f += 1
        g[ax] *= f
        n //= f
        ax = (ax - 1) % 3
    return tuple(g)


def block_epart(ne, pgrid) -> np.ndarray:
    """Element -> partition for an nx x ny (x nz) box split into pgrid blocks (element order of
    Builder.py:167-183: x fastest)."""
    nd = len(ne)
    idx = np.meshgrid(*[np.arange(n) for n in reversed(ne)], indexing="ij")
    ijk = [g.ravel() for g in reversed(idx)]
    part = np.zeros(len(ijk[0]), dtype=np.int32)
    mult = 1
    for a in range(nd):
        pa = np.minimum(ijk[a] * pgrid[a] // ne[a], pgrid[a] - 1)
        part += (pa * mult).astype(np.int32)
        mult *= pgrid[a]
    return part


def centroid_epart(m: Model, pgrid) -> np.ndarray:
    """Element -> partition by the position of the element centroid in the model's bounding box (a geometric block split
    for models that are not one lattice, e.g. a soil box wrapped in its PML layer); pgrid has one entry per axis."""
    npe_e = np.array([ELEM_NODES[int(k)] for k in np.unique(m.elem_kind)])[np.searchsorted(np.unique(m.elem_kind), m.elem_kind)]
    cen = np.zeros((m.n_elem, m.ndim))
    for npe in np.unique(npe_e):                           # vectorised per node count (millions of elements at bench sizes)
        sel = np.nonzero(npe_e == npe)[0]
        cen[sel] = m.coords[m.elem_conn[sel, :npe]].mean(axis=1)
    lo, hi = m.coords.min(axis=0), m.coords.max(axis=0)
    part = np.zeros(m.n_elem, dtype=np.int32)
    mult = 1
    for a in range(m.ndim):
        pa = np.minimum(((cen[:, a] - lo[a]) / (hi[a] - lo[a]) * pgrid[a]).astype(np.int64), pgrid[a] - 1)
        part += (pa * mult).astype(np.int32)
        mult *= pgrid[a]
    return part


def weighted_epart(m: Model, pgrid, pml_weight: float = 250.0) -> np.ndarray:
    """Element -> partition by recursive coordinate bisection with element WEIGHTS: per axis (z first, like proc_grid) every
    group of the previous level is cut into pgrid[axis] slabs of equal total weight, cuts only BETWEEN layers of equal centroid
    coordinate (so the interfaces stay planar).  A PML element costs ~250 soil elements per step on the device (9.0 ms for
    2.16 M PML elements vs 0.51 ms for 32.8 M soil elements), so the cuts balance the PML shell and the soil follows: what
    METIS does with vertex weights (`mpmetis` takes them the same way) instead of the plain geometric split of centroid_epart."""
    npe_e = np.array([ELEM_NODES[int(k)] for k in np.unique(m.elem_kind)])[np.searchsorted(np.unique(m.elem_kind), m.elem_kind)]
    cen = np.zeros((m.n_elem, m.ndim))
    for npe in np.unique(npe_e):
        sel = np.nonzero(npe_e == npe)[0]
        cen[sel] = m.coords[m.elem_conn[sel, :npe]].mean(axis=1)
    w = np.where(np.isin(m.elem_kind, (3, 4)), float(pml_weight), 1.0)
    part = np.zeros(m.n_elem, dtype=np.int32)
    mult = [1] * m.ndim
    for a in range(1, m.ndim):
        mult[a] = mult[a - 1] * pgrid[a - 1]
    groups = [np.arange(m.n_elem)]
    for a in reversed(range(m.ndim)):
        nxt = []
        for g in groups:
            if pgrid[a] == 1 or len(g) == 0:
                nxt.append(g)
                continue
            key = np.round(cen[g, a], 9)
            layers, inv = np.unique(key, return_inverse=True)
            lw = np.bincount(inv, weights=w[g], minlength=len(layers))
            cum = np.cumsum(lw)
            cuts = [0]
            for q in range(1, pgrid[a]):
                target = cum[-1] * q / pgrid[a]
                j = int(np.argmin(np.abs(cum - target))) + 1                 # cut after layer j-1
                cuts.append(min(max(j, cuts[-1] + 1), len(layers) - (pgrid[a] - q)))
            cuts.append(len(layers))
            slab = np.searchsorted(np.array(cuts[1:-1]), inv, side="right")
            for q in range(pgrid[a]):
                sel = g[slab == q]
                part[sel] += q * mult[a]
                nxt.append(sel)
        groups = nxt
    return part


def write_metis_graph(m: Model, path: str) -> str:
    """The METIS mesh file the reference's pre-processor hands to `mpmetis` (Core/Partition.py:87-144 SetMetisInputFile):
    first line = number of elements, then one line of 1-based node ids per element in ascending element order, every id
    followed by a blank.  The slave node of an EQUAL constraint carries the id of its master (:121-126), so that METIS sees
    the soil-PML interface as connected and keeps tied nodes together.  `mpmetis <path> <nparts>` then writes
    `<path>.epart.<nparts>`, which read_epart turns into the `epart` of split_model / write_reference_partitions."""
    ids = np.arange(1, m.n_nodes + 1, dtype=np.int64)
    if m.constraints:
        fd = np.asarray(m.freedof_flat)
        node_of_total = np.repeat(np.arange(m.n_nodes), np.diff(m.node_ptr))
        total_of_free = -np.ones(max(1, m.n_free), dtype=np.int64)
        total_of_free[fd[fd > -1]] = np.nonzero(fd > -1)[0]
        aux = ids.copy()
        for _, slave, masters, _ in m.constraints:
            if len(masters) == 1:                                   # EQUAL
                ids[node_of_total[slave]] = aux[node_of_total[total_of_free[masters[0]]]]
    with open(path, "w") as f:
        f.write(f"{m.n_elem}\n")
        for e in range(m.n_elem):
            f.write("".join(f"{ids[n]} " for n in m.elem_conn[e, :ELEM_NODES[int(m.elem_kind[e])]]) + "\n")
    return path


def read_epart(path: str) -> np.ndarray:
    """METIS `<graph>.epart.<nparts>` file: one partition id per element line (Partition.py:150-201)."""
    return np.loadtxt(path, dtype=np.int32)


def split_model(m: Model, epart: np.ndarray, nparts: int, tie_closure: str = "both", ranks=None) -> List[Model]:
    """Global model + element partition -> one sub-model per rank, each carrying
    `.halos = {peer: local node indices}`, `.global_nodes`, `.global_elems`.
    tie_closure: "both" (what the device needs: a partition that holds the slave OR a master of an EQUAL constraint gets all
    of its nodes) or "slave" (what the reference pre-processor writes: only the partition of the slave gets the masters,
    SeismoVLAB.py:381-392; the host driver completes the closure when it reads the files).
    ranks: build only these partitions' sub-models (the others are None in the returned list) -- what one rank of a torchrun job
    needs; the node sets of all partitions are still derived, so the halo lists are complete."""
    kinds = np.unique(m.elem_kind)
    npe_e = np.array([ELEM_NODES[int(k)] for k in kinds], dtype=np.int32)[np.searchsorted(kinds, m.elem_kind)]

    def nodes_of(el):
        """unique nodes of the listed elements (each element uses its own node count)"""
        el = np.asarray(el)
        if len(el) == 0:
            return np.zeros(0, dtype=np.int64)
        return np.unique(np.concatenate([m.elem_conn[el[npe_e[el] == npe], :npe].ravel() for npe in np.unique(npe_e[el])]))

    epart = np.asarray(epart, dtype=np.int32)
    if len(epart) < m.n_elem:
        # elements beyond the partitioned solids (ZeroLength1D dashpots appended after the lattice): each goes to the
        # lowest rank that holds one of its nodes, so its damping is counted exactly once (svlgpu_comm_init sums the
        # lumped mass / damping diagonals of the interface dofs over the ranks)
        first_rank = np.full(m.n_nodes, nparts, dtype=np.int32)
        for e in range(len(epart)):
            nn = m.elem_conn[e, :npe_e[e]]
            first_rank[nn] = np.minimum(first_rank[nn], epart[e])
        extra = []
        for e in range(len(epart), m.n_elem):
            r = int(first_rank[m.elem_conn[e, :npe_e[e]]].min())
            if r >= nparts:
                raise ValueError("element outside the partitioned mesh touches no partitioned node")
            extra.append(r)
        epart = np.concatenate([epart, np.array(extra, dtype=np.int32)])
    node_sets = []
    for r in range(nparts):
        el = np.nonzero(epart == r)[0]
        node_sets.append(nodes_of(el))
    # EQUAL constraints (the soil-PML ties of Builder.py:653-666): a partition that holds the slave or a master of a tie
    # gets all of its nodes, so every replica of a tied dof resolves to the same unknown and the interface lists stay
    # mirror images (the reference writes the constraint into every partition that holds its slave: SeismoVLAB.py:360-372)
    fd_all = np.asarray(m.freedof_flat)
    node_of_total = np.repeat(np.arange(m.n_nodes), np.diff(m.node_ptr))
    total_of_free = -np.ones(max(1, m.n_free), dtype=np.int64)
    total_of_free[fd_all[fd_all > -1]] = np.nonzero(fd_all > -1)[0]
    ties = []                                              # (tag, slave total dof, [master total dofs], factors)
    for tag, slave, masters, factors in m.constraints:
        ties.append((int(tag), int(slave), [int(total_of_free[f]) for f in masters], list(factors)))
    if ties:
        # (slave node, master node) pairs, one per master of every tie
        ta = np.array([node_of_total[sl] for _, sl, mt, _ in ties for _t in mt], dtype=np.int64)
        tb = np.array([node_of_total[t] for _, sl, mt, _ in ties for t in mt], dtype=np.int64)
        for r in range(nparts):
            have = np.zeros(m.n_nodes, dtype=bool)
            have[node_sets[r]] = True
            while True:
                ha, hb = have[ta], have[tb]
                if (ha == hb).all():
                    break
                if tie_closure == "both":
                    have[ta[hb]] = True
                have[tb[ha]] = True
                if tie_closure != "both":
                    break
            node_sets[r] = np.nonzero(have)[0]
    pml_anywhere = bool(np.isin(m.elem_kind, (3, 4)).any())
    # owner of a node = lowest rank that holds it
    owner = np.full(m.n_nodes, nparts, dtype=np.int32)
    for r in reversed(range(nparts)):
        owner[node_sets[r]] = r
    fd = np.asarray(m.freedof_flat)
    subs = []
    for r in range(nparts):
        if ranks is not None and r not in ranks:
            subs.append(None)
            continue
        el = np.nonzero(epart == r)[0]
        gn = node_sets[r]
        loc = -np.ones(m.n_nodes, dtype=np.int64)
        loc[gn] = np.arange(len(gn))
        s = Model(ndim=m.ndim, lumped=m.lumped)
        s.coords = m.coords[gn]
        s.node_ndof = m.node_ndof[gn]
        # free dofs get a placeholder (renumbered below), restrained (-1) and constrained (tag < -1) codes are kept
        s.freedof = [np.where(fd[m.node_ptr[n]:m.node_ptr[n + 1]] > -1, 0, fd[m.node_ptr[n]:m.node_ptr[n + 1]]).astype(np.int32)
                     for n in gn]
        s.materials = list(m.materials)
        conn = np.zeros((len(el), 8), dtype=np.int32)
        for npe in np.unique(npe_e[el]):
            sel = np.nonzero(npe_e[el] == npe)[0]
            conn[sel, :npe] = loc[m.elem_conn[el[sel], :npe]]
        s.elem_conn = conn
        s.elem_kind = m.elem_kind[el]
        s.elem_mat = m.elem_mat[el]
        s.elem_attr = m.elem_attr[el] if m.elem_attr is not None else None
        s.elem_am = m.elem_am[el] if m.elem_am is not None else None
        s.elem_ak = m.elem_ak[el] if m.elem_ak is not None else None
        s.masses = [(int(loc[n]), v) for n, v in m.masses if owner[n] == r]
        # a moving support moves on every replica of its node (it is a prescribed displacement, not a force to be summed)
        s.supports = [(int(loc[n]), d, sr, fc) for n, d, sr, fc in (getattr(m, "supports", None) or []) if loc[n] >= 0]
        s.point_loads = []
        for pl in m.point_loads:
            keep = [int(loc[n]) for n in pl.nodes if owner[n] == r]
            if keep:
                s.point_loads.append(PointLoad(np.array(keep, dtype=np.int32), pl.dir.copy(), pl.series.copy(), pl.factor))
        if m.drm is not None:
            d = m.drm
            eloc = -np.ones(m.n_elem, dtype=np.int64)
            eloc[el] = np.arange(len(el))
            mine = eloc[d.elems] >= 0
            if mine.any():
                de = eloc[d.elems[mine]].astype(np.int32)
                dn_glob = nodes_of(d.elems[mine])
                pos = np.searchsorted(d.nodes, dn_glob)
                assert (d.nodes[pos] == dn_glob).all()
                s.drm = DRMLoad(elems=de, nodes=loc[dn_glob].astype(np.int32), exterior=d.exterior[pos].copy(),
                                field=None if d.field is None else d.field[pos].copy(),
                                planewave=copy.deepcopy(d.planewave), factor=d.factor)
        s.dt, s.nt = m.dt, m.nt
        rn = [int(loc[n]) for n in (m.rec_nodes if m.rec_nodes is not None else []) if owner[n] == r]
        s.rec_nodes = np.array(rn, dtype=np.int32)
        s.rec_global = np.array([int(n) for n in (m.rec_nodes if m.rec_nodes is not None else []) if owner[n] == r],
                                dtype=np.int32)
        s.blocks = _sub_lattice_hint(m, gn)
        s.number_dofs()
        s.global_nodes, s.global_elems = gn, el

        def local_total(t):
            n = node_of_total[t]
            return int(s.node_ptr[loc[n]] + (t - m.node_ptr[n]))

        s.constraints = []
        for tag, sl, mt, fac in ties:
            if loc[node_of_total[sl]] < 0:
                continue
            s.constraints.append((tag, local_total(sl), [int(s.freedof_flat[local_total(t)]) for t in mt], list(fac)))
        # every rank of a model with PML elements takes part in the block solve's reductions, also a rank without PML
        # unknowns (svlgpu option "pml_collective")
        s.pml_collective = pml_anywhere
        s.halos = {}
        subs.append(s)
    for r in range(nparts):
        if subs[r] is None:
            continue
        for q in range(nparts):
            if q == r:
                continue
            shared = np.intersect1d(node_sets[r], node_sets[q], assume_unique=True)
            if len(shared):
                subs[r].halos[q] = np.searchsorted(node_sets[r], shared).astype(np.int32)
    return subs


def _sub_lattice_hint(m: Model, gn: np.ndarray):
    """If the global model is one lattice block and this partition's nodes form a full sub-box of it, the
    local numbering (ascending global id) is again x-fastest: hint the planner."""
    if len(m.blocks) != 1:
        return []
    n0, NX, NY, NZ = m.blocks[0]
    q = gn - n0
    inl = (q >= 0) & (q < NX * NY * NZ)
    nl = int(inl.sum())
    # nodes outside the lattice (e.g. the fixed twins of Lysmer dashpots) must follow the lattice nodes in the local numbering
    if nl == 0 or n0 != 0 or not inl[:nl].all():
        return []
    q = q[:nl]
    gn = gn[:nl]
    i, j, k = q % NX, (q // NX) % NY, q // (NX * NY)
    ni, nj, nk = i.max() - i.min() + 1, j.max() - j.min() + 1, k.max() - k.min() + 1
    if ni * nj * nk != len(gn):
        return []
    return [(0, int(ni), int(nj), int(nk))]


# -------------------------------------------------------------------------------
# direct per-rank construction of a block of a large synthetic box (bench.py): the global model is never
# materialised; the halo lists follow from the lattice arithmetic
# -------------------------------------------------------------------------------
def local_box(n_local, pgrid, rank, h=1.0, **kw) -> Model:
    """Rank `rank` of a (px*nx) x (py*ny) x (pz*nz) element box split into px x py x pz blocks of
    n_local = (nx, ny, nz) elements.  Bottom fixed on the lowest layer of ranks only."""
    px, py, pz = pgrid
    rx, ry, rz = rank % px, (rank // px) % py, rank // (px * py)
    nx, ny, nz = n_local
    origin = (rx * nx * h, ry * ny * h, rz * nz * h)
    m = make_box_model(n_local, h, fix="bottom" if rz == 0 else None, origin=origin, **kw)
    NX, NY, NZ = nx + 1, ny + 1, nz + 1
    i, j, k = np.meshgrid(np.arange(NX), np.arange(NY), np.arange(NZ), indexing="ij")
    lid = (i + NX * j + NX * NY * k)
    m.halos = {}
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                if dx == dy == dz == 0:
                    continue
                qx, qy, qz = rx + dx, ry + dy, rz + dz
                if not (0 <= qx < px and 0 <= qy < py and 0 <= qz < pz):
                    continue
                sel = [slice(None)] * 3
                for a, (d, N) in enumerate(((dx, NX), (dy, NY), (dz, NZ))):
                    if d == -1:
                        sel[a] = slice(0, 1)
                    elif d == 1:
                        sel[a] = slice(N - 1, N)
                nodes = np.sort(lid[tuple(sel)].ravel()).astype(np.int32)
                m.halos[qx + px * qy + px * py * qz] = nodes
    m.grid_pos = (rx, ry, rz)
    return m
