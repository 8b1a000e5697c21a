"""Host-side model container, synthetic structured generators and the writer of the
reference's per-rank JSON partition files.

This mirrors what the reference's pre-processor produces (01-Pre_Process):
  * node / element numbering of ``makeDomainVolume`` / ``makeDomainArea``
    (Method/Builder.py:134-141, 167-183, 244-417),
  * the ``PlainScheme`` total/free DOF numbering (Core/Numberer.py:185-233),
  * the per-rank JSON schema consumed by Driver.hpp:1981-2046.

A ``Model`` is plain numpy data; ``svl_b200.capi.upload`` turns it into a device
model through the C ABI (include/svlgpu.h), ``tests/oracle_lib.py`` turns the same
object into the oracle's input, and ``write_reference_json`` into the files the
reference executable reads.  Nothing here computes physics.
"""
from __future__ import annotations

import copy
import json
import math
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

LIN3DHEXA8, LIN2DQUAD4, PML3DHEXA8, PML2DQUAD4, ZEROLENGTH1D = 1, 2, 3, 4, 5
ELASTIC3DLINEAR, ELASTIC2DPLANESTRAIN, PLASTIC3DJ2, PLASTICPLANESTRAINJ2, VISCOUS1DLINEAR = 1, 2, 3, 4, 5

ELEM_NAME = {LIN3DHEXA8: "LIN3DHEXA8", LIN2DQUAD4: "LIN2DQUAD4", PML3DHEXA8: "PML3DHEXA8",
             PML2DQUAD4: "PML2DQUAD4", ZEROLENGTH1D: "ZEROLENGTH1D"}
ELEM_NODES = {LIN3DHEXA8: 8, LIN2DQUAD4: 4, PML3DHEXA8: 8, PML2DQUAD4: 4, ZEROLENGTH1D: 2}
ELEM_NATTR = {LIN3DHEXA8: 0, LIN2DQUAD4: 1, PML3DHEXA8: 9, PML2DQUAD4: 8, ZEROLENGTH1D: 1}
MAT_NAME = {ELASTIC3DLINEAR: "ELASTIC3DLINEAR", ELASTIC2DPLANESTRAIN: "ELASTIC2DPLANESTRAIN",
            PLASTIC3DJ2: "PLASTIC3DJ2", PLASTICPLANESTRAINJ2: "PLASTICPLANESTRAINJ2",
            VISCOUS1DLINEAR: "VISCOUS1DLINEAR"}
MAT_KEYS = {ELASTIC3DLINEAR: ["E", "nu", "rho"], ELASTIC2DPLANESTRAIN: ["E", "nu", "rho"],
            PLASTIC3DJ2: ["K", "G", "rho", "h", "beta", "Sy"],
            PLASTICPLANESTRAINJ2: ["K", "G", "rho", "h", "beta", "Sy"],
            VISCOUS1DLINEAR: ["eta"]}


@dataclass
class PointLoad:
    nodes: np.ndarray          # node indices (0-based)
    dir: np.ndarray            # ndim direction
    series: np.ndarray         # nt amplitudes (len 1 = constant)
    factor: float = 1.0


@dataclass
class DRMLoad:
    elems: np.ndarray          # element indices
    nodes: np.ndarray          # node indices
    exterior: np.ndarray       # uint8 per node
    field: Optional[np.ndarray] = None   # [nnodes, nt, 3*ndim]  (u, v, a) as in .drm files
    planewave: Optional[dict] = None     # analytic alternative, see capi
    factor: float = 1.0


@dataclass
class Model:
    ndim: int
    lumped: bool = True
    coords: np.ndarray = None            # [n, ndim]
    node_ndof: np.ndarray = None         # [n]
    freedof: List[np.ndarray] = None     # per node list, -1 = restrained, >=0 placeholder
    materials: List[tuple] = field(default_factory=list)        # (kind, [params])
    elem_kind: np.ndarray = None         # [ne]
    elem_conn: np.ndarray = None         # [ne, 8] (quads: first 4)
    elem_mat: np.ndarray = None          # [ne]
    elem_attr: np.ndarray = None         # [ne, 10]
    elem_am: Optional[np.ndarray] = None
    elem_ak: Optional[np.ndarray] = None
    constraints: List[tuple] = field(default_factory=list)      # (tag, slave_total, [master_free], [factor])
    masses: List[tuple] = field(default_factory=list)           # (node, [mass per dof])
    point_loads: List[PointLoad] = field(default_factory=list)
    drm: Optional[DRMLoad] = None
    blocks: List[tuple] = field(default_factory=list)           # (node0, nx, ny, nz) lattice hints
    dt: float = 0.0
    nt: int = 0
    rec_nodes: np.ndarray = None
    # support motions: (node index, local dof, series Xo, combination factor of the SUPPORTMOTION load that lists the node)
    # -- Driver.hpp:509-563 + :1725-1735, Assembler.cpp:493-533
    supports: List[tuple] = field(default_factory=list)

    # ---- derived numbering (PlainScheme) -------------------------------------
    def number_dofs(self):
        n = len(self.node_ndof)
        self.node_ptr = np.zeros(n + 1, dtype=np.int32)
        np.cumsum(self.node_ndof, out=self.node_ptr[1:])
        self.n_total = int(self.node_ptr[-1])
        self.totaldof = np.arange(self.n_total, dtype=np.int32)
        fd = np.concatenate(self.freedof).astype(np.int32) if isinstance(self.freedof, list) \
            else np.asarray(self.freedof, dtype=np.int32)
        free = fd.copy()
        is_free = fd > -1
        free[is_free] = np.arange(int(is_free.sum()), dtype=np.int32)
        self.freedof_flat = free
        self.n_free = int(is_free.sum())
        return self

    @property
    def n_nodes(self):
        return len(self.node_ndof)

    @property
    def n_elem(self):
        return len(self.elem_kind)

    def node_dofs(self, node):
        return self.totaldof[self.node_ptr[node]:self.node_ptr[node + 1]]

    def rec_dofs(self):
        return np.concatenate([self.node_dofs(n) for n in self.rec_nodes]).astype(np.int32)


# -------------------------------------------------------------------------------
# generators
# -------------------------------------------------------------------------------
def box_nodes(ne, P0, P1, P2, P3):
    """Builder.py:134-141: tag = 1 + i + (nx+1) j + (nx+1)(ny+1) k, coords P0 + i DX + j DY + k DZ."""
    nx, ny, nz = ne
    P0, P1, P2, P3 = (np.asarray(p, dtype=np.float64) for p in (P0, P1, P2, P3))
    DX, DY, DZ = (P1 - P0) / nx, (P2 - P0) / ny, (P3 - P0) / nz
    k, j, i = np.meshgrid(np.arange(nz + 1), np.arange(ny + 1), np.arange(nx + 1), indexing="ij")
    i, j, k = i.ravel(), j.ravel(), k.ravel()
    # same operation order as the reference: ((P0 + i*DX) + j*DY) + k*DZ
    return ((P0[None, :] + i[:, None] * DX[None, :]) + j[:, None] * DY[None, :]) + k[:, None] * DZ[None, :]


def box_hex8_conn(ne, node0=0):
    """Builder.py:167-183 (0-based): [n1..n8] in VTK hexahedron order."""
    nx, ny, nz = ne
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    i, j, k = i.ravel(), j.ravel(), k.ravel()
    sx, sy = nx + 1, (nx + 1) * (ny + 1)
    n1 = sy * k + j * sx + i
    conn = np.stack([n1, n1 + 1, n1 + sx + 1, n1 + sx, n1 + sy, n1 + sy + 1, n1 + sy + sx + 1, n1 + sy + sx],
                    axis=1)
    return (conn + node0).astype(np.int32)


def area_nodes(ne, P0, P1, P2):
    nx, ny = ne
    P0, P1, P2 = (np.asarray(p, dtype=np.float64) for p in (P0, P1, P2))
    DX, DY = (P1 - P0) / nx, (P2 - P0) / ny
    j, i = np.meshgrid(np.arange(ny + 1), np.arange(nx + 1), indexing="ij")
    i, j = i.ravel(), j.ravel()
    return (P0[None, :] + i[:, None] * DX[None, :]) + j[:, None] * DY[None, :]


def area_quad4_conn(ne, node0=0):
    """Builder.py:244-417 makeDomainArea QUAD4: [n1, n2, n3, n4] counter-clockwise."""
    nx, ny = ne
    j, i = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    i, j = i.ravel(), j.ravel()
    sx = nx + 1
    n1 = j * sx + i
    conn = np.stack([n1, n1 + 1, n1 + sx + 1, n1 + sx], axis=1)
    return (conn + node0).astype(np.int32)


def ricker(nt, dt, f0, t0):
    """Ricker pulse (1-2b) exp(-b), b = (pi f0 (t-t0))^2  (PlaneWave.py:222-223)."""
    t = np.arange(nt) * dt
    b = (math.pi * f0 * (t - t0)) ** 2
    return (1.0 - 2.0 * b) * np.exp(-b)


def make_box_model(ne, h=1.0, mat=(ELASTIC3DLINEAR, [1.3e7, 0.3, 2000.0]), fix="bottom",
                   dt=None, nt=0, load_node=None, load_dir=(0.0, 0.0, 1.0e4), series=None,
                   rec_nodes=None, jitter=0.0, seed=20260117, layers=None, origin=(0.0, 0.0, 0.0)) -> Model:
    """Config-1-like soil box: nx*ny*nz lin3DHexa8 on [0,nx h]x[0,ny h]x[0,nz h], bottom fixed,
    lumped mass, vertical point load on the top-centre node.  `layers`: optional list of
    (material tuple) assigned per element layer in z (cyclic)."""
    nx, ny, nz = ne
    ox, oy, oz = origin
    X = box_nodes(ne, [ox, oy, oz], [ox + nx * h, oy, oz], [ox, oy + ny * h, oz], [ox, oy, oz + nz * h])
    if jitter:
        rng = np.random.default_rng(seed)
        X = X + jitter * h * rng.uniform(-1.0, 1.0, X.shape)
    n = X.shape[0]
    m = Model(ndim=3)
    m.coords = X
    m.node_ndof = np.full(n, 3, dtype=np.int32)
    fd = np.zeros((n, 3), dtype=np.int32)
    NX, NY = nx + 1, ny + 1
    if fix == "bottom":
        fd[: NX * NY, :] = -1
    m.freedof = fd.reshape(-1)
    mats = [mat] if layers is None else list(layers)
    m.materials = mats
    conn = box_hex8_conn(ne)
    m.elem_conn = conn
    m.elem_kind = np.full(len(conn), LIN3DHEXA8, dtype=np.int32)
    if layers is None:
        m.elem_mat = np.zeros(len(conn), dtype=np.int32)
    else:
        kz = np.arange(len(conn)) // (nx * ny)
        m.elem_mat = (kz % len(mats)).astype(np.int32)
    m.elem_attr = None                      # lin3DHexa8 carries no attributes
    m.blocks = [(0, NX, NY, nz + 1)] if not jitter else []
    if dt is None:
        E, nu, rho = mats[0][1][:3] if mats[0][0] == ELASTIC3DLINEAR else (None, None, None)
        if E is None:
            K, G, rho = mats[0][1][0], mats[0][1][1], mats[0][1][2]
            vp = math.sqrt((K + 4.0 * G / 3.0) / rho)
        else:
            lam = E * nu / ((1 + nu) * (1 - 2 * nu)); mu = E / (2 * (1 + nu))
            vp = math.sqrt((lam + 2 * mu) / rho)
        dt = 0.5 * h / vp
    m.dt, m.nt = dt, nt
    if load_node is None:
        load_node = (nx // 2) + NX * (ny // 2) + NX * NY * nz
    if series is None and nt > 0:
        f0 = 1.0 / (20.0 * dt)
        series = ricker(nt, dt, f0, 1.2 / f0)
    if series is not None:
        m.point_loads = [PointLoad(np.array([load_node], dtype=np.int32), np.asarray(load_dir, float),
                                   np.asarray(series, float))]
    m.rec_nodes = np.asarray(rec_nodes if rec_nodes is not None else [load_node], dtype=np.int32)
    return m.number_dofs()


def make_area_model(ne, h=1.0, th=1.0, mat=(ELASTIC2DPLANESTRAIN, [1.3e7, 0.3, 2000.0]), fix="bottom",
                    dt=None, nt=0, load_node=None, load_dir=(0.0, 1.0e4), series=None,
                    rec_nodes=None, jitter=0.0, seed=20260117) -> Model:
    nx, ny = ne
    X = area_nodes(ne, [0, 0], [nx * h, 0], [0, ny * h])
    if jitter:
        rng = np.random.default_rng(seed)
        X = X + jitter * h * rng.uniform(-1.0, 1.0, X.shape)
    n = X.shape[0]
    m = Model(ndim=2)
    m.coords = X
    m.node_ndof = np.full(n, 2, dtype=np.int32)
    fd = np.zeros((n, 2), dtype=np.int32)
    if fix == "bottom":
        fd[: nx + 1, :] = -1
    m.freedof = fd.reshape(-1)
    m.materials = [mat]
    conn4 = area_quad4_conn(ne)
    conn = np.zeros((len(conn4), 8), dtype=np.int32)
    conn[:, :4] = conn4
    m.elem_conn = conn
    m.elem_kind = np.full(len(conn), LIN2DQUAD4, dtype=np.int32)
    m.elem_mat = np.zeros(len(conn), dtype=np.int32)
    m.elem_attr = np.zeros((len(conn), 10))
    m.elem_attr[:, 0] = th
    m.blocks = [(0, nx + 1, ny + 1, 1)] if not jitter else []
    if dt is None:
        E, nu, rho = mat[1][:3]
        lam = E * nu / ((1 + nu) * (1 - 2 * nu)); mu = E / (2 * (1 + nu))
        dt = 0.5 * h / math.sqrt((lam + 2 * mu) / rho)
    m.dt, m.nt = dt, nt
    if load_node is None:
        load_node = (nx // 2) + (nx + 1) * ny
    if series is None and nt > 0:
        f0 = 1.0 / (20.0 * dt)
        series = ricker(nt, dt, f0, 1.2 / f0)
    if series is not None:
        m.point_loads = [PointLoad(np.array([load_node], dtype=np.int32), np.asarray(load_dir, float),
                                   np.asarray(series, float))]
    m.rec_nodes = np.asarray(rec_nodes if rec_nodes is not None else [load_node], dtype=np.int32)
    return m.number_dofs()


def add_base_dashpots(m: Model, vs: float, vp: float, rho: float, h: float, th: float = 1.0) -> Model:
    """Lysmer-Kuhlemeyer base: every node of the bottom plane (z = min in 3-D, y = min in 2-D) gets a fixed twin
    node and one ZeroLength1D + Viscous1DLinear dashpot per direction, eta = rho * V * tributary area (Vp along
    the normal, Vs tangentially) -- what 01-Pre_Process/Method/Builder.py:1086-1131 emits.  The bottom plane must
    be unrestrained (fix=None).  Call before number_dofs-dependent data (loads, recorders) refer to new nodes;
    existing node / element indices are unchanged (new nodes and elements are appended)."""
    nd = m.ndim
    up = nd - 1
    zmin = m.coords[:, up].min()
    base = np.nonzero(np.abs(m.coords[:, up] - zmin) < 1e-9 * max(1.0, h))[0]
    lo = m.coords[base][:, :up].min(axis=0)
    hi = m.coords[base][:, :up].max(axis=0)
    n0 = m.n_nodes
    twins = np.arange(n0, n0 + len(base), dtype=np.int32)
    m.coords = np.vstack([m.coords, m.coords[base]])
    m.node_ndof = np.concatenate([m.node_ndof, np.full(len(base), nd, dtype=np.int32)])
    fd = np.asarray(m.freedof).reshape(-1)
    m.freedof = np.concatenate([fd, np.full(len(base) * nd, -1, dtype=np.int32)])
    mat_of: Dict[float, int] = {}
    conn, mats, attrs = [], [], []
    for b, t in zip(base, twins):
        on_edge = [(abs(m.coords[b, c] - lo[c]) < 1e-9 or abs(m.coords[b, c] - hi[c]) < 1e-9) for c in range(up)]
        area = (th if nd == 2 else 1.0) * float(np.prod([h * (0.5 if e else 1.0) for e in on_edge]))
        for d in range(nd):
            eta = rho * (vp if d == up else vs) * area
            if eta not in mat_of:
                mat_of[eta] = len(m.materials)
                m.materials = list(m.materials) + [(VISCOUS1DLINEAR, [eta])]
            row = np.zeros(8, dtype=np.int32)
            row[0], row[1] = t, b                 # fixed twin first, soil node second (Builder.py:1118)
            conn.append(row); mats.append(mat_of[eta]); a = np.zeros(10); a[0] = d; attrs.append(a)
    ne0 = m.n_elem
    m.elem_conn = np.vstack([m.elem_conn, np.array(conn, dtype=np.int32)])
    m.elem_kind = np.concatenate([m.elem_kind, np.full(len(conn), ZEROLENGTH1D, dtype=np.int32)])
    m.elem_mat = np.concatenate([m.elem_mat, np.array(mats, dtype=np.int32)])
    old_attr = m.elem_attr if m.elem_attr is not None else np.zeros((ne0, 10))
    m.elem_attr = np.vstack([old_attr, np.array(attrs)])
    if m.elem_am is not None:
        m.elem_am = np.concatenate([m.elem_am, np.zeros(len(conn))])
        m.elem_ak = np.concatenate([m.elem_ak, np.zeros(len(conn))])
    return m.number_dofs()


# -------------------------------------------------------------------------------
# reference JSON writer (Core/SeismoVLAB.py:49-298, Outputs.py:29-51)
# -------------------------------------------------------------------------------
def shuffle_numbering(m: Model, seed: int = 1, nodes: bool = True, elems: bool = True) -> Model:
    """The same mesh with node ids and element order permuted at random: what an unstructured mesher hands over.  The
    lattice hints are dropped (no Builder.py numbering any more).  Models with constraints are not handled."""
    if m.constraints:
        raise ValueError("shuffle_numbering: models with constraints are not handled")
    rng = np.random.default_rng(seed)
    n, ne = m.n_nodes, m.n_elem
    new_of_old = rng.permutation(n) if nodes else np.arange(n)
    old_of_new = np.argsort(new_of_old)
    eperm = rng.permutation(ne) if elems else np.arange(ne)              # new element q = old element eperm[q]
    enew_of_old = np.argsort(eperm)
    s = Model(ndim=m.ndim, lumped=m.lumped)
    s.coords = m.coords[old_of_new]
    s.node_ndof = m.node_ndof[old_of_new]
    fd = np.asarray(m.freedof_flat if hasattr(m, "freedof_flat") else m.freedof)
    ptr = np.concatenate([[0], np.cumsum(m.node_ndof)])
    s.freedof = [np.where(fd[ptr[o]:ptr[o + 1]] > -1, 0, fd[ptr[o]:ptr[o + 1]]).astype(np.int32) for o in old_of_new]
    s.materials = list(m.materials)
    npe = np.array([ELEM_NODES[int(k)] for k in m.elem_kind])
    conn = np.zeros_like(m.elem_conn)
    for q, e in enumerate(eperm):
        conn[q, :npe[e]] = new_of_old[m.elem_conn[e, :npe[e]]]
    s.elem_conn = conn
    s.elem_kind = m.elem_kind[eperm]; s.elem_mat = m.elem_mat[eperm]
    s.elem_attr = m.elem_attr[eperm] if m.elem_attr is not None else None
    s.elem_am = m.elem_am[eperm] if m.elem_am is not None else None
    s.elem_ak = m.elem_ak[eperm] if m.elem_ak is not None else None
    s.masses = [(int(new_of_old[q]), v) for q, v in m.masses]
    s.point_loads = [PointLoad(new_of_old[pl.nodes].astype(np.int32), pl.dir.copy(), pl.series.copy(), pl.factor) for pl in m.point_loads]
    s.supports = [(int(new_of_old[q]), d, sr, fc) for q, d, sr, fc in m.supports]
    if m.drm is not None:
        d = m.drm
        nn = new_of_old[d.nodes]
        o = np.argsort(nn)                                              # the reference keeps DRM nodes in ascending tag order
        s.drm = DRMLoad(elems=np.sort(enew_of_old[d.elems]).astype(np.int32), nodes=nn[o].astype(np.int32), exterior=d.exterior[o].copy(),
                        field=None if d.field is None else d.field[o].copy(), planewave=copy.deepcopy(d.planewave), factor=d.factor)
    s.dt, s.nt = m.dt, m.nt
    s.rec_nodes = new_of_old[m.rec_nodes].astype(np.int32) if m.rec_nodes is not None else None
    s.blocks = []
    return s.number_dofs()


def write_reference_json(m: Model, directory: str, name: str = "Model", combo: str = "Run",
                         resp=("disp",), integrator: str = "CENTRALDIFFERENCE", ndps: int = 16, newton=None,
                         _return_entities: bool = False, binary: bool = False) -> str:
    """Writes <directory>/Partition/<name>.1.0.json (+ load / .drm text files) in the schema the
    reference executable reads.  Tags are index+1.  Returns the partition directory.
    binary=True writes <name>.1.0.bin.json instead, with the Nodes / Elements / Constraints / Dampings tables in binary sidecars
    (the form pack_partition_tables produces), straight from the model's arrays: no per-entity dict is ever built, which is
    what makes 10^7-element models writable at all."""
    part = os.path.join(directory, "Partition")
    os.makedirs(part, exist_ok=True)
    # the reference's recorders write into <dir>/../Solution/<combo>/ and expect it to exist
    os.makedirs(os.path.join(directory, "Solution", combo), exist_ok=True)
    J: Dict[str, dict] = {}
    J["Global"] = {"ndim": m.ndim, "ntotal": m.n_total, "nfree": m.n_free, "update": "RESTARTABLE",
                   "massform": "LUMPED" if m.lumped else "CONSISTENT"}
    J["Materials"] = {}
    for i, (kind, par) in enumerate(m.materials):
        J["Materials"][str(i + 1)] = {"name": MAT_NAME[kind],
                                      "attributes": dict(zip(MAT_KEYS[kind], [float(p) for p in par]))}
    J["Nodes"] = {}
    for i in range(0 if binary else m.n_nodes):
        a, b = m.node_ptr[i], m.node_ptr[i + 1]
        J["Nodes"][str(i + 1)] = {"ndof": int(m.node_ndof[i]),
                                  "freedof": [int(v) for v in m.freedof_flat[a:b]],
                                  "totaldof": [int(v) for v in m.totaldof[a:b]],
                                  "coords": [float(v) for v in m.coords[i]]}
    if m.masses:
        J["Masses"] = {str(n + 1): {"ndof": len(v), "mass": [float(x) for x in v]} for n, v in m.masses}
    if m.constraints and not binary:
        J["Constraints"] = {str(t): {"stag": int(s), "mtag": [int(x) for x in mt], "factor": [float(x) for x in f]}
                            for t, s, mt, f in m.constraints}
    J["Elements"] = {}
    for e in range(0 if binary else m.n_elem):
        kind = int(m.elem_kind[e])
        nn = ELEM_NODES[kind]
        at = m.elem_attr[e] if m.elem_attr is not None else np.zeros(10)
        attr = {"material": int(m.elem_mat[e]) + 1, "rule": "GAUSS", "np": nn}
        if kind == ZEROLENGTH1D:               # Driver.hpp:1072-1078
            attr = {"material": int(m.elem_mat[e]) + 1, "dir": int(at[0])}
        elif kind == LIN2DQUAD4:
            attr["th"] = float(at[0])
        elif kind == PML3DHEXA8:
            attr.update({"n": float(at[0]), "L": float(at[1]), "R": float(at[2]),
                         "x0": [float(v) for v in at[3:6]], "npml": [float(v) for v in at[6:9]]})
        elif kind == PML2DQUAD4:
            attr.update({"th": float(at[0]), "n": float(at[1]), "L": float(at[2]), "R": float(at[3]),
                         "x0": [float(v) for v in at[4:6]], "npml": [float(v) for v in at[6:8]]})
        J["Elements"][str(e + 1)] = {"name": ELEM_NAME[kind],
                                     "conn": [int(v) + 1 for v in m.elem_conn[e, :nn]], "attributes": attr}
    # dampings: FREE on everything unless Rayleigh given (SeismoVLAB.py:704-707)
    if binary:
        stem = os.path.join(part, f"{name}.1.0")
        if not _return_entities:                                      # write_reference_partitions writes each rank's share itself
            _write_binary_tables(m, stem)
        J["Nodes"] = {"binary": os.path.basename(stem + ".nodes.bin"), "count": m.n_nodes}
        J["Elements"] = {"binary": os.path.basename(stem + ".elems.bin"), "count": m.n_elem}
        J["Dampings"] = {"binary": os.path.basename(stem + ".elems.bin")}
        if m.constraints:
            J["Constraints"] = {"binary": os.path.basename(stem + ".cons.bin"), "count": len(m.constraints)}
    elif m.elem_am is not None and (np.any(m.elem_am != 0) or np.any(m.elem_ak != 0)):
        groups: Dict[tuple, list] = {}
        for e in range(m.n_elem):
            groups.setdefault((float(m.elem_am[e]), float(m.elem_ak[e])), []).append(e + 1)
        J["Dampings"] = {}
        for i, ((am, ak), lst) in enumerate(groups.items()):
            if am == 0.0 and ak == 0.0:
                J["Dampings"][str(i + 1)] = {"name": "FREE", "attributes": {"list": lst}}
            else:
                J["Dampings"][str(i + 1)] = {"name": "RAYLEIGH", "attributes": {"am": am, "ak": ak, "list": lst}}
    else:
        J["Dampings"] = {"1": {"name": "FREE", "attributes": {"list": list(range(1, m.n_elem + 1))}}}
    J["Loads"] = {}
    tag = 0
    for pl in m.point_loads:
        tag += 1
        if len(pl.series) == 1:
            J["Loads"][str(tag)] = {"name": "POINTLOAD", "attributes": {
                "name": "CONSTANT", "type": "CONCENTRATED", "mag": float(pl.series[0]),
                "dir": [float(v) for v in pl.dir], "list": [int(v) + 1 for v in pl.nodes]}}
        else:
            fn = os.path.join(part, f"{name}_load{tag}.txt")
            with open(fn, "w") as f:
                f.write(f"{len(pl.series)}\n" + "\n".join(repr(float(v)) for v in pl.series) + "\n")
            J["Loads"][str(tag)] = {"name": "POINTLOAD", "attributes": {
                "name": "TIMESERIES", "type": "CONCENTRATED", "file": fn,
                "dir": [float(v) for v in pl.dir], "list": [int(v) + 1 for v in pl.nodes]}}
    factors = [float(pl.factor) for pl in m.point_loads]
    if m.drm is not None and m.drm.field is not None:
        tag += 1
        d = m.drm
        drmdir = os.path.join(part, "DRM")
        os.makedirs(drmdir, exist_ok=True)
        for li, node in enumerate(d.nodes):
            with open(os.path.join(drmdir, f"{name}-{tag}.{int(node) + 1}.drm"), "w") as f:
                fld = d.field[li]
                f.write(f"{fld.shape[0]} {fld.shape[1]} {int(d.exterior[li])}\n")
                for row in fld:
                    f.write(" ".join(repr(float(v)) for v in row) + "\n")
        J["Loads"][str(tag)] = {"name": "ELEMENTLOAD", "attributes": {
            "name": "TIMESERIES", "type": "GENERALWAVE",
            "file": os.path.join(drmdir, f"{name}-{tag}.$.drm"),
            "list": [int(v) + 1 for v in d.elems]}}
        factors.append(float(d.factor))
    if m.supports:
        # Supports{node tag: {type, file | value, dof (0-based)}} + one SUPPORTMOTION load per distinct factor
        by_node: Dict[int, list] = {}
        for node, dof, series, fac in m.supports:
            by_node.setdefault(int(node), []).append((int(dof), np.asarray(series, float), float(fac)))
        J["Supports"] = {}
        for node, lst in sorted(by_node.items()):
            if all(len(q[1]) == 1 for q in lst):
                J["Supports"][str(node + 1)] = {"type": "CONSTANT", "value": [float(q[1][0]) for q in lst], "dof": [q[0] for q in lst]}
            else:
                files = []
                for dof, series, _ in lst:
                    fn = os.path.join(part, f"{name}_support{node + 1}_{dof}.txt")
                    with open(fn, "w") as f:
                        f.write(f"{len(series)}\n" + "\n".join(repr(float(v)) for v in series) + "\n")
                    files.append(fn)
                J["Supports"][str(node + 1)] = {"type": "TIMESERIES", "file": files, "dof": [q[0] for q in lst]}
        for fac in sorted({q[2] for lst in by_node.values() for q in lst}):
            tag += 1
            nodes = sorted(n for n, lst in by_node.items() if any(q[2] == fac for q in lst))
            if any(q[2] != fac for n in nodes for q in by_node[n]):
                raise ValueError("write_reference_json: one combination factor per support node")
            J["Loads"][str(tag)] = {"name": "SUPPORTMOTION", "attributes": {"list": [n + 1 for n in nodes]}}
            factors.append(float(fac))
    J["Combinations"] = {"1": {"name": combo, "attributes": {
        "folder": combo, "load": list(range(1, tag + 1)), "factor": factors}}}
    J["Recorders"] = {}
    for i, r in enumerate(resp):
        J["Recorders"][str(i + 1)] = {"name": "NODE", "file": f"{r}.0.out", "ndps": ndps, "resp": r,
                                      "list": [int(v) + 1 for v in m.rec_nodes], "nsamp": 1}
    J["Simulations"] = {"combo": 1, "attributes": {
        "analysis": {"name": "DYNAMIC", "nt": int(m.nt)},
        "algorithm": ({"name": "LINEAR", "nstep": 1} if newton is None else
                      {"name": "NEWTON", "cnvgtol": float(newton[0]), "nstep": int(newton[1]), "cnvgtest": int(newton[2])}),
        "integrator": {"name": integrator, "dt": float(m.dt), "ktol": 1e-12, "mtol": 1e-12, "ftol": 1e-12},
        "solver": {"name": "EIGEN", "update": 1 if newton is None else 0}}}
    if _return_entities:
        return part, J
    with open(os.path.join(part, f"{name}.1.0.bin.json" if binary else f"{name}.1.0.json"), "w") as f:
        json.dump(J, f, indent=4)
    return part


def _write_binary_tables(m: Model, stem: str, nodes=None, elems=None, cons=None):
    """the sidecars of pack_partition_tables written from the model's arrays (tags = index + 1), vectorised.  nodes / elems /
    cons select a partition's share (ascending node / element indices, constraint list); numbering stays the model's."""
    nsel = np.arange(m.n_nodes) if nodes is None else np.asarray(nodes, dtype=np.int64)
    esel = np.arange(m.n_elem) if elems is None else np.asarray(elems, dtype=np.int64)
    cons = m.constraints if cons is None else cons
    n, ne = len(nsel), len(esel)
    ndof = np.asarray(m.node_ndof, "<i4")[nsel]
    if nodes is None:
        dofs = np.arange(m.n_total)
    else:                                                             # dofs of the selected nodes, node-major
        start = np.asarray(m.node_ptr)[nsel]
        dofs = np.repeat(start - np.concatenate([[0], np.cumsum(ndof)[:-1]]), ndof) + np.arange(int(ndof.sum()))
    with open(stem + ".nodes.bin", "wb") as f:
        f.write(b"SVLN" + np.array([1], "<u4").tobytes() + np.array([n], "<u8").tobytes() + np.array([m.ndim], "<u4").tobytes())
        for a in ((nsel + 1).astype("<u4"), ndof, np.asarray(m.coords, "<f8")[nsel],
                  np.asarray(m.totaldof, "<i4")[dofs], np.asarray(m.freedof_flat, "<i4")[dofs]):
            f.write(np.ascontiguousarray(a).tobytes())
    kind = np.asarray(m.elem_kind, "<i4")[esel]
    nconn = np.array([0] + [ELEM_NODES[k] for k in (1, 2, 3, 4, 5)], "<i4")[kind]
    conn = np.zeros((ne, 8), "<u4")
    econn = np.asarray(m.elem_conn)[esel]
    for npe in np.unique(nconn):
        sel = nconn == npe
        conn[sel, :npe] = econn[sel, :npe] + 1
    attr = np.zeros((ne, 10), "<f8") if m.elem_attr is None else np.asarray(m.elem_attr, "<f8")[esel].copy()
    nattr = np.array([0, 0, 1, 9, 8, 1])[kind]                       # meaningful leading entries per kind
    attr[np.arange(10)[None, :] >= nattr[:, None]] = 0.0
    am = np.zeros(ne, "<f8") if m.elem_am is None else np.asarray(m.elem_am, "<f8")[esel]
    ak = np.zeros(ne, "<f8") if m.elem_ak is None else np.asarray(m.elem_ak, "<f8")[esel]
    ray = ((am != 0) | (ak != 0)).astype("u1")
    with open(stem + ".elems.bin", "wb") as f:
        f.write(b"SVLE" + np.array([1], "<u4").tobytes() + np.array([ne], "<u8").tobytes())
        for a in ((esel + 1).astype("<u4"), kind, (np.asarray(m.elem_mat)[esel] + 1).astype("<u4"), nconn, conn, attr, am, ak, ray):
            f.write(np.ascontiguousarray(a).tobytes())
    if cons:
        if any(len(c[2]) != 1 for c in cons):
            raise ValueError("binary tables hold single-master (EQUAL) constraints only")
        with open(stem + ".cons.bin", "wb") as f:
            f.write(b"SVLC" + np.array([1], "<u4").tobytes() + np.array([len(cons)], "<u8").tobytes())
            f.write(np.array([c[0] for c in cons], "<i8").tobytes())
            f.write(np.array([c[1] for c in cons], "<i4").tobytes())
            f.write(np.array([c[2][0] for c in cons], "<i4").tobytes())
            f.write(np.array([c[3][0] for c in cons], "<f8").tobytes())


def write_reference_partitions(m: Model, epart, nparts: int, directory: str, name: str = "Model", combo: str = "Run",
                               resp=("disp",), integrator: str = "CENTRALDIFFERENCE", ndps: int = 16, binary: bool = False) -> str:
    """binary=True: <name>.1.<rank>.bin.json with the big tables in sidecars, written from the arrays (no per-entity dicts).
    One JSON file per rank, <directory>/Partition/<name>.1.<rank>.json, the way the reference pre-processor splits a model
    after METIS (Core/SeismoVLAB.py:300-420 createPartitions, :49-298 Entities2Processor): GLOBAL tags and GLOBAL total /
    free dof numbers everywhere (`Global.ntotal / nfree` are the whole model's), a partition holds the nodes of its elements
    plus the master nodes of the constraints whose slave dofs it holds (:381-392), nodal masses and the nodes of a point load
    go to the first partition that holds them (:117-122, :163-180), NODE recorders list the partition's own nodes and write
    `<resp>.<rank>.out` (:237-246), combinations keep the loads present in the partition.  `epart[e]` is the rank of element e
    (elements beyond len(epart), e.g. dashpots, follow partition.split_model's rule).  Returns the partition directory."""
    from . import partition as P
    part, J = write_reference_json(m, directory, name, combo, resp, integrator, ndps, _return_entities=True, binary=binary)
    subs = P.split_model(m, epart, nparts, tie_closure="slave")       # element / node sets only; numbering stays global here
    taken_mass, taken_load = set(), {k: set() for k in J["Loads"]}
    fd_all = np.asarray(m.freedof_flat)
    for r, s in enumerate(subs):
        ntags = [int(n) + 1 for n in s.global_nodes]
        nset = set(ntags)
        etags = [int(e) + 1 for e in s.global_elems]
        eset = set(etags)
        K = {"Global": J["Global"], "Materials": J["Materials"]}
        masses = {t: v for t, v in J.get("Masses", {}).items() if int(t) in nset and t not in taken_mass}
        taken_mass.update(masses)
        if binary:
            # constraints whose slave dof sits on a node of this partition, in the model's order
            held = np.zeros(m.n_nodes, dtype=bool)
            held[s.global_nodes] = True
            node_of_total = np.repeat(np.arange(m.n_nodes), np.diff(m.node_ptr))
            mine = [c for c in m.constraints if held[node_of_total[c[1]]]]
            stem = os.path.join(part, f"{name}.1.{r}")
            _write_binary_tables(m, stem, nodes=s.global_nodes, elems=s.global_elems, cons=mine)
            K["Nodes"] = {"binary": os.path.basename(stem + ".nodes.bin"), "count": len(ntags)}
            if masses:
                K["Masses"] = masses
            if mine:
                K["Constraints"] = {"binary": os.path.basename(stem + ".cons.bin"), "count": len(mine)}
            K["Elements"] = {"binary": os.path.basename(stem + ".elems.bin"), "count": len(etags)}
            K["Dampings"] = {"binary": os.path.basename(stem + ".elems.bin")}
        else:
            K["Nodes"] = {str(t): J["Nodes"][str(t)] for t in ntags}
            if masses:
                K["Masses"] = masses
            cons = {}
            for t in ntags:
                for f in J["Nodes"][str(t)]["freedof"]:
                    if f < -1:
                        cons[str(f)] = J["Constraints"][str(f)]
            if cons:
                K["Constraints"] = cons
            K["Elements"] = {str(t): J["Elements"][str(t)] for t in etags}
            K["Dampings"] = {}
            for d, D in J["Dampings"].items():
                lst = [t for t in D["attributes"]["list"] if t in eset]
                if lst:
                    K["Dampings"][d] = {"name": D["name"], "attributes": dict(D["attributes"], list=lst)}
        loads = {}
        for l, L in J["Loads"].items():
            if L["name"] == "POINTLOAD":
                lst = sorted(t for t in L["attributes"]["list"] if t in nset and t not in taken_load[l])
                taken_load[l].update(lst)
            elif L["name"] == "SUPPORTMOTION":          # every partition that holds the node moves it (SeismoVLAB.py:211-216)
                lst = sorted(t for t in L["attributes"]["list"] if t in nset)
            else:
                lst = sorted(t for t in L["attributes"]["list"] if t in eset)
            if lst:
                loads[l] = {"name": L["name"], "attributes": dict(L["attributes"], list=lst)}
        if loads:
            K["Loads"] = loads
        sup = {t: v for t, v in J.get("Supports", {}).items() if int(t) in nset}      # SeismoVLAB.py:124-128
        if sup:
            K["Supports"] = sup
        C = J["Combinations"]["1"]
        keep = [(l, f) for l, f in zip(C["attributes"]["load"], C["attributes"]["factor"]) if str(l) in loads]
        K["Combinations"] = {"1": {"name": C["name"], "attributes": (
            {"folder": C["attributes"]["folder"], "load": [l for l, _ in keep], "factor": [f for _, f in keep]} if keep else {})}}
        recs = {}
        for q, R in J["Recorders"].items():
            lst = sorted(t for t in R["list"] if t in nset)
            if lst:
                recs[q] = dict(R, file=R["file"].replace(".0.out", f".{r}.out"), list=lst)
        if recs:
            K["Recorders"] = recs
        K["Simulations"] = J["Simulations"]
        with open(os.path.join(part, f"{name}.1.{r}.bin.json" if binary else f"{name}.1.{r}.json"), "w") as f:
            json.dump(K, f, indent=4)
    return part


# -------------------------------------------------------------------------------
# binary sidecars for the big tables of a partition file
# -------------------------------------------------------------------------------
# The per-rank JSON spends > 99 % of its bytes on "Nodes" and "Elements" (one dict per entity, indent=4): ~400 B per node
# and ~300 B per element, i.e. tens of GB and minutes of parsing at 10^8 DOF.  pack_partition_tables moves these tables (and
# the constraint / damping lists that scale with them) into flat little-endian arrays next to the JSON; everything else
# (materials, loads, combinations, recorders, simulation) stays JSON, so the file remains the reference's schema with
# three keys replaced by {"binary": <file>, "count": n}.  Both readers (load_partition_json here, UpdateMesh in
# svl_host.cpp) accept either form.
#   <stem>.nodes.bin  "SVLN" u32 version u64 n u32 ncoord | tag u32[n] ndof i32[n] coords f64[n*ncoord] total i32[S] free i32[S]
#   <stem>.elems.bin  "SVLE" u32 version u64 n | tag u32[n] kind i32[n] material-tag u32[n] nconn i32[n] conn u32[n*8]
#                     attr f64[n*10] (order of read_reference_json) am f64[n] ak f64[n] rayleigh u8[n]
#   <stem>.cons.bin   "SVLC" u32 version u64 n | tag i64[n] slave-total i32[n] master-free i32[n] factor f64[n]
_ELEM_KIND_OF = {v: k for k, v in ELEM_NAME.items()}


def _elem_attr_row(kind, a):
    row = np.zeros(10)
    if kind == LIN2DQUAD4:
        row[0] = float(a.get("th", 1.0))
    elif kind == ZEROLENGTH1D:
        row[0] = int(a["dir"])
    elif kind == PML3DHEXA8:
        row[:9] = [a["n"], a["L"], a["R"], *a["x0"], *a["npml"]]
    elif kind == PML2DQUAD4:
        row[:8] = [a.get("th", 1.0), a["n"], a["L"], a["R"], *a["x0"], *a["npml"]]
    return row


def _elem_attr_dict(kind, mat_tag, row):
    if kind == ZEROLENGTH1D:
        return {"material": int(mat_tag), "dir": int(row[0])}
    a = {"material": int(mat_tag), "rule": "GAUSS", "np": ELEM_NODES[kind]}
    if kind == LIN2DQUAD4:
        a["th"] = float(row[0])
    elif kind == PML3DHEXA8:
        a.update({"n": float(row[0]), "L": float(row[1]), "R": float(row[2]), "x0": [float(v) for v in row[3:6]],
                  "npml": [float(v) for v in row[6:9]]})
    elif kind == PML2DQUAD4:
        a.update({"th": float(row[0]), "n": float(row[1]), "L": float(row[2]), "R": float(row[3]),
                  "x0": [float(v) for v in row[4:6]], "npml": [float(v) for v in row[6:8]]})
    return a


def pack_partition_tables(json_path: str, out_path: Optional[str] = None) -> str:
    """Rewrites a per-rank partition JSON (from the reference's pre-processor or from this module) with its Nodes / Elements /
    Constraints / Dampings tables moved into binary sidecars.  Returns the path of the new JSON (default <stem>.bin.json ->
    replace the '.json' of the -file pattern by '.bin.json')."""
    with open(json_path) as f:
        J = json.load(f)
    stem = json_path[:-5] if json_path.endswith(".json") else json_path
    out_path = out_path or stem + ".bin.json"
    ntags = sorted(J["Nodes"], key=int)
    n = len(ntags)
    ncoord = len(J["Nodes"][ntags[0]]["coords"]) if n else 0
    ndof = np.array([int(J["Nodes"][t]["ndof"]) for t in ntags], dtype="<i4")
    with open(stem + ".nodes.bin", "wb") as f:
        f.write(b"SVLN" + np.array([1], "<u4").tobytes() + np.array([n], "<u8").tobytes() + np.array([ncoord], "<u4").tobytes())
        f.write(np.array([int(t) for t in ntags], dtype="<u4").tobytes())
        f.write(ndof.tobytes())
        f.write(np.array([J["Nodes"][t]["coords"] for t in ntags], dtype="<f8").tobytes())
        f.write(np.array([v for t in ntags for v in J["Nodes"][t]["totaldof"]], dtype="<i4").tobytes())
        f.write(np.array([v for t in ntags for v in J["Nodes"][t]["freedof"]], dtype="<i4").tobytes())
    etags = sorted(J["Elements"], key=int)
    ne = len(etags)
    eidx = {int(t): i for i, t in enumerate(etags)}
    kind = np.zeros(ne, "<i4"); mat = np.zeros(ne, "<u4"); nconn = np.zeros(ne, "<i4"); conn = np.zeros((ne, 8), "<u4")
    attr = np.zeros((ne, 10), "<f8"); am = np.zeros(ne, "<f8"); ak = np.zeros(ne, "<f8"); ray = np.zeros(ne, "u1")
    for i, t in enumerate(etags):
        E = J["Elements"][t]
        if E["name"].upper() not in _ELEM_KIND_OF:
            raise ValueError(f"pack_partition_tables: element {E['name']} is not on this path")
        kind[i] = _ELEM_KIND_OF[E["name"].upper()]
        mat[i] = int(E["attributes"]["material"])
        nconn[i] = len(E["conn"])
        conn[i, :nconn[i]] = E["conn"]
        attr[i] = _elem_attr_row(int(kind[i]), E["attributes"])
    for D in J.get("Dampings", {}).values():
        if D["name"].upper() == "RAYLEIGH":
            for t in D["attributes"]["list"]:
                am[eidx[int(t)]] = float(D["attributes"]["am"]); ak[eidx[int(t)]] = float(D["attributes"]["ak"]); ray[eidx[int(t)]] = 1
        elif D["name"].upper() != "FREE":
            raise ValueError(f"pack_partition_tables: damping {D['name']} is not on this path")
    with open(stem + ".elems.bin", "wb") as f:
        f.write(b"SVLE" + np.array([1], "<u4").tobytes() + np.array([ne], "<u8").tobytes())
        for a in (np.array([int(t) for t in etags], dtype="<u4"), kind, mat, nconn, conn, attr, am, ak, ray):
            f.write(a.tobytes())
    K = dict(J)
    K["Nodes"] = {"binary": os.path.basename(stem + ".nodes.bin"), "count": n}
    K["Elements"] = {"binary": os.path.basename(stem + ".elems.bin"), "count": ne}
    K["Dampings"] = {"binary": os.path.basename(stem + ".elems.bin")}
    cons = J.get("Constraints", {})
    if cons and all(len(c["mtag"]) == 1 for c in cons.values()):
        ct = list(cons)                                            # the file's own order
        with open(stem + ".cons.bin", "wb") as f:
            f.write(b"SVLC" + np.array([1], "<u4").tobytes() + np.array([len(ct)], "<u8").tobytes())
            f.write(np.array([int(t) for t in ct], "<i8").tobytes())
            f.write(np.array([cons[t]["stag"] for t in ct], "<i4").tobytes())
            f.write(np.array([cons[t]["mtag"][0] for t in ct], "<i4").tobytes())
            f.write(np.array([cons[t]["factor"][0] for t in ct], "<f8").tobytes())
        K["Constraints"] = {"binary": os.path.basename(stem + ".cons.bin"), "count": len(ct)}
    with open(out_path, "w") as f:
        json.dump(K, f, indent=4)
    return out_path


def load_partition_json(path: str) -> dict:
    """json.load of a partition file with binary sidecars (pack_partition_tables) expanded back into the reference's
    dict-per-entity schema; a plain file is returned as it is."""
    with open(path) as f:
        J = json.load(f)
    here = os.path.dirname(os.path.abspath(path))

    def rd(f, dt, count):
        return np.frombuffer(f.read(np.dtype(dt).itemsize * count), dtype=dt, count=count)

    if isinstance(J.get("Nodes"), dict) and "binary" in J["Nodes"]:
        with open(os.path.join(here, J["Nodes"]["binary"]), "rb") as f:
            assert f.read(4) == b"SVLN" and rd(f, "<u4", 1)[0] == 1
            n = int(rd(f, "<u8", 1)[0]); nc = int(rd(f, "<u4", 1)[0])
            tags = rd(f, "<u4", n); ndof = rd(f, "<i4", n); xyz = rd(f, "<f8", n * nc).reshape(n, nc)
            S = int(ndof.sum())
            tot = rd(f, "<i4", S); fre = rd(f, "<i4", S)
        ptr = np.concatenate([[0], np.cumsum(ndof)])
        J["Nodes"] = {str(int(tags[i])): {"ndof": int(ndof[i]), "freedof": [int(v) for v in fre[ptr[i]:ptr[i + 1]]],
                                          "totaldof": [int(v) for v in tot[ptr[i]:ptr[i + 1]]],
                                          "coords": [float(v) for v in xyz[i]]} for i in range(n)}
    if isinstance(J.get("Elements"), dict) and "binary" in J["Elements"]:
        with open(os.path.join(here, J["Elements"]["binary"]), "rb") as f:
            assert f.read(4) == b"SVLE" and rd(f, "<u4", 1)[0] == 1
            ne = int(rd(f, "<u8", 1)[0])
            tags = rd(f, "<u4", ne); kind = rd(f, "<i4", ne); mat = rd(f, "<u4", ne); nconn = rd(f, "<i4", ne)
            conn = rd(f, "<u4", ne * 8).reshape(ne, 8); attr = rd(f, "<f8", ne * 10).reshape(ne, 10)
            am = rd(f, "<f8", ne); ak = rd(f, "<f8", ne); ray = rd(f, "u1", ne)
        J["Elements"] = {str(int(tags[i])): {"name": ELEM_NAME[int(kind[i])], "conn": [int(v) for v in conn[i, :nconn[i]]],
                                             "attributes": _elem_attr_dict(int(kind[i]), mat[i], attr[i])} for i in range(ne)}
        groups: Dict[tuple, list] = {}
        for i in range(ne):
            groups.setdefault((bool(ray[i]), float(am[i]), float(ak[i])), []).append(int(tags[i]))
        J["Dampings"] = {str(q + 1): ({"name": "RAYLEIGH", "attributes": {"am": a_, "ak": k_, "list": lst}} if r_ else
                                      {"name": "FREE", "attributes": {"list": lst}})
                         for q, ((r_, a_, k_), lst) in enumerate(groups.items())}
    if isinstance(J.get("Constraints"), dict) and "binary" in J["Constraints"]:
        with open(os.path.join(here, J["Constraints"]["binary"]), "rb") as f:
            assert f.read(4) == b"SVLC" and rd(f, "<u4", 1)[0] == 1
            nc_ = int(rd(f, "<u8", 1)[0])
            tg = rd(f, "<i8", nc_); st = rd(f, "<i4", nc_); mt = rd(f, "<i4", nc_); fc = rd(f, "<f8", nc_)
        J["Constraints"] = {str(int(tg[i])): {"stag": int(st[i]), "mtag": [int(mt[i])], "factor": [float(fc[i])]} for i in range(nc_)}
    return J


def read_reference_json(path: str, base_dir: Optional[str] = None) -> Model:
    """Inverse of write_reference_json for the subset of the schema on this path (SURVEY.md App. D;
    12-Utilities/Driver.hpp:1981-2046): builds a Model from a per-rank JSON file the reference's pre-processor wrote.
    Entities are taken in ascending tag order; node / element indices are positions in that order.  Relative load
    files are resolved against `base_dir` (default: the directory the run is started from = parent of Partition/).
    Adds m.integrator (the JSON's integrator name) and m.rec_spec = [(resp, [node indices])]."""
    J = load_partition_json(path)                                 # plain JSON, or JSON + binary sidecars
    base = base_dir or os.path.dirname(os.path.dirname(os.path.abspath(path)))
    G = J["Global"]
    m = Model(ndim=int(G["ndim"]), lumped=str(G.get("massform", "LUMPED")).upper() == "LUMPED")
    ntags = sorted(J["Nodes"], key=int)
    nidx = {int(t): i for i, t in enumerate(ntags)}
    m.coords = np.array([[float(v) for v in J["Nodes"][t]["coords"]][:m.ndim] for t in ntags])
    m.node_ndof = np.array([int(J["Nodes"][t]["ndof"]) for t in ntags], dtype=np.int32)
    # the file's own total / free numbering (any scheme of 01-Pre_Process/Core/Numberer.py) is mapped to the Model's
    # node-major numbering: constraints refer to slave TOTAL and master FREE dofs of the file (Driver.hpp:440-505)
    fd, tot_map, free_map, nfree = [], {}, {}, 0
    ntot = 0
    for t in ntags:
        N = J["Nodes"][t]
        row = []
        for ft, tt in zip(N["freedof"], N["totaldof"]):
            tot_map[int(tt)] = ntot; ntot += 1
            if int(ft) > -1:
                free_map[int(ft)] = nfree; nfree += 1
                row.append(0)
            else:
                row.append(int(ft))                                  # -1 restrained, < -1 constraint tag
        fd.append(np.array(row, dtype=np.int32))
    m.freedof = fd
    for ctag, Cn in J.get("Constraints", {}).items():
        m.constraints.append((int(ctag), tot_map[int(Cn["stag"])], [free_map[int(v)] for v in Cn["mtag"]],
                              [float(v) for v in Cn["factor"]]))
    kinds = {v: k for k, v in MAT_NAME.items()}
    mtags = sorted(J["Materials"], key=int)
    midx = {int(t): i for i, t in enumerate(mtags)}
    for t in mtags:
        M_ = J["Materials"][t]
        kind = kinds[M_["name"].upper()]
        m.materials.append((kind, [float(M_["attributes"][k]) for k in MAT_KEYS[kind]]))
    ekinds = {v: k for k, v in ELEM_NAME.items()}
    etags = sorted(J["Elements"], key=int)
    eidx = {int(t): i for i, t in enumerate(etags)}
    ne = len(etags)
    m.elem_kind = np.zeros(ne, dtype=np.int32); m.elem_conn = np.zeros((ne, 8), dtype=np.int32)
    m.elem_mat = np.zeros(ne, dtype=np.int32); m.elem_attr = np.zeros((ne, 10))
    for i, t in enumerate(etags):
        E = J["Elements"][t]
        kind = ekinds[E["name"].upper()]
        a = E["attributes"]
        m.elem_kind[i] = kind
        m.elem_conn[i, :len(E["conn"])] = [nidx[int(n)] for n in E["conn"]]
        m.elem_mat[i] = midx[int(a["material"])]
        if kind == LIN2DQUAD4:
            m.elem_attr[i, 0] = float(a.get("th", 1.0))
        elif kind == ZEROLENGTH1D:
            m.elem_attr[i, 0] = int(a["dir"])
        elif kind == PML3DHEXA8:
            m.elem_attr[i, :9] = [a["n"], a["L"], a["R"], *a["x0"], *a["npml"]]
        elif kind == PML2DQUAD4:
            m.elem_attr[i, :8] = [a.get("th", 1.0), a["n"], a["L"], a["R"], *a["x0"], *a["npml"]]
    m.elem_am = np.zeros(ne); m.elem_ak = np.zeros(ne)
    for D in J.get("Dampings", {}).values():
        if D["name"].upper() == "RAYLEIGH":
            for t in D["attributes"]["list"]:
                if m.elem_kind[eidx[int(t)]] != ZEROLENGTH1D:          # ZeroLength1D::SetDamping does nothing
                    m.elem_am[eidx[int(t)]] = float(D["attributes"]["am"])
                    m.elem_ak[eidx[int(t)]] = float(D["attributes"]["ak"])
    if not m.elem_am.any() and not m.elem_ak.any():
        m.elem_am = m.elem_ak = None
    for t, v in J.get("Masses", {}).items():
        m.masses.append((nidx[int(t)], [float(x) for x in v["mass"]]))
    sim = J["Simulations"]
    combo = J["Combinations"][str(sim["combo"])]["attributes"]
    for ltag, factor in zip(combo.get("load", []), combo.get("factor", [])):
        L = J["Loads"][str(ltag)]
        a = L["attributes"]
        if L["name"].upper() == "ELEMENTLOAD" and a["type"].upper() == "GENERALWAVE":
            # DRM: one `.drm` text file per node of the listed elements, `nt nFields cond` then nt rows (Driver.hpp:1689-1721);
            # '$' in the file pattern -> node tag; nodes in ascending tag order (the reference's std::map)
            if m.drm is not None:
                raise ValueError("read_reference_json: one GENERALWAVE load per combination")
            elems = np.array([eidx[int(t)] for t in a["list"]], dtype=np.int32)
            etag_of = {i: t for t, i in eidx.items()}
            tags = sorted({int(n) for e in elems for n in J["Elements"][str(etag_of[int(e)])]["conn"]})
            fields, ext = [], []
            for t in tags:
                fn = a["file"].replace("$", str(t))
                fn = fn if os.path.isabs(fn) else os.path.join(base, fn)
                tok = open(fn).read().split()
                nt_, nf_, cond = int(tok[0]), int(tok[1]), int(tok[2])
                fields.append(np.array([float(v) for v in tok[3:3 + nt_ * nf_]]).reshape(nt_, nf_))
                ext.append(1 if cond else 0)
            m.drm = DRMLoad(elems=elems, nodes=np.array([nidx[t] for t in tags], dtype=np.int32), exterior=np.array(ext, dtype=np.uint8),
                            field=np.stack(fields), factor=float(factor))
            continue
        if L["name"].upper() == "SUPPORTMOTION":
            # Driver.hpp:1725-1735 + :509-563: the load lists nodes, their motions sit in Supports{node tag}
            for ntag in a["list"]:
                S = J.get("Supports", {}).get(str(int(ntag)))
                if S is None:
                    continue
                for q, dof in enumerate(S["dof"]):
                    if S["type"].upper() == "CONSTANT":
                        series = np.array([float(S["value"][q])])
                    else:
                        fn = S["file"][q] if os.path.isabs(S["file"][q]) else os.path.join(base, S["file"][q])
                        tok = open(fn).read().split()
                        series = np.array([float(v) for v in tok[1:1 + int(tok[0])]])
                    m.supports.append((nidx[int(ntag)], int(dof), series, float(factor)))
            continue
        if L["name"].upper() != "POINTLOAD" or a["type"].upper() != "CONCENTRATED":
            raise ValueError("read_reference_json: only CONCENTRATED point loads and GENERALWAVE element loads are handled by this reader")
        if a["name"].upper() == "CONSTANT":
            series = np.array([float(a["mag"])])
        else:
            fn = a["file"] if os.path.isabs(a["file"]) else os.path.join(base, a["file"])
            tok = open(fn).read().split()                           # Driver.hpp:1514-1527: count, then the values
            series = np.array([float(v) for v in tok[1:1 + int(tok[0])]])
        d = np.zeros(3); d[:len(a["dir"])] = a["dir"]
        m.point_loads.append(PointLoad(np.array([nidx[int(n)] for n in a["list"]], dtype=np.int32), d[:max(m.ndim, len(a["dir"]))][:3],
                                       series, float(factor)))
    A = sim["attributes"]
    m.dt, m.nt = float(A["integrator"]["dt"]), int(A["analysis"]["nt"])
    m.integrator = A["integrator"]["name"].upper()
    alg = A.get("algorithm", {})
    m.newton = ((float(alg.get("cnvgtol", 1e-6)), int(alg.get("nstep", 1)), int(alg.get("cnvgtest", 4)))
                if str(alg.get("name", "LINEAR")).upper() == "NEWTON" else None)
    m.rec_spec = []
    for r in sorted(J.get("Recorders", {}), key=int):
        R = J["Recorders"][r]
        if R["name"].upper() == "NODE":
            m.rec_spec.append((R["resp"].lower(), [nidx[int(n)] for n in R["list"]]))
    m.rec_nodes = np.array(m.rec_spec[0][1] if m.rec_spec else [0], dtype=np.int32)
    m.blocks = []
    return m.number_dofs()


def read_node_recorder(path: str) -> np.ndarray:
    """NODE recorder text file (Recorder.cpp:73-105, 239-269) -> [nrows, ncols] array."""
    with open(path) as f:
        first = f.readline().split()
        nn = int(first[0])
        for _ in range(nn):
            f.readline()
        return np.loadtxt(f, ndmin=2)


# -------------------------------------------------------------------------------
# DRM (Method/Builder.py:1010-1073 setDRMDomain; Core/PlaneWave.py:272-318 file layout)
# -------------------------------------------------------------------------------
def ricker_uva(tau, f0):
    """Ricker displacement pulse and its first two time derivatives."""
    w = math.pi * f0
    b = (w * tau) ** 2
    e = np.exp(-b)
    u = (1.0 - 2.0 * b) * e
    v = (-6.0 * w * w * tau + 4.0 * w ** 4 * tau ** 3) * e
    a = (-6.0 * w * w + 24.0 * w ** 4 * tau ** 2 - 8.0 * w ** 6 * tau ** 4) * e
    return u, v, a


def add_drm_box(m: Model, x0, xl, planewave: dict, tabulate_nt: int = 0, factor: float = 1.0) -> Model:
    """DRM element layer = elements cut by the box |x - x0| <= xl (setDRMDomain); nodes inside the
    box are 'interior/boundary' (cond 0), outside 'exterior' (cond 1).  planewave = {dir, pol, xref,
    c, f0, t0, amp}: u(x,t) = amp pol ricker(t - t0 - (x - xref).dir / c).  tabulate_nt > 0 also
    tabulates the [nnodes, nt, 3 ndim] field exactly as the reference's .drm files hold it."""
    X = m.coords
    x0 = np.asarray(x0, float); xl = np.asarray(xl, float)
    inside = np.all(np.abs(X - x0[None, :]) <= xl[None, :], axis=1)
    npe = 8 if m.ndim == 3 else 4
    cnt = inside[m.elem_conn[:, :npe]].sum(axis=1)
    elems = np.nonzero((cnt > 0) & (cnt < npe))[0].astype(np.int32)
    nodes = np.unique(m.elem_conn[elems, :npe]).astype(np.int32)
    ext = (~inside[nodes]).astype(np.uint8)
    d = DRMLoad(elems=elems, nodes=nodes, exterior=ext, planewave=dict(planewave), factor=factor)
    if tabulate_nt:
        pw = planewave
        dirv = np.asarray(pw["dir"], float); pol = np.asarray(pw["pol"], float); xr = np.asarray(pw["xref"], float)
        t = np.arange(tabulate_nt) * m.dt
        s = ((X[nodes] - xr[None, :]) @ dirv) / pw["c"]
        tau = t[None, :] - pw["t0"] - s[:, None]
        u, v, a = ricker_uva(tau, pw["f0"])
        nd = m.ndim
        fld = np.zeros((len(nodes), tabulate_nt, 3 * nd))
        for c in range(nd):
            fld[:, :, c] = pw["amp"] * pol[c] * u
            fld[:, :, nd + c] = pw["amp"] * pol[c] * v
            fld[:, :, 2 * nd + c] = pw["amp"] * pol[c] * a
        d.field = fld
    m.drm = d
    return m


# -------------------------------------------------------------------------------
# PML layer (Method/Builder.py:804-1008 setPMLattributes / setPMLDomain, :581-667 mergeDomain)
# -------------------------------------------------------------------------------
def make_pml_model(ne, npml, h=1.0, soil=None, pml_mat=None, pml_n=2.0, pml_R=1.0e-5, th=1.0,
                   dt=None, nt=0, load_dir=None, series=None, rec_nodes=None) -> Model:
    """Soil box (2-D: nx x ny quads, 3-D: nx x ny x nz hexes, free top) wrapped by a `npml`-cell PML
    layer on its sides and bottom, assembled the way the reference pre-processor does it:
      * the PML is its own mesh with 5 (2-D) / 9 (3-D) dofs per node; its interface nodes duplicate
        the soil boundary nodes and are tied to them by EQUAL constraints on the displacement dofs
        (mergeDomain, Builder.py:653-666); soil nodes / elements are numbered first;
      * per-element PML attributes x0 / npml from the element centroid (setPMLattributes);
      * displacement dofs of the outer PML boundary are restrained (B10 / J13 model scripts).
    Element attrs: 3-D [n, L, R, x0(3), npml(3)], 2-D [th, n, L, R, x0(2), npml(2)]."""
    nd = len(ne)
    is3 = nd == 3
    if soil is None:
        soil = (ELASTIC3DLINEAR if is3 else ELASTIC2DPLANESTRAIN, [1.3e7, 0.3, 2000.0])
    if pml_mat is None:
        pml_mat = (soil[0], list(soil[1]))
    p = int(npml)
    L = p * h
    n_cells = list(ne)
    top = nd - 1                                   # vertical axis: last one, free surface at its max
    # ---- soil lattice -------------------------------------------------------------
    if is3:
        nx, ny, nz = ne
        Xs = box_nodes(ne, [0, 0, 0], [nx * h, 0, 0], [0, ny * h, 0], [0, 0, nz * h])
        conn_s = box_hex8_conn(ne)
    else:
        nx, ny = ne
        Xs = area_nodes(ne, [0, 0], [nx * h, 0], [0, ny * h])
        conn_s = area_quad4_conn(ne)
    ns = Xs.shape[0]
    sdim = [c + 1 for c in n_cells]

    def soil_id(ijk):
        idx = ijk[0] + sdim[0] * ijk[1]
        if is3:
            idx = idx + sdim[0] * sdim[1] * ijk[2]
        return idx

    # ---- extended lattice: indices lo..hi per axis ----------------------------------
    lo = [-p] * nd
    hi = [c + p for c in n_cells]
    hi[top] = n_cells[top]
    ext_dim = [hi[a] - lo[a] + 1 for a in range(nd)]
    grids = np.meshgrid(*[np.arange(lo[a], hi[a] + 1) for a in reversed(range(nd))], indexing="ij")
    IJK = np.stack([g.ravel() for g in reversed(grids)], axis=1)          # x fastest
    strictly_in = np.ones(len(IJK), dtype=bool)
    on_or_in = np.ones(len(IJK), dtype=bool)
    for a in range(nd):
        if a == top:
            strictly_in &= IJK[:, a] > 0
            on_or_in &= IJK[:, a] >= 0
        else:
            strictly_in &= (IJK[:, a] > 0) & (IJK[:, a] < n_cells[a])
            on_or_in &= (IJK[:, a] >= 0) & (IJK[:, a] <= n_cells[a])
    is_pml_node = ~strictly_in
    pml_index = -np.ones(len(IJK), dtype=np.int64)
    pml_index[is_pml_node] = ns + np.arange(int(is_pml_node.sum()))
    Xp = IJK[is_pml_node].astype(np.float64) * h
    npn = Xp.shape[0]

    def ext_id(ijk):
        idx = (ijk[0] - lo[0]) + ext_dim[0] * (ijk[1] - lo[1])
        if is3:
            idx = idx + ext_dim[0] * ext_dim[1] * (ijk[2] - lo[2])
        return idx

    # ---- PML cells ----------------------------------------------------------------------
    cgr = np.meshgrid(*[np.arange(lo[a], hi[a]) for a in reversed(range(nd))], indexing="ij")
    C = np.stack([g.ravel() for g in reversed(cgr)], axis=1)
    cell_in = np.ones(len(C), dtype=bool)
    for a in range(nd):
        if a == top:
            cell_in &= C[:, a] >= 0
        else:
            cell_in &= (C[:, a] >= 0) & (C[:, a] < n_cells[a])
    Cp = C[~cell_in]
    offs = kHexPos if is3 else kQuadPos
    conn_p = np.zeros((len(Cp), 8), dtype=np.int32)
    for l, o in enumerate(offs):
        q = Cp + np.asarray(o)[None, :]
        conn_p[:, l] = pml_index[ext_id(q.T)]
    assert (conn_p[:, :len(offs)] >= ns).all()
    # attributes (setPMLattributes): centroid classification against the soil box
    cen = (Cp + 0.5) * h
    x0c = np.array([n_cells[a] * h / 2.0 for a in range(nd)]); x0c[top] = n_cells[top] * h
    xl = np.array([n_cells[a] * h / 2.0 for a in range(nd)]); xl[top] = n_cells[top] * h
    attr_p = np.zeros((len(Cp), 10))
    for e in range(len(Cp)):
        sgn = np.zeros(nd)
        for a in range(nd):
            if a == top:
                sgn[a] = -1.0 if cen[e, a] < x0c[a] - xl[a] else 0.0
            else:
                sgn[a] = -1.0 if cen[e, a] < x0c[a] - xl[a] else (1.0 if cen[e, a] > x0c[a] + xl[a] else 0.0)
        cnt = int(np.count_nonzero(sgn))
        npv = sgn / math.sqrt(cnt)
        x0e = x0c + sgn * xl
        if sgn[top] == 0.0:
            x0e[top] = x0c[top] - xl[top] / 2.0
        if is3:
            attr_p[e, :3] = [pml_n, L, pml_R]; attr_p[e, 3:6] = x0e; attr_p[e, 6:9] = npv
        else:
            attr_p[e, :4] = [th, pml_n, L, pml_R]; attr_p[e, 4:6] = x0e; attr_p[e, 6:8] = npv

    # ---- assemble the model ----------------------------------------------------------------
    m = Model(ndim=nd)
    m.coords = np.vstack([Xs, Xp])
    ndof_p = 9 if is3 else 5
    m.node_ndof = np.concatenate([np.full(ns, nd), np.full(npn, ndof_p)]).astype(np.int32)
    fd = [np.zeros(nd, dtype=np.int32) for _ in range(ns)] + [np.zeros(ndof_p, dtype=np.int32) for _ in range(npn)]
    pIJK = IJK[is_pml_node]
    iface = on_or_in[is_pml_node]
    outer = np.zeros(npn, dtype=bool)
    for a in range(nd):
        outer |= pIJK[:, a] == lo[a]
        if a != top:
            outer |= pIJK[:, a] == hi[a]
    tag = -1
    cons = []
    ptr_p = ns * nd + ndof_p * np.arange(npn)
    for q in range(npn):
        if outer[q]:
            fd[ns + q][:nd] = -1
        elif iface[q]:
            master = soil_id(pIJK[q])
            for k in range(nd):
                tag -= 1
                fd[ns + q][k] = tag
                cons.append((tag, int(ptr_p[q] + k), (int(master), k)))
    m.freedof = fd
    m.materials = [soil, pml_mat]
    conn = np.zeros((len(conn_s) + len(conn_p), 8), dtype=np.int32)
    conn[:len(conn_s), :conn_s.shape[1]] = conn_s
    conn[len(conn_s):] = conn_p
    m.elem_conn = conn
    m.elem_kind = np.concatenate([np.full(len(conn_s), LIN3DHEXA8 if is3 else LIN2DQUAD4),
                                  np.full(len(conn_p), PML3DHEXA8 if is3 else PML2DQUAD4)]).astype(np.int32)
    m.elem_mat = np.concatenate([np.zeros(len(conn_s)), np.ones(len(conn_p))]).astype(np.int32)
    m.elem_attr = np.zeros((len(conn), 10))
    if not is3:
        m.elem_attr[:len(conn_s), 0] = th
    m.elem_attr[len(conn_s):] = attr_p
    m.blocks = [(0, sdim[0], sdim[1], sdim[2] if is3 else 1)]
    if dt is None:
        E, nu, rho = soil[1][:3]
        lam = E * nu / ((1 + nu) * (1 - 2 * nu)); mu = E / (2 * (1 + nu))
        dt = 0.5 * h / math.sqrt((lam + 2 * mu) / rho)
    m.dt, m.nt = dt, nt
    m.number_dofs()
    # constraints in the JSON form: slave TOTAL dof <- master FREE dof (SeismoVLAB.py:130-138)
    m.constraints = [(t, s, [int(m.freedof_flat[m.node_ptr[mn] + k])], [1.0]) for t, s, (mn, k) in cons]
    ijk_load = [c // 2 for c in n_cells]; ijk_load[top] = n_cells[top]
    load_node = soil_id(ijk_load)
    if load_dir is None:
        load_dir = (0.0, 0.0, 1.0e4) if is3 else (0.0, 1.0e4)
    if series is None and nt > 0:
        f0 = 1.0 / (20.0 * dt)
        series = ricker(nt, dt, f0, 1.2 / f0)
    if series is not None:
        m.point_loads = [PointLoad(np.array([load_node], dtype=np.int32), np.asarray(load_dir, float),
                                   np.asarray(series, float))]
    m.rec_nodes = np.asarray(rec_nodes if rec_nodes is not None else [load_node], dtype=np.int32)
    m.n_soil_nodes, m.n_soil_elems = ns, len(conn_s)
    return m


kHexPos = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
kQuadPos = [(0, 0), (1, 0), (1, 1), (0, 1)]
