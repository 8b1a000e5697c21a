"""Named parity cases shared by the golden-vector generator (tests/golden/make_golden.py), the CPU
tests of the oracle and the GPU parity tests.  Every case is a deterministic svl_b200.model.Model;
the golden file of a case holds the NODE-recorder history the UNMODIFIED reference executable
(oracle/_ref/SeismoVLAB.exe) wrote for exactly this model (CentralDifference + Linear + Eigen)."""
import hashlib
import math

import numpy as np

from svl_b200 import model as M

SOIL = [1.3e7, 0.3, 2000.0]            # fixture J05
PMLMAT = [5.0e7, 0.25, 2000.0]         # fixtures J12 / B10
PML_DT = 0.25 / math.sqrt(1.3e7 * 0.7 / (1.3 * 0.4) / 2000.0)   # 0.25 h / Vp(SOIL), h = 1
J2 = [2.9e7, 2.0e7, 2000.0, 1.0e7, 0.5, 1.0e4]   # fixture F07 (beta 0.5 as in SURVEY App. B.5)


def kat444():
    """SURVEY.md App. B.5 CentralDifference known-answer case."""
    nt = 51
    series = np.array([math.sin(2 * math.pi * k / 20) for k in range(nt)])
    return M.make_box_model((4, 4, 4), 1.0, dt=0.004, nt=nt, load_node=112, load_dir=(2e3, -1e3, 1e4),
                            series=series, rec_nodes=[62, 112, 124])


def kat444_masses():
    """kat444 with nodal point masses (Assembler::AssembleNodalMass, Assembler.cpp:622-657; JSON "Masses"): unequal per-dof values,
    one of them zero (dropped by the |m| > mtol filter), on the loaded node and on an interior node."""
    m = kat444()
    m.masses = [(112, [500.0, 500.0, 250.0]), (62, [100.0, 0.0, 300.0])]
    return m


def hex8_distorted():
    nt = 41
    series = np.array([math.sin(2 * math.pi * k / 16) for k in range(nt)])
    return M.make_box_model((3, 4, 5), 0.8, dt=0.003, nt=nt, load_dir=(1e3, 2e3, -5e3), series=series,
                            rec_nodes=[30, 77, 119], jitter=0.15)


def hex8_layered_rayleigh():
    mats = [(M.ELASTIC3DLINEAR, SOIL), (M.ELASTIC3DLINEAR, [5.0e7, 0.25, 2200.0])]
    m = M.make_box_model((4, 3, 6), 1.0, nt=60, layers=mats, rec_nodes=[22, 70, 139])
    m.elem_am = np.where(m.elem_mat == 0, 0.8, 0.3)
    m.elem_ak = np.zeros(m.n_elem)
    return m


def quad4_area():
    return M.make_area_model((8, 6), 0.5, th=0.8, nt=70, rec_nodes=[12, 40, 58])


def quad4_distorted():
    return M.make_area_model((6, 5), 0.5, th=1.3, nt=50, rec_nodes=[10, 27, 41], jitter=0.2)


def j2_column():
    return M.make_box_model((2, 2, 6), 1.0, mat=(M.PLASTIC3DJ2, J2), nt=100, load_dir=(3.0e5, 0.0, 1.0e5),
                            rec_nodes=[40, 58, 62])


def drm_box():
    nt = 50
    ne = (6, 6, 5)
    m = M.make_box_model(ne, 1.0, nt=nt, series=None, fix=None)
    m.point_loads = []
    nx, ny, nz = ne
    vs = math.sqrt(SOIL[0] / (2 * (1 + SOIL[1])) / SOIL[2])
    pw = dict(dir=[0.0, 0.0, 1.0], pol=[1.0, 0.0, 0.0], xref=[0.0, 0.0, 0.0], c=vs, f0=1.0 / (25 * m.dt),
              t0=30 * m.dt, amp=1e-3)
    M.add_drm_box(m, x0=[nx / 2, ny / 2, nz], xl=[nx / 2 - 1.5, ny / 2 - 1.5, nz - 1.5], planewave=pw,
                  tabulate_nt=nt)
    m.rec_nodes = np.array([0, m.n_nodes // 2, m.n_nodes - 1, int(m.drm.nodes[3])], dtype=np.int32)
    return m


def drm_area():
    nt = 50
    ne = (8, 6)
    m = M.make_area_model(ne, 1.0, nt=nt, series=None, fix=None)
    m.point_loads = []
    nx, ny = ne
    vs = math.sqrt(SOIL[0] / (2 * (1 + SOIL[1])) / SOIL[2])
    pw = dict(dir=[0.0, 1.0], pol=[1.0, 0.0], xref=[0.0, 0.0], c=vs, f0=1.0 / (25 * m.dt), t0=30 * m.dt, amp=1e-3)
    M.add_drm_box(m, x0=[nx / 2, ny], xl=[nx / 2 - 1.5, ny - 1.5], planewave=pw, tabulate_nt=nt)
    m.rec_nodes = np.array([0, m.n_nodes // 2, m.n_nodes - 1, int(m.drm.nodes[2])], dtype=np.int32)
    return m


def pml2d():
    # dt = 0.25 h / Vp: the reference's CentralDifference + PML pair (G term dropped) grows
    # slowly at the usual 0.5 h / Vp; the parity window stays in the stable regime
    m = M.make_pml_model((6, 5), 3, 1.0, soil=(M.ELASTIC2DPLANESTRAIN, SOIL), nt=120, dt=PML_DT)
    ns = m.n_soil_nodes
    m.rec_nodes = np.array([0, 17, int(m.point_loads[0].nodes[0]), ns + 5, ns + 40], dtype=np.int32)
    return m


def pml3d():
    m = M.make_pml_model((3, 3, 3), 2, 1.0, soil=(M.ELASTIC3DLINEAR, SOIL), nt=80, dt=PML_DT)
    ns = m.n_soil_nodes
    m.rec_nodes = np.array([5, int(m.point_loads[0].nodes[0]), ns + 7, ns + 100], dtype=np.int32)
    return m


def _soil_speeds():
    lam = SOIL[0] * SOIL[1] / ((1 + SOIL[1]) * (1 - 2 * SOIL[1])); mu = SOIL[0] / (2 * (1 + SOIL[1]))
    return math.sqrt(mu / SOIL[2]), math.sqrt((lam + 2 * mu) / SOIL[2])


def lysmer_column():
    """Soil column on a Lysmer-Kuhlemeyer base: ZeroLength1D + Viscous1DLinear dashpots between every base node and a
    fixed twin (Builder.py:1086-1131), as in fixture J05 but under CentralDifference."""
    m = M.make_box_model((3, 3, 6), 1.0, nt=90, fix=None, load_dir=(4e3, -2e3, 1e4), rec_nodes=[0, 5, 53, 111])
    vs, vp = _soil_speeds()
    return M.add_base_dashpots(m, vs, vp, SOIL[2], 1.0)


def lysmer_area():
    m = M.make_area_model((6, 5), 0.5, th=0.8, nt=90, fix=None, load_dir=(3e3, 1e4), rec_nodes=[0, 3, 20, 41])
    vs, vp = _soil_speeds()
    return M.add_base_dashpots(m, vs, vp, SOIL[2], 0.5, th=0.8)


def j2ps_area():
    """lin2DQuad4 + PlasticPlaneStrainJ2 (PlasticPlaneStrainJ2.cpp:227-278), loaded into yield."""
    vp = math.sqrt((J2[0] + 4.0 * J2[1] / 3.0) / J2[2])
    return M.make_area_model((4, 8), 1.0, th=1.0, mat=(M.PLASTICPLANESTRAINJ2, J2), nt=110, dt=0.5 / vp,
                             load_dir=(3.0e5, 1.0e5), rec_nodes=[12, 27, 40, 44])


def c1_column20():
    """BASELINE.json configs[0] at its full size: 3-D linear elastic soil column, 20 x 20 x 20 lin3DHexa8 + Elastic3DLinear
    (fixture J05's soil), bottom fixed, lumped CentralDifference at dt = 0.5 h / Vp, Ricker point load at the centre of the free
    surface; 80 steps carry the P front to the base and back.  16 recorded nodes: a 4 x 4 grid over the free surface plus
    interior points on the load axis."""
    n = 20
    N1 = n + 1
    top = [i + N1 * j + N1 * N1 * n for j in (0, 7, 13, 20) for i in (0, 7, 13, 20)]
    axis = [10 + N1 * 10 + N1 * N1 * k for k in (5, 10, 15)]
    return M.make_box_model((n, n, n), 1.0, mat=(M.ELASTIC3DLINEAR, SOIL), nt=81, rec_nodes=top[:13] + axis)


# ---- reactions and support motion (a7) ---------------------------------------------------
def reaction_box():
    """Bottom-fixed layered box with mass-proportional Rayleigh damping, point masses (one on a FIXED node: its inertia force
    enters the reaction, Assembler.cpp:568-590) and a point load that also acts on a fixed node (Fext is subtracted from the
    reaction, CentralDifference.cpp:168).  Recorded: base corner / edge / interior nodes and two free nodes (zero rows)."""
    m = hex8_layered_rayleigh()
    m.masses = [(7, [300.0, 200.0, 100.0]), (70, [150.0, 150.0, 150.0])]
    pl = m.point_loads[0]
    m.point_loads.append(M.PointLoad(np.array([6, 27], dtype=np.int32), np.array([5e2, -3e2, 2e3]), 0.5 * pl.series, 1.5))
    m.rec_nodes = np.array([0, 2, 7, 12, 19, 27, 70], dtype=np.int32)
    return m


def reaction_area():
    m = M.make_area_model((8, 6), 0.5, th=0.8, nt=70, rec_nodes=[0, 4, 8, 12, 40])
    m.elem_am = np.full(m.n_elem, 0.6); m.elem_ak = np.zeros(m.n_elem)
    return m


def _pulse(nt, dt, amp, f0):
    t = np.arange(nt) * dt
    return amp * np.sin(2 * np.pi * f0 * t) * np.exp(-((t - 0.5 * nt * dt) / (0.25 * nt * dt)) ** 2)


def support_column():
    """Soil column shaken through its fixed base: TIMESERIES support motion on every base node in x (factor 1.25 from the
    combination), in z on the same nodes with another history, one CONSTANT entry (which the dynamic loop never applies:
    Node.cpp:236-243 with k >= 1) -- plus a point load at the top.  Reference semantics reproduced: the support increment
    reaches the elements one step late."""
    m = M.make_box_model((3, 3, 6), 1.0, nt=90, load_dir=(4e2, -2e2, 1e3), rec_nodes=[0, 5, 15, 53, 111])
    gx = _pulse(m.nt, m.dt, 2e-3, 1.0 / (30 * m.dt)); gz = _pulse(m.nt, m.dt, -1e-3, 1.0 / (22 * m.dt))
    for n in range(16):
        m.supports.append((n, 0, gx, 1.25))
        m.supports.append((n, 2, gz, 1.25))
    m.supports.append((3, 1, np.array([5e-4]), 1.25))
    return m


def support_area():
    m = M.make_area_model((6, 5), 0.5, th=0.8, nt=90, load_dir=(3e2, 1e3), rec_nodes=[0, 3, 6, 20, 41])
    m.elem_am = np.full(m.n_elem, 0.4); m.elem_ak = np.zeros(m.n_elem)
    g = _pulse(m.nt, m.dt, 1e-3, 1.0 / (25 * m.dt))
    for n in range(7):
        m.supports.append((n, 0, g, 1.0))
    for n in (0, 6):
        m.supports.append((n, 1, 0.5 * g[::-1].copy(), 1.0))
    return m


def reaction_lysmer():
    """lysmer_column with the fixed dashpot twins recorded: their reaction is the force the Lysmer base absorbs
    (ZeroLength1D::ComputeInternalDynamicForces, ZeroLength1D.cpp:257-277)."""
    m = lysmer_column()
    ns = 4 * 4 * 7
    m.rec_nodes = np.array([0, 5, ns, ns + 5, ns + 15], dtype=np.int32)
    return m


# ---- mid-size goldens in the shapes of BASELINE configs[1] / [2] / [4] (the toy meshes above exercise one tile of
# the lattice kernels; these are the largest the reference executable finishes in minutes with the shim's envelope LDL^T) -------
def mid_quad4_pml():
    """configs[1] shape: 200 x 100 lin2DQuad4 half-space + 5-cell PML2DQuad4 layer (left / right / bottom), point source."""
    m = M.make_pml_model((200, 100), 5, 1.0, soil=(M.ELASTIC2DPLANESTRAIN, SOIL), nt=160, dt=PML_DT)
    ns = m.n_soil_nodes
    N1 = 201
    m.rec_nodes = np.array([0, 100 + N1 * 50, 100 + N1 * 100, 30 + N1 * 80, 199 + N1 * 3, int(m.point_loads[0].nodes[0]), ns + 7, ns + 900], dtype=np.int32)
    return m


def mid_j2():
    """configs[4] shape: 10 x 10 x 40 lin3DHexa8 + Plastic3DJ2 column, base shear + vertical load through the free surface,
    loaded into yield (checked by the GPU test against the elastic twin)."""
    m = M.make_box_model((10, 10, 40), 1.0, mat=(M.PLASTIC3DJ2, J2), nt=140, load_dir=(3.0e5, 0.0, 1.0e5))
    N1 = 11
    top = np.arange(N1 * N1 * 40, N1 * N1 * 41, dtype=np.int32)
    s = m.point_loads[0].series
    m.point_loads = [M.PointLoad(top, np.array([6.0e3, 0.0, 2.0e3]), s.copy())]
    m.rec_nodes = np.array([5 + N1 * 5 + N1 * N1 * k for k in (40, 30, 20, 10, 3)] + [0 + N1 * 0 + N1 * N1 * 40], dtype=np.int32)
    return m


MID_CASES = {f.__name__: f for f in (mid_quad4_pml, mid_j2)}
# (a mid-size PML3DHexa8 golden is out of reach of the reference build in this image: the shim's envelope LDL^T of the ~46 000-dof
# coupled system of a 14 x 14 x 12 + 3-cell model is near-dense -- it had not finished its first factorisation after 45 minutes)

# cases whose goldens also hold `reaction` (and vel / accel where the support moves): tests/golden/make_golden.py
REACTION_CASES = ("reaction_box", "reaction_area", "support_column", "support_area", "reaction_lysmer")

CASES = {f.__name__: f for f in (c1_column20, kat444, kat444_masses, hex8_distorted, hex8_layered_rayleigh, quad4_area, quad4_distorted,
                                 j2_column, drm_box, drm_area, pml2d, pml3d, lysmer_column, lysmer_area, j2ps_area)}
# tolerance of |oracle - reference| and |device - oracle| per case (max_t|d| / max_t|ref| per dof)
REACTION_CASE_FUNCS = {f.__name__: f for f in (reaction_box, reaction_area, support_column, support_area, reaction_lysmer)}
TOL = {name: 1e-10 for name in list(CASES) + list(REACTION_CASE_FUNCS)}
TOL.update({"mid_quad4_pml": 1e-9, "mid_j2": 1e-8})
TOL["j2ps_area"] = 1e-8
TOL["j2_column"] = 1e-8        # plastic: looser bound (BASELINE.json north_star), stated in DESIGN.md
TOL["pml2d"] = 1e-9            # PML: Keff is not diagonal -> iterative block solve (rtol 1e-14), see DESIGN.md
TOL["pml3d"] = 1e-9


# NewmarkBeta + Linear cases: the same generators run at 4x the explicit time step, i.e. beyond the
# CentralDifference stability limit; goldens `newmark_<case>.npz` come from the unmodified reference executable with
# integrator NEWMARK (Driver.hpp:1811-1813).
NEWMARK_CASES = ("kat444", "quad4_area", "lysmer_column", "hex8_layered_rayleigh", "hex8_distorted", "drm_box")
TOL_NEWMARK = 1e-9             # the device solves Keff dU = Feff by conjugate gradients (rtol 1e-13), the reference by LDL^T


def newmark_case(name):
    m = CASES[name]()
    m.dt *= 4.0
    return m


# NewmarkBeta + NewtonRaphson on the plastic cases (fixtures F03 / F07): (cnvgtol, nstep, cnvgtest) per run, one
# setting for each convergence test of Algorithm.cpp:122-186; goldens `newton_<case>.npz` from the unmodified reference executable.
# The settings stop the iteration before a Gauss point's trial state lands ON the yield surface to the last bit: beyond that the
# reference's `TrialF <= 0` branch (and with it the tangent of the next step's first iteration) is decided by rounding, so
# two correct implementations of the same algorithm separate at ~cnvgtol.
NEWTON_CASES = {
    "j2_column": ((1e-6, 50, 5), (1e1, 30, 1), (1e-7, 30, 3), (1e-4, 30, 4)),
    "j2ps_area": ((1e-8, 20, 2), (1e-10, 30, 6), (1e-5, 30, 7), (1.0, 30, 8)),
}
TOL_NEWTON = 1e-12


def fixture_j05():
    """The reference's own validation fixture 03-Validations/01-Debugging/J05-DY_Lin_3DSoilColumn_Elastic_Hexa8 rebuilt with
    the repo's model layer: 1 x 1 x 100 lin3DHexa8 column (Elastic3DLinear 1.3e7 / 0.3 / 2000), z restrained everywhere,
    Rayleigh am = 0.1244195110, ak = 0.0007878958 on the solids, four ZeroLength1D + Viscous1DLinear (eta 2.5e4) dashpots
    per horizontal direction between base nodes 1-4 and fixed twins 405-408, point loads 25000 * ricker.in in x on the base
    nodes, lumped mass, Newmark + Linear, dt 0.01, nt 1000; recorded nodes 1 and 401.  Its golden data are the OpenSees
    histories shipped with the fixture (tests/golden/J05/opensees.npz, 6 significant digits)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "J05", "opensees.npz"))
    m = M.make_box_model((1, 1, 100), 1.0, nt=1000, dt=0.01, fix=None, series=np.zeros(1000), rec_nodes=[0, 400])
    fd = np.asarray(m.freedof).reshape(-1, 3).copy()
    fd[:, 2] = -1                                                   # addRestrain(dof=[3]) on every soil node
    n0 = m.n_nodes
    m.coords = np.vstack([m.coords, m.coords[:4]])                  # Lysmer twins 405..408
    m.node_ndof = np.concatenate([m.node_ndof, np.full(4, 3, dtype=np.int32)])
    m.freedof = np.concatenate([fd.reshape(-1), np.full(12, -1, dtype=np.int32)])
    m.materials = [(M.ELASTIC3DLINEAR, SOIL), (M.VISCOUS1DLINEAR, [2.5e4])]
    conn, attr = [], []
    for b in range(4):
        for d in (0, 1):                                            # 'dir': 1, 2 in the script = 0, 1 in the JSON (Attach.py:496-497)
            row = np.zeros(8, dtype=np.int32); row[0], row[1] = b, n0 + b
            a = np.zeros(10); a[0] = d
            conn.append(row); attr.append(a)
    ne0 = m.n_elem
    m.elem_conn = np.vstack([m.elem_conn, np.array(conn, dtype=np.int32)])
    m.elem_kind = np.concatenate([m.elem_kind, np.full(8, M.ZEROLENGTH1D, dtype=np.int32)])
    m.elem_mat = np.concatenate([m.elem_mat, np.ones(8, dtype=np.int32)])
    m.elem_attr = np.vstack([np.zeros((ne0, 10)), np.array(attr)])
    m.elem_am = np.concatenate([np.full(ne0, 0.1244195110), np.zeros(8)])
    m.elem_ak = np.concatenate([np.full(ne0, 0.0007878958), np.zeros(8)])
    series = np.asarray(g["ricker"], float)
    m.point_loads = [M.PointLoad(np.array([b], dtype=np.int32), np.array([25000.0, 0.0, 0.0]), series.copy()) for b in range(4)]
    return m.number_dofs()


# The reference's own validation fixtures that sit on this path, as INPUT FILES: tests/golden/fixtures/<name>/Partition/*.json
# is what the reference's pre-processor (01-Pre_Process, run on the fixture's script by tests/golden/make_fixture_inputs.py)
# wrote, the load files are the fixture's, opensees.npz holds the fixture's OpenSees golden histories.  `cols` maps an
# OpenSees column to a column of our NODE recorder (as each fixture's LaTeX/cmpResults.py pairs them).
REF_FIXTURES = {
    "F02": dict(json="Debugging_F02.1.0.json", cols=((1, 0), (2, 1)), tol=5e-6),             # 1 lin2DQuad4, lumped
    "F06": dict(json="Debugging_F06.1.0.json", cols=((1, 0), (3, 2)), tol=5e-6),             # quad4 column + dashpots + Rayleigh
    "J02": dict(json="Debugging_J02.1.0.json", cols=((3, 0), (1, 1), (2, 2)), tol=5e-6),     # 1 lin3DHexa8, CONSISTENT mass
}
# NewmarkBeta + NewtonRaphson + PlasticPlaneStrainJ2 fixtures: reference.npz (the unmodified reference executable on these files)
# pins the algorithm; opensees.npz is the fixture's shipped golden (a different Newton scheme: the reference itself is only this
# close to it).  tol_exe: max |d| / max |ref|;  tol_os: relative RMS per column as cmpResults.py pairs them (disp, vel, accel).
NEWTON_FIXTURES = {
    "F03": dict(json="Debugging_F03.1.0.json", cols=((1, 0),), tol_exe=1e-11, tol_os=(2e-3, 2e-2, 3e-2)),   # 1 quad, rho = 0
    "F07": dict(json="Debugging_F07.1.0.json", cols=((1, 0), (3, 2)), tol_exe=5e-5, tol_os=(1e-4, 2e-4, 5e-4)),
}
# Fixtures the reference validates by a plot only (no shipped numbers): reference.npz = the unmodified reference executable on
# exactly these input files.  Both carry EQUAL constraints (soil-PML ties) and the consistent mass matrix.
EXE_FIXTURES = {
    "F11": dict(json="Debugging_F11.1.0.json"),      # lin2DQuad4 column + PML2DQuad4, NEWMARK
    "J12": dict(json="Debugging_J12.1.0.json"),      # lin3DHexa8 rod + PML3DHexa8, EXTENDEDNEWMARK (history matrix G)
}


def fixture_dir(name):
    import os
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fixtures", name)


def fixture_model(name):
    import os
    spec = REF_FIXTURES.get(name) or EXE_FIXTURES.get(name) or NEWTON_FIXTURES[name]
    return M.read_reference_json(os.path.join(fixture_dir(name), "Partition", spec["json"]))


def fixture_errors(name, hist, key):
    """max over the paired columns of the relative RMS error against the fixture's OpenSees history `key`"""
    import os
    g = np.load(os.path.join(fixture_dir(name), "opensees.npz"))[key]
    errs = []
    for col, ours in (REF_FIXTURES.get(name) or NEWTON_FIXTURES[name])["cols"]:
        ref = g[:, col]
        errs.append(np.sqrt(np.mean((hist[:, ours] - ref) ** 2)) / np.sqrt(np.mean(ref ** 2)))
    return max(errs)


def fingerprint(m) -> str:
    """Hash of the model inputs, stored beside each golden history so that drift of a generator is
    detected instead of silently comparing different models."""
    h = hashlib.sha256()
    for a in (m.coords, m.elem_conn, m.elem_kind, m.elem_mat, m.freedof_flat, m.node_ndof):
        h.update(np.ascontiguousarray(a).tobytes())
    h.update(repr((m.dt, m.nt, [(k, list(map(float, p))) for k, p in m.materials])).encode())
    for pl in m.point_loads:
        h.update(np.ascontiguousarray(pl.series).tobytes())
    if m.masses:                                   # only models that carry nodal masses (older goldens keep their hash)
        h.update(repr([(int(n), list(map(float, v))) for n, v in m.masses]).encode())
    for n, d, series, fac in getattr(m, "supports", None) or []:
        h.update(repr((int(n), int(d), float(fac))).encode()); h.update(np.ascontiguousarray(series, dtype=float).tobytes())
    return h.hexdigest()[:16]


def rel_err(a, b):
    """max over recorded dofs of max_t|a-b| / max_t|b|.  Dofs whose whole history stays
    below 1e-3 of the global peak are normalised by that floor instead (zero by symmetry / ahead of the
    wave front: both sides only hold rounding noise there, and the reference's |f| <= ftol assembly filter,
    Assembler.cpp:262, decides which precursors exist at all)."""
    scale = np.maximum(np.abs(b).max(axis=0), 1e-3 * np.abs(b).max())
    scale[scale == 0] = 1.0
    return (np.abs(a - b).max(axis=0) / scale).max()
