"""Parity of the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Tolerance (BASELINE.json north_star): nodal displacement histories within 1e-10 of the
reference for the linear elastic cases, measured as max_t|u - u_ref| / max_t|u_ref| per
recorded dof; the plastic case uses the looser 1e-8 stated in DESIGN.md."""
import math

import numpy as np
import pytest

from svl_b200 import model as M

pytestmark = pytest.mark.gpu

TOL_LINEAR = 1e-10
TOL_PLASTIC = 1e-8


def rel_err(a, b):
    """max over recorded dofs of max_t|a-b| / max_t|b|.  Dofs whose whole
    history stays below 1e-3 of the global peak are normalised by that floor instead: they are
    zero by symmetry or ahead of the wave front, both codes only hold ~1e-19 rounding noise
    there (and the reference's |f| <= ftol = 1e-12 assembly filter, Assembler.cpp:262, which
    the device path does not apply, decides which precursors exist at all)."""
    scale = np.maximum(np.abs(b).max(axis=0), 1e-3 * np.abs(b).max())
    scale[scale == 0] = 1.0
    return (np.abs(a - b).max(axis=0) / scale).max()


def _device(m, **kw):
    from svl_b200.capi import DeviceModel
    return DeviceModel(m, **kw)


def kat_model(hint=True, jitter=0.0):
    nt = 51
    series = np.array([math.sin(2 * math.pi * k / 20) for k in range(nt)])
    m = M.make_box_model((4, 4, 4), 1.0, dt=0.004, nt=nt, load_node=112, load_dir=(2e3, -1e3, 1e4),
                         series=series, rec_nodes=[62, 112, 124], jitter=jitter)
    if not hint:
        m.blocks = []
    return m


def test_kat_block_stencil(oracle):
    m = kat_model()
    ref, _ = oracle.run(m)
    d = _device(m)
    out = d.run()[0]
    c = d.counters()
    assert c["n_block_nodes"] == 125 and c["n_generic_elements"] == 0
    assert rel_err(out, ref) < TOL_LINEAR
    # SURVEY.md App. B.5 known answer produced by the reference executable
    assert abs(out[0, 3] - 9.888543819998318e-06) < 1e-19
    assert np.allclose(out[24, 3:6], [5.382280138183359e-04, -2.691140069091679e-04, 2.370556193531373e-03],
                       rtol=1e-11, atol=0)


def test_kat_generic_path(oracle):
    m = kat_model(hint=False)
    ref, _ = oracle.run(m)
    d = _device(m)
    out = d.run()[0]
    c = d.counters()
    assert c["n_block_nodes"] == 0 and c["n_generic_elements"] == 64
    assert rel_err(out, ref) < TOL_LINEAR


def test_distorted_mesh_generic(oracle):
    m = kat_model(jitter=0.15)
    ref, _ = oracle.run(m)
    out = _device(m).run()[0]
    assert rel_err(out, ref) < TOL_LINEAR


@pytest.mark.parametrize("ne", [(7, 5, 3), (33, 9, 6), (40, 35, 20)])
def test_box_block_vs_oracle(oracle, ne):
    nt = 40
    m = M.make_box_model(ne, 0.5, nt=nt, rec_nodes=None)
    nn = m.n_nodes
    m.rec_nodes = np.array(sorted({0, nn - 1, nn // 2, nn // 3, m.point_loads[0].nodes[0]}), dtype=np.int32)
    ref, Uref = oracle.run(m, nthreads=8)
    d = _device(m)
    out = d.run()[0]
    assert d.counters()["n_block_nodes"] == nn
    assert rel_err(out, ref) < TOL_LINEAR
    U = d.get_state(0)
    assert np.abs(U - Uref).max() / np.abs(Uref).max() < TOL_LINEAR


NS = {"SVLGPU_NO_SEP": "1"}            # the separable kernel is the default for the interior class: switch it off to reach the others


@pytest.mark.parametrize("env", [{}, {"SVLGPU_STENCIL_KZ": "4"}, {"SVLGPU_STENCIL_KZ": "1"}, {"SVLGPU_STENCIL_KZ": "2"}, {"SVLGPU_STENCIL_KZ": "5"},
                                 NS, {**NS, "SVLGPU_STENCIL_V": "3"}, {**NS, "SVLGPU_NO_SYM": "1"}, {**NS, "SVLGPU_STENCIL_R": "6"},
                                 {**NS, "SVLGPU_STENCIL_R": "6", "SVLGPU_STENCIL_KZ": "5"}, {**NS, "SVLGPU_STENCIL_KZ": "4"},
                                 {**NS, "SVLGPU_STENCIL_NOBAR": "1"}, {**NS, "SVLGPU_STENCIL_NOBAR": "1", "SVLGPU_STENCIL_KZ": "3"},
                                 {**NS, "SVLGPU_STENCIL_KZ": "1"}, {**NS, "SVLGPU_STENCIL_KZ": "2"}])
@pytest.mark.parametrize("ne", [(40, 35, 20), (67, 30, 13)])
def test_dominant_class_kernel_variants(oracle, monkeypatch, env, ne):
    """Every variant of the dominant-class block-stencil kernel (separable v5 -- the default --, TMA v3, barrier-free v4
    with / without the symmetric-coefficient table, 4 or 6 rows per thread, short z-chunks) against the oracle."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    nt = 30
    m = M.make_box_model(ne, 0.5, nt=nt, rec_nodes=None)
    nn = m.n_nodes
    m.rec_nodes = np.array(sorted({0, nn - 1, nn // 2, nn // 3, m.point_loads[0].nodes[0]}), dtype=np.int32)
    ref, Uref = oracle.run(m, nthreads=8)
    d = _device(m)
    out = d.run()[0]
    assert d.counters()["n_block_nodes"] == nn
    assert rel_err(out, ref) < TOL_LINEAR
    U = d.get_state(0)
    assert np.abs(U - Uref).max() / np.abs(Uref).max() < TOL_LINEAR


@pytest.mark.parametrize("ne,hs", [((37, 35, 21), (0.5, 0.8, 1.25)), ((66, 19, 9), (1.0, 1.0, 0.4))])
def test_separable_kernel_on_rectangular_cells(oracle, ne, hs):
    """k_stencil3_sep fits its tensor-product coefficients to the class table, so rectangular (non-cubic) cells and odd
    lattice dimensions (row-alignment shifts of the TMA copies) go through the same kernel."""
    nt = 30
    m = M.make_box_model(ne, min(hs), nt=nt, rec_nodes=None)          # dt from the smallest edge
    m.coords = m.coords * (np.array(hs) / min(hs))
    nn = m.n_nodes
    m.rec_nodes = np.array(sorted({0, nn - 1, nn // 2, nn // 3, m.point_loads[0].nodes[0]}), dtype=np.int32)
    rng = np.random.default_rng(7)
    V0 = rng.uniform(-1.0, 1.0, m.n_total)              # every dof moves from the first step on (U0 = 0: SURVEY App. C q2)
    V0[np.asarray(m.totaldof)[np.asarray(m.freedof_flat) < 0]] = 0.0
    ref, Uref = oracle.run(m, nthreads=8, V0=V0)
    d = _device(m, V0=V0)
    out = d.run()[0]
    assert d.counters()["n_block_nodes"] == nn
    assert rel_err(out, ref) < TOL_LINEAR
    U = d.get_state(0)
    assert np.abs(U - Uref).max() / np.abs(Uref).max() < TOL_LINEAR
    F = d.internal_force()                              # force-only mode of the same kernel
    Fref = oracle.internal_force(m, U)
    assert np.abs(F - Fref).max() / np.abs(Fref).max() < 1e-11


@pytest.mark.parametrize("kind", ["hex8_layers", "hex8", "quad4"])
def test_unstructured_numbering_uses_neighbour_list_node_classes(oracle, kind):
    """Random node / element numbering (no lattice block): nodes whose assembled row of K repeats are advanced from
    pre-summed row blocks + explicit neighbour lists (k_nbr_nodes), the rest by the Gauss-point kernels; both against
    the oracle, and against each other with the option switched off."""
    mats = [(M.ELASTIC3DLINEAR, [1.3e7, 0.3, 2000.0]), (M.ELASTIC3DLINEAR, [5.0e7, 0.25, 2200.0])]
    if kind == "quad4":
        m = M.make_area_model((40, 31), 0.5, th=0.8, nt=60)
    else:
        m = M.make_box_model((13, 11, 10), 1.0, nt=50, layers=mats if kind == "hex8_layers" else None)
        if kind == "hex8_layers":
            m.dt *= 0.5
    m.rec_nodes = np.array(sorted({0, m.n_nodes - 1, m.n_nodes // 2, m.n_nodes // 3, int(m.point_loads[0].nodes[0])}), dtype=np.int32)
    s = M.shuffle_numbering(m, 11)
    V0 = np.random.default_rng(5).uniform(-1.0, 1.0, s.n_total)
    V0[np.asarray(s.totaldof)[np.asarray(s.freedof_flat) < 0]] = 0.0
    ref, Uref = oracle.run(s, nthreads=8, V0=V0)
    d = _device(s, V0=V0)
    out = d.run()[0]
    c = d.counters()
    assert c["n_block_nodes"] == 0 and c["n_nbr_nodes"] > s.n_nodes // 2 and c["n_generic_elements"] < s.n_elem
    assert rel_err(out, ref) < TOL_LINEAR
    U = d.get_state(0)
    assert np.abs(U - Uref).max() / np.abs(Uref).max() < TOL_LINEAR
    F = d.internal_force()
    Fref = oracle.internal_force(s, U)
    assert np.abs(F - Fref).max() / np.abs(Fref).max() < 1e-11
    d2 = _device(s, V0=V0, options={"nbr_classes": 0.0})
    out2 = d2.run()[0]
    assert d2.counters()["n_nbr_nodes"] == 0 and d2.counters()["n_generic_elements"] == s.n_elem
    assert rel_err(out2, out) < TOL_LINEAR


def test_vel_accel_recorders(oracle):
    m = kat_model()
    d = _device(m, fields=(0, 1, 2))
    out = d.run()
    for f in (0, 1, 2):
        ref, _ = oracle.run(m, field=f)
        assert rel_err(out[f], ref) < 1e-9, f


def test_layered_materials(oracle):
    mats = [(M.ELASTIC3DLINEAR, [1.3e7, 0.3, 2000.0]), (M.ELASTIC3DLINEAR, [5.0e7, 0.25, 2200.0])]
    m = M.make_box_model((6, 5, 8), 1.0, nt=60, layers=mats)
    ref, _ = oracle.run(m)
    out = _device(m).run()[0]
    assert rel_err(out, ref) < TOL_LINEAR


def test_internal_force_and_mass(oracle):
    m = kat_model()
    rng = np.random.default_rng(42)
    U0 = rng.uniform(-1e-3, 1e-3, m.n_total)
    d = _device(m, U0=U0)
    F = d.internal_force()
    Fref = oracle.internal_force(m, U0)
    assert np.abs(F - Fref).max() / np.abs(Fref).max() < 1e-12
    assert np.abs(d.mass_diagonal() - oracle.mass_diagonal(m)).max() < 1e-9
    # generic path gives the same vector
    m2 = kat_model(hint=False)
    F2 = _device(m2, U0=U0).internal_force()
    assert np.abs(F2 - Fref).max() / np.abs(Fref).max() < 1e-12


def test_internal_force_leaves_the_state_buffers_alone(oracle):
    """svlgpu_internal_force works in a scratch vector of its own: VEL / ACCEL read after it (Integrator::GetVelocities after
    Assembler::ComputeInternalForceVector in the host facade) still equal the recorder rows of the last step."""
    m = kat_model()
    d = _device(m, fields=(0, 1, 2))
    d.step(1, 30, True)
    rows = [d.read_recorder(f)[-1] for f in (0, 1, 2)]
    d.internal_force()
    dofs = np.concatenate([m.totaldof[m.node_ptr[n]:m.node_ptr[n + 1]] for n in m.rec_nodes]).astype(np.int32)
    for f in (0, 1, 2):
        assert np.array_equal(d.get_state(f)[dofs], rows[f]), f
    d.step(30, m.nt, True)                             # and the run continues as if nothing had happened
    ref, _ = oracle.run(m)
    assert rel_err(d.read_recorder(0), ref) < TOL_LINEAR


def test_diverged_run_stops_with_an_error():
    """NaN / Inf in U must surface as stop (the reference does not check).  A time step ten times above the
    stability limit overflows within a few hundred steps; the node that is recorded sits far from where it starts."""
    from svl_b200.capi import SvlError
    m = M.make_box_model((6, 6, 6), 1.0, dt=0.05, nt=1500, rec_nodes=[0])
    d = _device(m)
    with pytest.raises(SvlError, match="NaN / Inf"):
        for k in range(1, 1500, 100):
            d.step(k, k + 100, True)


def test_j2_plastic_column(oracle):
    mat = (M.PLASTIC3DJ2, [2.9e7, 2.0e7, 2000.0, 1.0e7, 0.5, 1.0e4])
    m = M.make_box_model((3, 3, 8), 1.0, mat=mat, nt=120, load_dir=(3.0e5, 0.0, 1.0e5))
    ref, _ = oracle.run(m)
    d = _device(m)
    out = d.run()[0]
    assert d.counters()["n_generic_elements"] == m.n_elem
    assert rel_err(out, ref) < TOL_PLASTIC
    # the load must actually drive Gauss points past yield for this to test the return map
    K, G = 2.9e7, 2.0e7
    lin = M.make_box_model((3, 3, 8), 1.0, nt=120, load_dir=(3.0e5, 0.0, 1.0e5), dt=m.dt,
                           mat=(M.ELASTIC3DLINEAR, [9 * K * G / (3 * K + G), (3 * K - 2 * G) / (2 * (3 * K + G)), 2000.0]))
    assert rel_err(out, oracle.run(lin)[0]) > 1e-3


def test_quad4_block_and_generic(oracle):
    m = M.make_area_model((12, 9), 0.5, th=0.8, nt=80)
    ref, _ = oracle.run(m)
    d = _device(m)
    out = d.run()[0]
    assert d.counters()["n_block_nodes"] == m.n_nodes
    assert rel_err(out, ref) < TOL_LINEAR
    m.blocks = []
    out2 = _device(m).run()[0]
    assert rel_err(out2, ref) < TOL_LINEAR
    mj = M.make_area_model((12, 9), 0.5, th=0.8, nt=80, jitter=0.2)
    assert rel_err(_device(mj).run()[0], oracle.run(mj)[0]) < TOL_LINEAR


def test_step_host_roundtrip(oracle):
    m = kat_model()
    ref, _ = oracle.run(m)
    d = _device(m)
    row = np.zeros(9)
    for k in range(1, m.nt):
        d.step_host(k, [m.point_loads[0].series[k]], rec=0, row=row)
        assert np.abs(row - ref[k - 1]).max() <= TOL_LINEAR * np.abs(ref).max()


def drm_model(ne=(8, 8, 6), nt=60, tabulate=True):
    m = M.make_box_model(ne, 1.0, nt=nt, series=None, fix=None)
    m.point_loads = []
    nx, ny, nz = ne
    vs = math.sqrt(1.3e7 / (2 * 1.3) / 2000.0)
    pw = dict(dir=[0.0, 0.0, 1.0], pol=[1.0, 0.0, 0.0], xref=[0.0, 0.0, 0.0], c=vs, f0=1.0 / (25 * m.dt),
              t0=30 * m.dt, amp=1e-3)
    # box whose faces cut the 2nd element layer from the sides / bottom; open at the top
    M.add_drm_box(m, x0=[nx / 2, ny / 2, nz], xl=[nx / 2 - 1.5, ny / 2 - 1.5, nz - 1.5], planewave=pw,
                  tabulate_nt=nt if tabulate else 0)
    m.rec_nodes = np.array([0, m.n_nodes // 2, m.n_nodes - 1, int(m.drm.nodes[3])], dtype=np.int32)
    return m


def test_drm_tabulated_field(oracle):
    m = drm_model()
    assert len(m.drm.elems) > 0 and m.drm.exterior.sum() > 0
    ref, _ = oracle.run(m)
    assert np.abs(ref).max() > 1e-5
    out = _device(m).run()[0]
    assert rel_err(out, ref) < TOL_LINEAR
    m.blocks = []
    assert rel_err(_device(m).run()[0], ref) < TOL_LINEAR


def test_drm_analytic_planewave_matches_tabulated(oracle):
    m = drm_model()
    ref, _ = oracle.run(m)
    m.drm.field = None                       # device evaluates the Ricker plane wave itself
    out = _device(m).run()[0]
    assert rel_err(out, ref) < 1e-9


# ---- device vs the golden histories written by the reference executable (tests/golden/) ----------
import os

import cases

DEVICE_CASES = list(cases.CASES)


@pytest.mark.parametrize("name", DEVICE_CASES)
def test_device_matches_reference_golden(oracle, name):
    m = cases.CASES[name]()
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"{name}.npz"))
    assert str(g["fingerprint"]) == cases.fingerprint(m)
    d = _device(m)
    out = d.run()[0]
    assert d.counters()["total_launches"] > 0
    assert cases.rel_err(out, g["disp"]) < cases.TOL[name]
    ref, _ = oracle.run(m)
    assert cases.rel_err(out, ref) < cases.TOL[name]
    c = d.counters()
    if name.startswith("pml"):
        assert c["n_pml_elements"] > 0 and c["pml_solves"] == m.nt - 1
        print(f"{name}: {c['n_pml_unknowns']} block unknowns, {c['pml_iterations'] / c['pml_solves']:.1f} BiCGStab iterations/step, "
              f"err vs reference golden {cases.rel_err(out, g['disp']):.2e}")


@pytest.mark.parametrize("name", ["hex8_distorted", "quad4_distorted"])
def test_gauss_point_strain_stress_at_the_current_state(oracle, name):
    """svlgpu_get_gauss (Element::GetStrain / GetStress at Gauss points, lin3DHexa8.cpp:150-200): after an internal-force pass on
    the current state the kept Gauss-point strains equal the oracle's B u of that state and the stresses C eps."""
    import ctypes as C
    dp = C.POINTER(C.c_double)
    m = cases.CASES[name]()
    d = _device(m, options={"keep_gauss": 1.0})
    d.step(1, m.nt // 2, True)
    d.internal_force()                                   # Gauss-point pass on the current state, nothing committed
    U = d.get_state(0)
    elems = np.arange(m.n_elem, dtype=np.int32)
    eps, sig = d.gauss(0, elems), d.gauss(1, elems)
    is3 = m.ndim == 3
    npe, ngp, ncomp = (8, 8, 6) if is3 else (4, 4, 3)
    Cm = np.zeros(ncomp * ncomp)
    E, nu = m.materials[0][1][:2]
    (oracle.lib.svlo_elastic3d_C if is3 else oracle.lib.svlo_planestrain_C)(C.c_double(E), C.c_double(nu), Cm.ctypes.data_as(dp))
    Cm = Cm.reshape(ncomp, ncomp)
    ref = np.zeros((m.n_elem, ngp, ncomp))
    for e in range(m.n_elem):
        conn = m.elem_conn[e, :npe]
        X = np.ascontiguousarray(m.coords[conn].ravel())
        Ue = np.ascontiguousarray(np.concatenate([U[m.node_ptr[n]:m.node_ptr[n] + m.ndim] for n in conn]))
        (oracle.lib.svlo_hex8_strain if is3 else oracle.lib.svlo_quad4_strain)(X.ctypes.data_as(dp), Ue.ctypes.data_as(dp),
                                                                              ref[e].ctypes.data_as(dp))
    assert np.abs(ref).max() > 0
    assert np.abs(eps - ref).max() <= 1e-12 * np.abs(ref).max()
    assert np.abs(sig - ref @ Cm.T).max() <= 1e-12 * np.abs(ref @ Cm.T).max()


@pytest.mark.parametrize("name", list(cases.MID_CASES))
def test_device_matches_reference_golden_on_mid_size_config_shapes(name):
    """BASELINE configs[1] / [4] shapes at the largest size the reference executable finishes in minutes (200 x 100 quad4 +
    PML2D, 10 x 10 x 40 J2): many tiles / chunks / classes instead of the toy meshes' one."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"{name}.npz")
    if not os.path.exists(path):
        pytest.skip(f"{name}.npz has not been generated (tests/golden/make_golden.py mid)")
    m = cases.MID_CASES[name]()
    g = np.load(path)
    assert str(g["fingerprint"]) == cases.fingerprint(m)
    d = _device(m)
    out = d.run()[0]
    c = d.counters()
    assert cases.rel_err(out, g["disp"]) < cases.TOL[name]
    if name == "mid_j2":
        assert c["n_generic_elements"] == m.n_elem
        K, G = cases.J2[0], cases.J2[1]
        lin = cases.mid_j2()
        lin.materials = [(M.ELASTIC3DLINEAR, [9 * K * G / (3 * K + G), (3 * K - 2 * G) / (2 * (3 * K + G)), 2000.0])]
        assert cases.rel_err(_device(lin).run()[0], g["disp"]) > 1e-3          # the load really drives Gauss points past yield
    else:
        assert c["n_pml_elements"] > 0 and c["n_block_nodes"] > 0
        print(f"{name}: {c['n_pml_unknowns']} block unknowns, {c['pml_iterations'] / c['pml_solves']:.1f} BiCGStab iterations/step, "
              f"err vs reference golden {cases.rel_err(out, g['disp']):.2e}")


@pytest.mark.parametrize("name", list(cases.REACTION_CASES))
def test_reactions_and_support_motion_match_reference_golden(oracle, name):
    """REACTION recorders (restrained-row reaction pass at recorded steps) and support motion on the device against the
    recorder files of the unmodified reference executable and against the oracle: disp, vel, accel, reaction."""
    m = cases.REACTION_CASE_FUNCS[name]()
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"{name}.npz"))
    assert str(g["fingerprint"]) == cases.fingerprint(m)
    for hint in (True, False):                           # block-stencil kernels, then the Gauss-point kernels
        if not hint:
            m.blocks = []
        d = _device(m, fields=(0, 1, 2, 3))
        out = d.run()
        for f, key in ((0, "disp"), (1, "vel"), (2, "accel"), (3, "reaction")):
            assert cases.rel_err(out[f], g[key]) < cases.TOL[name], (key, hint, "vs reference")
            ref, _ = oracle.run(m, field=f)
            assert cases.rel_err(out[f], ref) < cases.TOL[name], (key, hint, "vs oracle")
        # the getters report the true support displacement as well (the state buffers hold what the elements see)
        dofs = np.asarray(m.rec_dofs(), np.int32)
        for f in (0, 1, 2):
            assert np.array_equal(d.get_state(f, dofs), out[f][-1]), f
        d.close()


def test_host_driver_writes_reaction_recorders_and_moves_supports(tmp_path):
    """The C++ host driver reads Supports / SUPPORTMOTION / resp = reaction from the reference's JSON and writes the
    reference's recorder files; compared with the files of the reference executable (goldens)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "svl_b200", "SeismoVLAB_gpu.exe")
    for name in ("support_column", "reaction_lysmer"):
        m = cases.REACTION_CASE_FUNCS[name]()
        g = np.load(os.path.join(root, "tests", "golden", f"{name}.npz"))
        work = os.path.join(str(tmp_path), name)
        part = M.write_reference_json(m, work, "Case", "Run", resp=("disp", "accel", "reaction"), ndps=17)
        r = subprocess.run([exe, "-dir", part, "-file", "Case.1.$.json"], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        for key in ("disp", "accel", "reaction"):
            out = M.read_node_recorder(os.path.join(work, "Solution", "Run", f"{key}.0.out"))
            assert cases.rel_err(out, g[key]) < cases.TOL[name], (name, key)


def test_pml_internal_force(oracle):
    for name in ("pml2d", "pml3d"):
        m = cases.CASES[name]()
        rng = np.random.default_rng(3)
        U0 = rng.uniform(-1e-3, 1e-3, m.n_total)
        F = _device(m, U0=U0).internal_force()
        Fref = oracle.internal_force(m, U0)
        assert np.abs(F - Fref).max() / np.abs(Fref).max() < 1e-12


def test_multigpu_interface_exchange():
    """2-rank (or more) run of tests/multigpu_check.py when the box has several GPUs."""
    import subprocess
    import sys

    import torch
    n = min(torch.cuda.device_count(), 8)
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 2 if n < 4 else (4 if n < 8 else 8)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29533",
                        os.path.join(root, "tests", "multigpu_check.py")], capture_output=True, text=True, timeout=900)
    print(r.stdout[-4000:], r.stderr[-2000:])
    assert r.returncode == 0


@pytest.mark.parametrize("name", ["kat444", "pml2d", "hex8_layered_rayleigh"])
def test_host_driver_runs_from_binary_partition_tables(tmp_path, name):
    """the same run from a partition file whose Nodes / Elements / Constraints / Dampings tables sit in binary
    sidecars (model.pack_partition_tables) -- same recorder file as from the plain JSON, same golden."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "svl_b200", "SeismoVLAB_gpu.exe")
    m = cases.CASES[name]()
    g = np.load(os.path.join(root, "tests", "golden", f"{name}.npz"))
    part = M.write_reference_json(m, str(tmp_path), "Case", "Run", resp=("disp",), ndps=17)
    M.pack_partition_tables(os.path.join(part, "Case.1.0.json"))
    out_file = os.path.join(str(tmp_path), "Solution", "Run", "disp.0.out")
    outs = []
    for pattern in ("Case.1.$.json", "Case.1.$.bin.json"):
        r = subprocess.run([exe, "-dir", part, "-file", pattern], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        outs.append((open(out_file).readline(), M.read_node_recorder(out_file)))
        os.remove(out_file)
    assert outs[0][0] == outs[1][0]                                   # identical header line
    assert cases.rel_err(outs[1][1], outs[0][1]) < 1e-13              # the same object graph reaches the device
    assert cases.rel_err(outs[1][1], g["disp"]) < cases.TOL[name]


# ---- the reference's command line / file formats on top of the C ABI (svl_b200/host) ----------------------
@pytest.mark.parametrize("name", ["kat444", "quad4_area", "drm_box", "j2_column", "hex8_layered_rayleigh", "pml2d", "pml3d",
                                  "lysmer_column", "lysmer_area", "j2ps_area"])
def test_host_driver_reads_reference_files_and_writes_reference_recorders(tmp_path, name):
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "svl_b200", "SeismoVLAB_gpu.exe")
    assert os.path.exists(exe), "run __graft_entry__.build() first"
    m = cases.CASES[name]()
    g = np.load(os.path.join(root, "tests", "golden", f"{name}.npz"))
    part = M.write_reference_json(m, str(tmp_path), "Case", "Run", resp=("disp",), ndps=17)
    r = subprocess.run([exe, "-dir", part, "-file", "Case.1.$.json"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    out_file = os.path.join(str(tmp_path), "Solution", "Run", "disp.0.out")
    head = open(out_file).readline().split()
    assert int(head[0]) == len(m.rec_nodes) and int(head[2]) == m.n_total and int(head[3]) == m.nt   # Recorder.cpp:92
    out = M.read_node_recorder(out_file)
    assert out.shape == g["disp"].shape
    assert cases.rel_err(out, g["disp"]) < cases.TOL[name]
    if name == "kat444":
        assert "lattice nodes 125" in r.stdout          # the planner recognised the makeDomainVolume lattice by itself
    # when the reference executable travelled to this box, run it on the very same files and compare the text outputs
    ref_exe = os.path.join(root, "oracle", "_ref", "SeismoVLAB.exe")
    if os.path.exists(ref_exe):
        os.rename(out_file, out_file + ".gpu")
        subprocess.run([ref_exe, "-dir", part, "-file", "Case.1.$.json"], stdout=subprocess.DEVNULL, check=True, timeout=600)
        ref = M.read_node_recorder(out_file)
        assert cases.rel_err(out, ref) < cases.TOL[name]
        assert open(out_file).readline() == open(out_file + ".gpu").readline()        # identical header line


@pytest.mark.parametrize("name", ["kat444", "c1_column20", "kat444_masses", "hex8_layered_rayleigh", "quad4_area", "j2_column", "lysmer_column"])
def test_reference_executable_with_the_gpu_integrator_linked_in(tmp_path, name):
    """The drop-in at link time: oracle/_ref/SeismoVLAB_refgpu.exe is the reference's own objects (Driver, DynamicAnalysis,
    Recorder, Mesh, elements ...) with CentralDifference.o replaced by integration/GPUCentralDifference.cpp + libsvlgpu.so.
    It reads the same JSON, its own Recorder writes the files -- and they must equal the golden files the unmodified
    executable wrote (and that executable's output on this very input, when it travelled to this box)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "oracle", "_ref", "SeismoVLAB_refgpu.exe")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/SeismoVLAB_refgpu.exe was not built (needs /root/reference at build time)")
    m = cases.CASES[name]()
    g = np.load(os.path.join(root, "tests", "golden", f"{name}.npz"))
    part = M.write_reference_json(m, str(tmp_path), "Case", "Run", resp=("disp", "vel", "accel"), ndps=17)
    r = subprocess.run([exe, "-dir", part, "-file", "Case.1.$.json"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "GPU CentralDifference:" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    out = {k: M.read_node_recorder(os.path.join(str(tmp_path), "Solution", "Run", f"{k}.0.out")) for k in ("disp", "vel", "accel")}
    assert out["disp"].shape == g["disp"].shape
    assert cases.rel_err(out["disp"], g["disp"]) < cases.TOL[name]
    ref_exe = os.path.join(root, "oracle", "_ref", "SeismoVLAB.exe")
    if os.path.exists(ref_exe):
        for k in out:
            os.rename(os.path.join(str(tmp_path), "Solution", "Run", f"{k}.0.out"), os.path.join(str(tmp_path), f"{k}.gpu"))
        subprocess.run([ref_exe, "-dir", part, "-file", "Case.1.$.json"], stdout=subprocess.DEVNULL, check=True, timeout=900)
        for k in out:
            ref = M.read_node_recorder(os.path.join(str(tmp_path), "Solution", "Run", f"{k}.0.out"))
            assert cases.rel_err(out[k], ref) < max(cases.TOL[name], 1e-9 if k != "disp" else 0.0), k


@pytest.mark.parametrize("name", ["kat444", "drm_box", "quad4_area", "j2_column"])
def test_cuda_graph_replay_matches(oracle, name):
    """steps replayed from a CUDA graph (device-resident step index / recorder row) give the same history"""
    m = cases.CASES[name]()
    ref, _ = oracle.run(m)
    d = _device(m, options={"cuda_graph": 1})
    out = d.run()[0]
    assert cases.rel_err(out, ref) < cases.TOL[name]
    plain = _device(m).run()[0]
    assert np.array_equal(out, plain)


# ---- NewmarkBeta + Linear on the device ----------------------------------------------------------
@pytest.mark.parametrize("name", list(cases.NEWMARK_CASES))
def test_newmark_device_matches_reference_golden(oracle, name):
    """Matrix-free conjugate-gradient Newmark step against the histories the unmodified reference executable wrote with
    integrator NEWMARK (direct LDL^T solve) and against the oracle; displacement, velocity and acceleration."""
    m = cases.newmark_case(name)
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"newmark_{name}.npz"))
    assert str(g["fingerprint"]) == cases.fingerprint(m)
    d = _device(m, fields=(0, 1, 2), options={"integrator": 1.0})
    out = d.run()
    for f, key in ((0, "disp"), (1, "vel"), (2, "accel")):
        ref, _ = oracle.run(m, field=f, integrator="NEWMARK")
        assert cases.rel_err(out[f], g[key]) < cases.TOL_NEWMARK, (key, "vs reference")
        assert cases.rel_err(out[f], ref) < cases.TOL_NEWMARK, (key, "vs oracle")
    U = d.get_state(0)
    _, Uref = oracle.run(m, integrator="NEWMARK")
    assert np.abs(U - Uref).max() / np.abs(Uref).max() < cases.TOL_NEWMARK
    assert d.counters()["total_launches"] > 0


def test_newmark_generic_path_and_refusals(oracle):
    m = cases.newmark_case("kat444")
    m.blocks = []                                   # Gauss-point kernels as the K operator
    ref, _ = oracle.run(m, integrator="NEWMARK")
    out = _device(m, options={"integrator": 1.0}).run()[0]
    assert cases.rel_err(out, ref) < cases.TOL_NEWMARK
    from svl_b200.capi import SvlError
    with pytest.raises(SvlError):                   # non-linear material: the tangent is not the elastic stiffness
        _device(cases.j2_column(), options={"integrator": 1.0})
    with pytest.raises(SvlError):                   # PML needs ExtendedNewmarkBeta
        _device(cases.pml2d(), options={"integrator": 1.0})


def test_host_driver_newmark(tmp_path):
    """The C++ host driver reads integrator NEWMARK from the reference's JSON and writes the reference's recorder files."""
    import subprocess
    name = "lysmer_column"
    m = cases.newmark_case(name)
    part = M.write_reference_json(m, str(tmp_path), "Case", "Run", resp=("disp",), ndps=17, integrator="NEWMARK")
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "svl_b200", "SeismoVLAB_gpu.exe")
    subprocess.run([exe, "-dir", part, "-file", "Case.1.$.json"], check=True, stdout=subprocess.DEVNULL)
    out = M.read_node_recorder(os.path.join(str(tmp_path), "Solution", "Run", "disp.0.out"))
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"newmark_{name}.npz"))
    assert cases.rel_err(out, g["disp"]) < cases.TOL_NEWMARK


def test_newmark_device_matches_reference_fixture_j05(oracle):
    """Fixture J05 of the reference's own validation suite (OpenSees golden histories, 6 significant digits) and the
    oracle, on the device: Newmark with both Rayleigh coefficients, Lysmer dashpots, 1000 steps."""
    m = cases.fixture_j05()
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "J05", "opensees.npz"))
    d = _device(m, fields=(0, 1, 2), options={"integrator": 1.0})
    out = d.run()
    for f, key in ((0, "disp"), (1, "vel"), (2, "accel")):
        for col, ours in ((1, 0), (3, 3)):
            ref = g[key][:, col]
            err = np.sqrt(np.mean((out[f][:, ours] - ref) ** 2)) / np.sqrt(np.mean(ref ** 2))
            assert err < 5e-6, (key, col, err)
        ref, _ = oracle.run(m, field=f, integrator="NEWMARK")
        assert cases.rel_err(out[f], ref) < 1e-8, key


@pytest.mark.parametrize("name", ["F02", "F06"])
def test_device_matches_reference_fixture_from_its_own_input_files(oracle, tmp_path, name):
    """The reference's fixtures F02 / F06 as the reference's pre-processor wrote them: (i) through the Python binding,
    (ii) through the C++ host driver reading the JSON + load files and writing the reference's recorder files; both
    against the fixture's OpenSees golden histories and the oracle."""
    import shutil
    import subprocess
    m = cases.fixture_model(name)
    d = _device(m, fields=(0, 1, 2), options={"integrator": 1.0})
    out = d.run()
    for f, key in ((0, "disp"), (1, "vel"), (2, "accel")):
        assert cases.fixture_errors(name, out[f], key) < cases.REF_FIXTURES[name]["tol"], key
        ref, _ = oracle.run(m, field=f, integrator="NEWMARK")
        assert cases.rel_err(out[f], ref) < 1e-8, key
    work = os.path.join(str(tmp_path), name)
    shutil.copytree(cases.fixture_dir(name), work)
    J = __import__("json").load(open(os.path.join(work, "Partition", cases.REF_FIXTURES[name]["json"])))
    combo = J["Combinations"][str(J["Simulations"]["combo"])]["attributes"]["folder"]
    os.makedirs(os.path.join(work, "Solution", combo), exist_ok=True)
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "svl_b200", "SeismoVLAB_gpu.exe")
    subprocess.run([exe, "-dir", os.path.join(work, "Partition"), "-file", cases.REF_FIXTURES[name]["json"].replace(".0.json", ".$.json")],
                   check=True, stdout=subprocess.DEVNULL, cwd=work)
    for key, fn in (("disp", "Displacement.0.out"), ("vel", "Velocity.0.out"), ("accel", "Acceleration.0.out")):
        hist = M.read_node_recorder(os.path.join(work, "Solution", combo, fn))
        assert cases.fixture_errors(name, hist, key) < 2e-5, key          # recorder files carry ndps = 8 digits


def test_host_driver_element_recorders_match_fixture_f02_opensees_stress(tmp_path):
    """The reference's fixture F02 as shipped (NODE + ELEMENT recorders) through the C++ driver: Stress.0.out / Strain.0.out in the
    reference's layout, compared the way the fixture's own cmpResults.py does it (4th Gauss point <-> OpenSees columns 7 8 9)."""
    import json
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "svl_b200", "SeismoVLAB_gpu.exe")
    shutil.copytree(cases.fixture_dir("F02"), str(tmp_path / "fx"))
    part = str(tmp_path / "fx" / "Partition")
    J = json.load(open(os.path.join(part, "Debugging_F02.1.0.json")))
    folder = J["Combinations"][str(J["Simulations"]["combo"])]["attributes"]["folder"]
    os.makedirs(os.path.join(str(tmp_path / "fx"), "Solution", folder))
    r = subprocess.run([exe, "-dir", part, "-file", "Debugging_F02.1.$.json"], capture_output=True, text=True, timeout=300,
                       cwd=str(tmp_path / "fx"), env=dict(os.environ, SVLGPU_ELEMENT_RECORDERS="1"))
    assert r.returncode == 0, r.stdout + r.stderr
    g = np.load(os.path.join(cases.fixture_dir("F02"), "opensees_gauss.npz"))
    stress = np.loadtxt(os.path.join(str(tmp_path / "fx"), "Solution", folder, "Stress.0.out"), skiprows=2)
    rrms = lambda a, b: np.sqrt(np.mean((a - b) ** 2)) / np.sqrt(np.mean(b ** 2))      # noqa: E731
    for ours, col in ((9, 7), (10, 8), (11, 9)):
        assert rrms(stress[:, ours], g["stress"][:, col]) < 5e-6


def test_consistent_mass_fixture_is_refused_loudly():
    from svl_b200.capi import SvlError
    with pytest.raises(SvlError):
        _device(cases.fixture_model("J02"), options={"integrator": 1.0})
