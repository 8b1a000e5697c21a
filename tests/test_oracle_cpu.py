"""CPU suite (-m "not gpu"): pins the oracle (oracle/svl_oracle.c) against
  (a) the golden NODE-recorder histories written by the UNMODIFIED reference executable
      (tests/golden/<case>.npz, generator tests/golden/make_golden.py),
  (b) element / material vectors produced by the reference's own classes (tests/golden/element_kat.npz),
  (c) the known answers of SURVEY.md App. B.5,
and checks the C-ABI library without touching a GPU."""
import os
import re

import numpy as np
import pytest

import cases
from oracle_lib import RefProbe
from svl_b200 import model as M

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gold(name):
    return np.load(os.path.join(GOLD, f"{name}.npz"))


@pytest.fixture(scope="module")
def kat():
    return gold("element_kat")


# ---- (a) whole-analysis histories --------------------------------------------------------------
@pytest.mark.parametrize("name", list(cases.CASES))
def test_oracle_history_matches_reference_executable(oracle, name):
    m = cases.CASES[name]()
    g = gold(name)
    assert str(g["fingerprint"]) == cases.fingerprint(m), "case generator drifted: regenerate tests/golden"
    out, _ = oracle.run(m)
    assert out.shape == g["disp"].shape
    assert np.abs(g["disp"]).max() > 0
    assert cases.rel_err(out, g["disp"]) < cases.TOL[name]


def test_oracle_matches_reference_executable_on_the_mid_size_j2_column(oracle):
    """configs[4] shape (10 x 10 x 40 lin3DHexa8 + Plastic3DJ2, loaded into yield) against the reference executable.  (The
    mid-size PML2D golden is checked against the oracle by the opt-in test below: 7 minutes of envelope LDL^T.)"""
    m = cases.mid_j2()
    g = gold("mid_j2")
    assert str(g["fingerprint"]) == cases.fingerprint(m)
    out, _ = oracle.run(m, nthreads=8)
    assert cases.rel_err(out, g["disp"]) < cases.TOL["mid_j2"]


@pytest.mark.skipif(os.environ.get("SVL_SLOW_TESTS", "0") != "1", reason="7 minutes of envelope LDL^T on one core: SVL_SLOW_TESTS=1 "
                    "(measured 2.8e-13 against the reference executable's golden, DESIGN.md section 4)")
def test_oracle_matches_reference_executable_on_the_mid_size_quad4_pml_model(oracle):
    """configs[1] shape (200 x 100 lin2DQuad4 + 5-cell PML2DQuad4 layer, 16 000 coupled unknowns) against the reference executable."""
    m = cases.mid_quad4_pml()
    g = gold("mid_quad4_pml")
    assert str(g["fingerprint"]) == cases.fingerprint(m)
    out, _ = oracle.run(m, nthreads=8)
    assert cases.rel_err(out, g["disp"]) < cases.TOL["mid_quad4_pml"]


@pytest.mark.parametrize("name", list(cases.REACTION_CASES))
def test_oracle_reactions_and_support_motion_match_reference_executable(oracle, name):
    """Integrator::ComputeReactionForce rows (DynamicAnalysis.cpp:130-150, CentralDifference.cpp:155-171) and support motion
    (Assembler.cpp:493-533, CentralDifference.cpp:135,189-202) against the disp / vel / accel / reaction NODE recorders the
    unmodified reference executable wrote for the same model."""
    m = cases.REACTION_CASE_FUNCS[name]()
    g = gold(name)
    assert str(g["fingerprint"]) == cases.fingerprint(m), "case generator drifted: regenerate tests/golden"
    for f, key in ((0, "disp"), (1, "vel"), (2, "accel"), (3, "reaction")):
        out, _ = oracle.run(m, field=f)
        assert out.shape == g[key].shape and np.abs(g[key]).max() > 0
        assert cases.rel_err(out, g[key]) < cases.TOL[name], key
    if name.startswith("support"):
        # the support really moves and really shakes the column
        sup_dofs = {m.node_ptr[n] + d for n, d, sr, _ in m.supports if len(sr) > 1}
        cols = [i for i, q in enumerate(m.rec_dofs()) if q in sup_dofs]
        assert cols and np.abs(g["disp"][:, cols]).max() > 1e-4


def test_reference_json_round_trip_with_supports_and_reaction_recorder(oracle, tmp_path):
    from svl_b200 import model as M
    m = cases.support_column()
    part = M.write_reference_json(m, str(tmp_path), "Case", "Run", resp=("disp", "reaction"))
    m2 = M.read_reference_json(os.path.join(part, "Case.1.0.json"))
    assert len(m2.supports) == len(m.supports) and [r for r, _ in m2.rec_spec] == ["disp", "reaction"]
    for f in (0, 3):
        a, _ = oracle.run(m, field=f)
        b, _ = oracle.run(m2, field=f)
        assert np.abs(a - b).max() <= 1e-14 * np.abs(a).max()


@pytest.mark.parametrize("name", list(cases.NEWMARK_CASES))
def test_oracle_newmark_history_matches_reference_executable(oracle, name):
    """NewmarkBeta + Linear (10-Integrators/03-Newmark/NewmarkBeta.cpp) at 4x the explicit step: displacement, velocity
    and acceleration histories of the unmodified reference executable."""
    m = cases.newmark_case(name)
    g = gold(f"newmark_{name}")
    assert str(g["fingerprint"]) == cases.fingerprint(m), "case generator drifted: regenerate tests/golden"
    for f, key in ((0, "disp"), (1, "vel"), (2, "accel")):
        out, _ = oracle.run(m, field=f, integrator="NEWMARK")
        assert out.shape == g[key].shape and np.abs(g[key]).max() > 0
        assert cases.rel_err(out, g[key]) < 1e-10, key


def _j05_errors(hist, g):
    """relative RMS error per OpenSees column (t, node 1 x, node 1 y, node 401 x, node 401 y) against our recorder
    layout (node 1 xyz, node 401 xyz), as 03-Validations/.../LaTeX/cmpResults.py compares them"""
    errs = []
    for col, ours in ((1, 0), (3, 3)):
        ref = g[:, col]
        errs.append(np.sqrt(np.mean((hist[:, ours] - ref) ** 2)) / np.sqrt(np.mean(ref ** 2)))
    return max(errs)


def test_oracle_newmark_matches_reference_fixture_j05_opensees_golden(oracle):
    """The reference's OWN golden vector for this path: fixture J05 (soil column on Lysmer dashpots, Rayleigh damping
    with both coefficients, Newmark) ships OpenSees displacement / velocity / acceleration histories printed with
    6 significant digits."""
    m = cases.fixture_j05()
    g = np.load(os.path.join(GOLD, "J05", "opensees.npz"))
    for f, key, tol in ((0, "disp", 5e-6), (1, "vel", 5e-6), (2, "accel", 5e-6)):
        out, _ = oracle.run(m, field=f, integrator="NEWMARK")
        assert out.shape == (999, 6)
        assert _j05_errors(out, g[key]) < tol, key


@pytest.mark.parametrize("name", list(cases.REF_FIXTURES))
def test_oracle_matches_reference_fixture_from_its_own_input_files(oracle, name):
    """Fixtures F02 / F06 / J02 of the reference's validation suite, read from the JSON the reference's pre-processor wrote,
    against the OpenSees golden histories the fixtures ship (6 significant digits)."""
    m = cases.fixture_model(name)
    assert m.integrator == "NEWMARK"
    for f, key in ((0, "disp"), (1, "vel"), (2, "accel")):
        out, _ = oracle.run(m, field=f, integrator=m.integrator)
        assert cases.fixture_errors(name, out, key) < cases.REF_FIXTURES[name]["tol"], key


@pytest.mark.parametrize("name", ["J02", "F02"])
def test_oracle_gauss_point_strain_stress_match_reference_fixture_opensees_recorders(oracle, name):
    """Fixtures J02 (one lin3DHexa8) and F02 (one lin2DQuad4) ship the OpenSees ELEMENT recorder files strain.out / stress.out.
    The oracle's kinematics (B matrix, Gauss-point order, Voigt order) and elastic law applied to the oracle's own nodal
    history reproduce them at the Gauss point and in the column pairing the fixtures' LaTeX/cmpResults.py use
    (J02: 8th point, SVL [11 22 33 12 23 13] <-> OpenSees columns 43 45 44 48 47 46; F02: 4th point <-> columns 7 8 9),
    to the 6 printed digits."""
    import ctypes as C
    dp = C.POINTER(C.c_double)
    m = cases.fixture_model(name)
    g = np.load(os.path.join(cases.fixture_dir(name), "opensees_gauss.npz"))
    out, _ = oracle.run(m, integrator=m.integrator, rec_dofs=np.arange(m.n_total, dtype=np.int32))
    assert m.n_elem == 1 and out.shape[0] == g["strain"].shape[0]
    is3 = m.ndim == 3
    npe, ngp, ncomp = (8, 8, 6) if is3 else (4, 4, 3)
    conn = m.elem_conn[0, :npe]
    X = np.ascontiguousarray(m.coords[conn].ravel())
    Cm = np.zeros(ncomp * ncomp)
    E, nu = m.materials[0][1][:2]
    (oracle.lib.svlo_elastic3d_C if is3 else oracle.lib.svlo_planestrain_C)(C.c_double(E), C.c_double(nu), Cm.ctypes.data_as(dp))
    Cm = Cm.reshape(ncomp, ncomp)
    eps = np.zeros((out.shape[0], ngp, ncomp))
    for k in range(out.shape[0]):
        Ue = np.ascontiguousarray(np.concatenate([out[k, m.node_ptr[n]:m.node_ptr[n] + m.ndim] for n in conn]))
        e = np.zeros((ngp, ncomp))
        (oracle.lib.svlo_hex8_strain if is3 else oracle.lib.svlo_quad4_strain)(X.ctypes.data_as(dp), Ue.ctypes.data_as(dp),
                                                                              e.ctypes.data_as(dp))
        eps[k] = e
    sig = eps @ Cm.T
    gp, cols = (7, (43, 45, 44, 48, 47, 46)) if is3 else (3, (7, 8, 9))
    rrms = lambda a, b: np.sqrt(np.mean((a - b) ** 2)) / np.sqrt(np.mean(b ** 2))      # noqa: E731
    for c, col in enumerate(cols):
        assert rrms(sig[:, gp, c], g["stress"][:, col]) < 5e-6, ("stress", c)
        if is3:                                                                          # F02's script compares stresses only
            assert rrms(eps[:, gp, c], g["strain"][:, col]) < 5e-6, ("strain", c)


@pytest.mark.parametrize("name", list(cases.EXE_FIXTURES))
def test_oracle_matches_reference_executable_on_pml_fixture_inputs(oracle, name):
    """Fixtures F11 (PML2DQuad4, Newmark) and J12 (PML3DHexa8, ExtendedNewmarkBeta with the history matrix G), read from
    the reference pre-processor's JSON (EQUAL constraints, consistent mass), against the unmodified reference executable
    run on the same files.  Accelerations carry 4/dt^2 of the displacement rounding."""
    m = cases.fixture_model(name)
    g = np.load(os.path.join(cases.fixture_dir(name), "reference.npz"))
    assert m.integrator in ("NEWMARK", "EXTENDEDNEWMARK") and len(m.constraints) > 0 and not m.lumped
    for f, key, tol in ((0, "disp", 1e-10), (1, "vel", 1e-8), (2, "accel", 1e-6)):
        out, _ = oracle.run(m, field=f, integrator=m.integrator, nthreads=4)
        assert out.shape == g[key].shape
        assert cases.rel_err(out, g[key]) < tol, key


@pytest.mark.parametrize("name", ["pml2d", "pml3d"])
def test_oracle_extended_newmark_matches_reference_executable(oracle, name):
    m = cases.CASES[name]()
    m.dt *= 2.0
    g = gold(f"extnewmark_{name}")
    assert str(g["fingerprint"]) == cases.fingerprint(m)
    out, _ = oracle.run(m, integrator="EXTENDEDNEWMARK")
    assert cases.rel_err(out, g["disp"]) < 1e-10


@pytest.mark.parametrize("name", list(cases.NEWTON_CASES))
def test_oracle_newton_raphson_matches_reference_executable(oracle, name):
    """NewmarkBeta + NewtonRaphson with the consistent tangents of Plastic3DJ2 / PlasticPlaneStrainJ2, one setting per
    convergence test of Algorithm.cpp:122-186 (unbalanced force, increment, energy, their relative forms, total relative
    increment, fixed iteration count), against the unmodified reference executable."""
    m = cases.CASES[name]()
    g = gold(f"newton_{name}")
    assert str(g["fingerprint"]) == cases.fingerprint(m)
    tests_seen = set()
    for i, nw in enumerate(cases.NEWTON_CASES[name]):
        assert tuple(g["settings"][i]) == tuple(map(float, nw))
        out, _ = oracle.run(m, integrator="NEWMARK", newton=nw)
        assert cases.rel_err(out, g[f"disp_{i}"]) < cases.TOL_NEWTON, nw
        tests_seen.add(nw[2])
    lin, _ = oracle.run(m, integrator="NEWMARK")                   # the iteration matters: Linear gives another history
    assert cases.rel_err(lin, g["disp_0"]) > 1e-6
    assert len(tests_seen) == 4


@pytest.mark.parametrize("name", list(cases.NEWTON_FIXTURES))
def test_oracle_newton_raphson_matches_reference_fixtures_f03_f07(oracle, name):
    """The reference's own Newton fixtures, read from the JSON its pre-processor wrote: F03 (one PlasticPlaneStrainJ2 quad loaded
    far into yield, no mass) and F07 (100-quad J2 soil column on Lysmer dashpots, stiffness-proportional Rayleigh damping from
    the INITIAL tangent, lin2DQuad4.cpp:363).  Against the reference executable run on the same files the bound is set by
    quirk q10 (F07: a Gauss point whose trial state sits on the yield surface to the last bit picks its tangent branch by
    rounding); against the fixtures' OpenSees histories the bounds are what the reference itself achieves."""
    spec = cases.NEWTON_FIXTURES[name]
    m = cases.fixture_model(name)
    assert m.integrator == "NEWMARK" and m.newton is not None and m.newton[2] == 5
    ref = np.load(os.path.join(cases.fixture_dir(name), "reference.npz"))
    for f, key, amp in ((0, "disp", 1.0), (1, "vel", 10.0), (2, "accel", 100.0)):
        out, _ = oracle.run(m, field=f, integrator=m.integrator, newton=m.newton)
        assert out.shape == ref[key].shape
        assert cases.rel_err(out, ref[key]) < spec["tol_exe"] * amp, key
        assert cases.fixture_errors(name, out, key) < spec["tol_os"][f], key
        assert abs(cases.fixture_errors(name, out, key) - cases.fixture_errors(name, ref[key], key)) < spec["tol_exe"] * amp


@pytest.mark.parametrize("name,par", [("F03", [1.333333e8, 8.0e7, 0.0, 8.0e7, 1.0, 4.0e7]),
                                      ("F07", [2.9e7, 2.0e7, 2000.0, 1.0e7, 1.0, 1.0e4])])
def test_j2_plane_strain_return_map_matches_reference_fixture_stress_paths(oracle, name, par):
    """PlasticPlaneStrainJ2 (the embedding of PlasticPlaneStrainJ2.cpp:235 around the 3-D return map) driven by the
    Gauss-point strain histories of the reference's fixtures F03 / F07 reproduces their OpenSees stress histories
    (6 printed digits, accumulated along a plastic path).  F03 is loaded well into yield (|s| reaches Sy)."""
    import ctypes as C
    g = np.load(os.path.join(GOLD, "fixtures", "J2PS", "opensees_stress_strain.npz"))
    dp = C.POINTER(C.c_double)
    p = np.array(par, float)
    for gp in (0, 2):
        strain, ref = g[f"{name}_strain_gp{gp}"], g[f"{name}_stress_gp{gp}"]
        st = np.zeros(13)
        sig = np.zeros_like(ref)
        yielded = 0
        for k, e in enumerate(strain):
            e6 = np.array([0.0, e[0], e[1], 0.0, e[2], 0.0]); s6 = np.zeros(6); before = st[12]
            oracle.lib.svlo_j2_update(p.ctypes.data_as(dp), e6.ctypes.data_as(dp), st.ctypes.data_as(dp), s6.ctypes.data_as(dp))
            yielded += st[12] != before
            sig[k] = (s6[1], s6[2], s6[4])
        err = np.abs(sig - ref).max(axis=0) / np.abs(ref).max(axis=0)
        assert err.max() < 3e-5, (gp, err)
        if name == "F03":
            assert yielded > 5


@pytest.mark.parametrize("name", ["kat444", "lysmer_area", "pml2d", "hex8_layered_rayleigh", "j2ps_area", "drm_box"])
def test_reference_json_writer_reader_round_trip(oracle, tmp_path, name):
    """write_reference_json -> read_reference_json gives a model the oracle advances to the same history (the reader
    renumbers nodes / dofs by ascending tag and maps the constraints' slave-total / master-free dofs)."""
    from svl_b200 import model as M
    m = cases.CASES[name]()
    part = M.write_reference_json(m, str(tmp_path), "Case", "Run")
    m2 = M.read_reference_json(os.path.join(part, "Case.1.0.json"))
    assert m2.n_nodes == m.n_nodes and m2.n_elem == m.n_elem and m2.n_free == m.n_free and len(m2.constraints) == len(m.constraints)
    assert m2.integrator == "CENTRALDIFFERENCE" and m2.dt == m.dt and m2.nt == m.nt
    a, _ = oracle.run(m)
    b, _ = oracle.run(m2)
    assert a.shape == b.shape and np.abs(a).max() > 0
    assert np.abs(a - b).max() <= 1e-14 * np.abs(a).max()


def test_oracle_vel_accel_match_reference_executable(oracle):
    m = cases.kat444()
    g = gold("kat444")
    for f, key in ((1, "vel"), (2, "accel")):
        out, _ = oracle.run(m, field=f)
        assert cases.rel_err(out, g[key]) < 1e-9


def test_survey_b5_central_difference_known_answer(oracle):
    out, _ = oracle.run(cases.kat444())
    assert abs(out[0, 3] - 9.888543819998318e-06) < 1e-19
    assert np.allclose(out[24], [1.305711949530492e-06, -6.528559747652594e-07, -1.540661613904458e-04,
                                 5.382280138183359e-04, -2.691140069091679e-04, 2.370556193531373e-03,
                                 1.165693300755441e-04, -1.614276542374741e-04, -2.752663997074770e-05],
                       rtol=1e-10, atol=1e-18)
    assert np.allclose(out[49, 3:6], [2.280569168526373e-04, -1.140284584263187e-04, 1.169099120176147e-03],
                       rtol=1e-10, atol=0)


# ---- (b) element / material level -----------------------------------------------------------------
def test_hex8_element_vectors(oracle, kat):
    E, nu, rho = cases.SOIL
    f = oracle.hex8_elastic_force(kat["hex_X"], kat["hex_U"], E, nu)
    assert np.abs(f - kat["hex_f"]).max() < 1e-12 * np.abs(kat["hex_f"]).max()
    assert np.abs(oracle.hex8_stiffness(kat["hex_X"], E, nu) - kat["hex_K"]).max() < 1e-12 * np.abs(kat["hex_K"]).max()
    assert np.abs(oracle.hex8_mass(kat["hex_X"], rho, True) - kat["hex_Mlumped"]).max() < 1e-10
    assert np.abs(oracle.hex8_mass(kat["hex_X"], rho, False) - kat["hex_Mcons"]).max() < 1e-10
    cube = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], float)
    U7 = np.zeros((8, 3)); U7[6] = [1e-3, 2e-3, -1e-3]
    f7 = oracle.hex8_elastic_force(cube, U7, E, nu)
    assert np.abs(f7 - kat["hex_unit_f"]).max() < 1e-9
    # SURVEY App. B.5 literal values (first / last entries) and lumped mass 250
    assert abs(f7[0] + 1.284722222222221e+03) < 1e-9 and abs(f7[23] - 5.555555555555557e+02) < 1e-9
    assert np.allclose(np.diag(oracle.hex8_mass(cube, rho, True)), 250.0, rtol=1e-13)


def test_quad4_element_vectors(oracle, kat):
    E, nu, rho = cases.SOIL
    th = float(kat["quad_th"])
    f = oracle.quad4_elastic_force(kat["quad_X"], kat["quad_U"], th, E, nu)
    assert np.abs(f - kat["quad_f"]).max() < 1e-12 * np.abs(kat["quad_f"]).max()
    assert np.abs(oracle.quad4_stiffness(kat["quad_X"], th, E, nu) - kat["quad_K"]).max() < 1e-12 * np.abs(kat["quad_K"]).max()
    assert np.abs(oracle.quad4_mass(kat["quad_X"], th, rho, True) - kat["quad_Mlumped"]).max() < 1e-10
    sq = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], float)
    U3 = np.zeros((4, 2)); U3[2] = [1e-3, -2e-3]
    f3 = oracle.quad4_elastic_force(sq, U3, 1.0, E, nu)
    assert np.allclose(f3, [2.5e3, 4.375e3, 0.0, 9.375e3, 1.25e3, -1.1875e4, -3.75e3, -1.875e3], rtol=1e-12, atol=1e-8)


def test_j2_return_map_path(oracle, kat):
    sig, _ = oracle.j2_path(cases.J2, kat["j2_eps"])
    assert np.abs(sig - kat["j2_sig"]).max() < 1e-11 * np.abs(kat["j2_sig"]).max()
    # the path must actually yield, otherwise this only tests elasticity
    K, G = cases.J2[0], cases.J2[1]
    e = kat["j2_eps"][-1]
    tr = e[:3].sum()
    el = np.concatenate([K * tr + 2 * G * (e[:3] - tr / 3), G * e[3:]])
    assert np.abs(sig[-1] - el).max() > 1e-3 * np.abs(el).max()
    # SURVEY App. B.5 two-step known answer
    s2, _ = oracle.j2_path(cases.J2, [[1e-3, -2e-4, 3e-4, 8e-4, -5e-4, 2e-4], [1.5e-3, -1e-4, 2e-4, -4e-4, -9e-4, 6e-4]])
    assert np.allclose(s2[0], [3.957938883279864e+04, 2.502896788644333e+04, 3.109164328075804e+04,
                               4.850140315451767e+03, -3.031337697157354e+03, 1.212535078862942e+03], rtol=1e-12)
    assert np.allclose(s2[1], [5.493074194109687e+04, 4.211363736707318e+04, 4.215562069182994e+04,
                               -6.324585637498332e+03, -4.356109137812740e+03, 3.476490645298842e+03], rtol=1e-12)


def test_pml_matrices(oracle, kat):
    E, nu, rho = cases.PMLMAT
    for tag in ("face", "corner"):
        got = dict(zip("MCKG", oracle.pml3d(kat[f"pml3_{tag}_X"], E, nu, rho, kat[f"pml3_{tag}_par"])))
        for k in "MCKG":
            ref = kat[f"pml3_{tag}_{k}"]
            scale = max(np.abs(ref).max(), 1e-300)
            assert np.abs(got[k] - ref).max() <= 1e-12 * scale, (tag, k)
    got = dict(zip("MCK", oracle.pml2d(kat["pml2_X"], E, nu, rho, kat["pml2_par"])))
    for k in "MCK":
        ref = kat[f"pml2_{k}"]
        assert np.abs(got[k] - ref).max() <= 1e-12 * np.abs(ref).max(), k
    # SURVEY App. B.5 literal values
    cube = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], float)
    Mm, Cc, Kk, Gg = oracle.pml3d(cube, 5e7, 0.25, 2000.0, [2, 10, 1e-5, 0.5, 0.5, 1.0, 0, 0, -1])
    assert abs(Mm[0, 0] - 7.482028220606291e+01) < 1e-11 and abs(Mm[3, 3] + 7.482028220606293e-10) < 1e-21
    assert abs(Cc[0, 0] - 1.292470397625685e+02) < 1e-10 and abs(Cc[0, 3] + 5.611521165454719e-02) < 1e-14
    assert abs(np.linalg.norm(Kk) - 1.155283173875213e+00) < 1e-12 and np.linalg.norm(Gg) == 0.0


def test_drm_element_forces(oracle, kat):
    E, nu, rho = cases.SOIL
    f = oracle.hex8_drm(kat["hex_X"], E, nu, rho, True, kat["drm_hex_ext"], kat["drm_hex_field"])
    assert np.abs(f - kat["drm_hex_f"]).max() < 1e-12 * np.abs(kat["drm_hex_f"]).max()
    f = oracle.quad4_drm(kat["quad_X"], float(kat["quad_th"]), E, nu, rho, True, kat["drm_quad_ext"], kat["drm_quad_field"])
    assert np.abs(f - kat["drm_quad_f"]).max() < 1e-12 * np.abs(kat["drm_quad_f"]).max()


@pytest.mark.skipif(not RefProbe.available(), reason="reference probe only exists in the build container")
def test_oracle_against_live_reference_classes(oracle):
    rp = RefProbe()
    rng = np.random.default_rng(7)
    cube = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], float)
    for _ in range(5):
        X = cube + 0.2 * rng.uniform(-1, 1, (8, 3))
        U = 1e-3 * rng.uniform(-1, 1, (8, 3))
        a = oracle.hex8_elastic_force(X, U, *cases.SOIL[:2])
        b = rp.internal_force(1, X, U, 1, cases.SOIL)
        assert np.abs(a - b).max() < 1e-12 * np.abs(b).max()


# ---- host-side model logic ----------------------------------------------------------------------
def test_structured_numbering_follows_builder():
    m = M.make_box_model((1, 1, 3), 1.0)
    # Builder.py:174-183: a 1x1xk column gives conn [1,2,4,3,5,6,8,7] (1-based)
    assert list(m.elem_conn[0] + 1) == [1, 2, 4, 3, 5, 6, 8, 7]
    assert m.n_total == 48 and m.n_free == 36


def test_pml_generator_topology():
    m = M.make_pml_model((4, 3), 2, 1.0)
    assert m.n_soil_nodes == 20 and (m.elem_kind[:12] == M.LIN2DQUAD4).all() and (m.elem_kind[12:] == M.PML2DQUAD4).all()
    # every constraint ties a PML displacement dof to the soil dof at the same location
    for tag, slave, masters, factors in m.constraints:
        assert tag < -1 and factors == [1.0]
        node_s = np.searchsorted(m.node_ptr, slave, side="right") - 1
        free = np.asarray(m.freedof_flat)
        master_total = int(np.nonzero(free == masters[0])[0][0])
        node_m = np.searchsorted(m.node_ptr, master_total, side="right") - 1
        assert node_s >= m.n_soil_nodes > node_m
        assert np.allclose(m.coords[node_s], m.coords[node_m])


# ---- C ABI (no GPU here) --------------------------------------------------------------------------
def test_cabi_exports_every_declared_symbol():
    from svl_b200 import capi
    lib = capi.load_library()
    hdr = open(os.path.join(os.path.dirname(GOLD), "..", "include", "svlgpu.h")).read()
    declared = sorted(set(re.findall(r"\b(svlgpu_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 30
    missing = [n for n in declared if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(capi.EXPORTS) == declared


def test_cabi_builder_argument_errors_and_no_cpu_fallback():
    import ctypes as C
    from svl_b200 import capi
    L = capi.load_library()
    assert not L.svlgpu_create(4, 1)
    assert b"ndim" in L.svlgpu_last_error()
    h = L.svlgpu_create(3, 1)
    assert h
    assert L.svlgpu_finalize(h, 0.01, 0) != 0           # empty model is refused
    assert b"empty" in L.svlgpu_last_error()
    L.svlgpu_destroy(h)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        m = cases.kat444()
        with pytest.raises(capi.SvlError, match="no CUDA device|no usable CUDA"):
            capi.DeviceModel(m)


def test_bench_reads_the_hbm_peak_from_any_reasonable_measured_peaks_schema():
    """bench.py takes roofline.peak from the driver-written MEASURED_PEAKS.json, whose schema this repo does not control."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    assert b.pick_hbm_peak({"hbm_gbs": 6540.0}) == (6540.0, "hbm_gbs")
    assert b.pick_hbm_peak({"hbm_copy_gbs": {"burst": 7100, "sustained": 6540}, "bf16_dense_tflops": 1500})[0] == 6540.0
    assert b.pick_hbm_peak({"peaks": {"hbm_tb_s": 6.54, "bf16_tf_s": 1600}})[0] == pytest.approx(6540.0)
    assert b.pick_hbm_peak({"copy_bandwidth_GBps": 6600, "cublas_bf16_TFLOPs": 1700})[0] == 6600.0
    assert b.pick_hbm_peak({"bf16_tflops": 1700}) is None and b.pick_hbm_peak({}) is None
    # a key that names HBM wins over an unrelated "*bandwidth*" one, and an out-of-range first candidate does not end the search
    assert b.pick_hbm_peak({"l2_bandwidth_gbs": 9000.0, "hbm_gbs": 6551.0})[0] == 6551.0
    assert b.pick_hbm_peak({"nvlink_bandwidth_gbs": 900.0, "dram_copy_gbs": 6500.0})[0] == 6500.0


def test_interior_hex8_stencil_is_a_sum_of_tensor_products(oracle):
    """Groundwork for the next kernel round: on a uniform lattice of cubes the 27 x (3 x 3) stencil the
    block-stencil kernel applies at an interior node -- assembled here from the oracle's lin3DHexa8 stiffness (lin3DHexa8.cpp:288-318)
    over the 8 surrounding elements -- equals, to rounding,
        K_aa = (lam + 2 mu) S_a M_b M_c + mu (M_a S_b M_c + M_a M_b S_c),      K_ab = -(lam + mu) D_a D_b M_c   (a != b)
    with the 1-D stencils M = h/6 [1 4 1], S = 1/h [-1 2 -1], D = 1/2 [-1 0 1]: about 70 FP64 instructions per node when applied
    direction by direction, instead of the 153 DFMA of the symmetric 27-point form."""
    import ctypes as C
    dp = C.POINTER(C.c_double)
    E, nu, h = 1.3e7, 0.3, 0.7
    lam, mu = E * nu / ((1 + nu) * (1 - 2 * nu)), E / (2 * (1 + nu))
    Cm = np.zeros(36)
    oracle.lib.svlo_elastic3d_C(C.c_double(E), C.c_double(nu), Cm.ctypes.data_as(dp))
    nid = lambda i, j, k: i + 3 * j + 9 * k                                  # noqa: E731
    pos = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
    K = np.zeros((81, 81))
    for ck in range(2):
        for cj in range(2):
            for ci in range(2):
                nodes = [nid(ci + a, cj + b, ck + c) for a, b, c in pos]
                X = np.array([[(ci + a) * h, (cj + b) * h, (ck + c) * h] for a, b, c in pos]).ravel()
                Ke = np.zeros(576)
                oracle.lib.svlo_hex8_stiffness(X.ctypes.data_as(dp), Cm.ctypes.data_as(dp), Ke.ctypes.data_as(dp))
                Ke = Ke.reshape(24, 24)
                for p, n_p in enumerate(nodes):
                    for q, n_q in enumerate(nodes):
                        K[3 * n_p:3 * n_p + 3, 3 * n_q:3 * n_q + 3] += Ke[3 * p:3 * p + 3, 3 * q:3 * q + 3]
    c = nid(1, 1, 1)
    st = np.zeros((3, 3, 3, 3, 3))                                           # [di+1][dj+1][dk+1][a][b]
    for k in range(3):
        for j in range(3):
            for i in range(3):
                st[i, j, k] = K[3 * c:3 * c + 3, 3 * nid(i, j, k):3 * nid(i, j, k) + 3]
    M1, S1, D1 = h / 6 * np.array([1, 4, 1.0]), 1 / h * np.array([-1, 2, -1.0]), 0.5 * np.array([-1, 0, 1.0])
    T = lambda a, b, c_: np.einsum("i,j,k->ijk", a, b, c_)                  # noqa: E731
    F = np.zeros_like(st)
    for a in range(3):
        for b in range(3):
            t = [M1, M1, M1]
            if a == b:
                t[a] = S1
                F[..., a, a] += (lam + 2 * mu) * T(*t)
            else:
                t[b] = S1
                F[..., a, a] += mu * T(*t)
                t = [M1, M1, M1]
                t[a] = D1
                t[b] = D1
                F[..., a, b] += -(lam + mu) * T(*t)
    assert np.abs(F - st).max() <= 1e-14 * np.abs(st).max()


def test_every_runtime_option_of_the_library_is_documented_in_the_header():
    """svlgpu_set_option: the names cabi.cu accepts and the names include/svlgpu.h documents are the same set."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "svl_b200", "csrc", "cabi.cu")).read()
    body = src[src.index("int svlgpu_set_option"):]
    body = body[:body.index("GUARD_END")]
    accepted = set(re.findall(r'n == "([a-z_0-9]+)"', body))
    hdr = open(os.path.join(root, "include", "svlgpu.h")).read()
    doc = hdr[hdr.index("Planner / solver options"):hdr.index("int svlgpu_set_option")]
    documented = set(re.findall(r'"([a-z_0-9]+)"', doc))
    assert accepted and accepted == documented, (accepted - documented, documented - accepted)
