#!/usr/bin/env python
"""Generates the fixtures under tests/golden/ FROM THE REFERENCE ITSELF (run in the build container,
where /root/reference exists and `make -C oracle ref` has produced oracle/_ref/):

  <case>.npz          NODE-recorder histories (disp; kat444 also vel/accel) written by the unmodified
                      reference executable oracle/_ref/SeismoVLAB.exe for tests/cases.py:<case>, run with
                      CentralDifference + Linear + EigenSolver through the reference's own JSON driver.
  element_kat.npz     element / material level vectors from the reference's own classes
                      (oracle/_ref/libsvlref_probe.so): internal forces, M/C/K/G matrices, J2 stress path,
                      DRM element forces.

Usage:  python tests/golden/make_golden.py [case ...]
The fixtures travel to the GPU box; the reference does not.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from svl_b200 import model as M  # noqa: E402
import cases  # noqa: E402
from oracle_lib import RefProbe  # noqa: E402

EXE = os.path.join(ROOT, "oracle", "_ref", "SeismoVLAB.exe")


def run_reference(m, resp=("disp",), integrator="CENTRALDIFFERENCE", newton=None):
    tmp = tempfile.mkdtemp(prefix="svlgold_")
    part = M.write_reference_json(m, tmp, "Case", "Run", resp=resp, ndps=17, integrator=integrator, newton=newton)
    subprocess.run([EXE, "-dir", part, "-file", "Case.1.$.json"], stdout=subprocess.DEVNULL, check=True)
    return {r: M.read_node_recorder(os.path.join(tmp, "Solution", "Run", f"{r}.0.out")) for r in resp}


def history_cases(names):
    for name in names:
        m = cases.CASES[name]()
        resp = ("disp", "vel", "accel") if name == "kat444" else ("disp",)
        out = run_reference(m, resp)
        assert out["disp"].shape[0] == m.nt - 1, (name, out["disp"].shape)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), fingerprint=cases.fingerprint(m),
                            rec_nodes=m.rec_nodes, dt=m.dt, nt=m.nt, **out)
        print(f"{name}: {out['disp'].shape} peak |u| = {np.abs(out['disp']).max():.6e}")


def reaction_cases(names):
    """disp / vel / accel / reaction NODE recorders of the reference executable (Recorder.cpp:246-260) for the reaction and
    support-motion cases"""
    for name in names:
        m = cases.REACTION_CASE_FUNCS[name]()
        out = run_reference(m, ("disp", "vel", "accel", "reaction"))
        assert out["disp"].shape[0] == m.nt - 1, (name, out["disp"].shape)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), fingerprint=cases.fingerprint(m),
                            rec_nodes=m.rec_nodes, dt=m.dt, nt=m.nt, **out)
        print(f"{name}: {out['disp'].shape} peak |u| = {np.abs(out['disp']).max():.6e} peak |R| = {np.abs(out['reaction']).max():.6e}")


def newmark_cases(names):
    for name in names:
        m = cases.newmark_case(name)
        out = run_reference(m, ("disp", "vel", "accel"), integrator="NEWMARK")
        assert out["disp"].shape[0] == m.nt - 1, (name, out["disp"].shape)
        np.savez_compressed(os.path.join(HERE, f"newmark_{name}.npz"), fingerprint=cases.fingerprint(m),
                            rec_nodes=m.rec_nodes, dt=m.dt, nt=m.nt, **out)
        print(f"newmark_{name}: {out['disp'].shape} peak |u| = {np.abs(out['disp']).max():.6e}")


def newton_cases():
    for name, settings in cases.NEWTON_CASES.items():
        m = cases.CASES[name]()
        out = {f"disp_{i}": run_reference(m, ("disp",), integrator="NEWMARK", newton=nw)["disp"] for i, nw in enumerate(settings)}
        np.savez_compressed(os.path.join(HERE, f"newton_{name}.npz"), fingerprint=cases.fingerprint(m), rec_nodes=m.rec_nodes,
                            dt=m.dt, nt=m.nt, settings=np.array(settings, float), **out)
        print(f"newton_{name}: {len(settings)} x {out['disp_0'].shape} peak |u| = {np.abs(out['disp_0']).max():.6e}")


def extended_newmark_cases():
    for name in ("pml2d", "pml3d"):
        m = cases.CASES[name]()
        m.dt *= 2.0
        out = run_reference(m, ("disp",), integrator="EXTENDEDNEWMARK")
        np.savez_compressed(os.path.join(HERE, f"extnewmark_{name}.npz"), fingerprint=cases.fingerprint(m),
                            rec_nodes=m.rec_nodes, dt=m.dt, nt=m.nt, **out)
        print(f"extnewmark_{name}: {out['disp'].shape} peak |u| = {np.abs(out['disp']).max():.6e}")


def element_kat():
    rp = RefProbe()
    rng = np.random.default_rng(20260117)
    cube = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], float)
    sq = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], float)
    G = {}
    Xh = cube * [1.2, 0.9, 1.1] + 0.12 * rng.uniform(-1, 1, (8, 3))
    Uh = 1e-3 * rng.uniform(-1, 1, (8, 3))
    Xq = sq * [1.3, 0.8] + 0.1 * rng.uniform(-1, 1, (4, 2))
    Uq = 1e-3 * rng.uniform(-1, 1, (4, 2))
    G["hex_X"], G["hex_U"], G["quad_X"], G["quad_U"] = Xh, Uh, Xq, Uq
    G["hex_f"] = rp.internal_force(1, Xh, Uh, 1, cases.SOIL)
    mats = rp.matrices(1, Xh, 1, cases.SOIL, lumped=True, want="MK")
    G["hex_Mlumped"], G["hex_K"] = mats["M"], mats["K"]
    G["hex_Mcons"] = rp.matrices(1, Xh, 1, cases.SOIL, lumped=False, want="M")["M"]
    qa = np.zeros(10); qa[0] = 0.7
    G["quad_th"] = 0.7
    G["quad_f"] = rp.internal_force(2, Xq, Uq, 2, cases.SOIL, qa)
    mats = rp.matrices(2, Xq, 2, cases.SOIL, qa, lumped=True, want="MK")
    G["quad_Mlumped"], G["quad_K"] = mats["M"], mats["K"]
    # SURVEY App. B.5 unit-cube known answer (node 7 displaced)
    U7 = np.zeros((8, 3)); U7[6] = [1e-3, 2e-3, -1e-3]
    G["hex_unit_f"] = rp.internal_force(1, cube, U7, 1, cases.SOIL)
    # J2 path: 24 strain increments, loading / unloading / reloading
    eps = np.cumsum(4e-4 * rng.uniform(-1, 1, (24, 6)), axis=0)
    eps[8:16] *= 0.3
    G["j2_eps"] = eps
    G["j2_sig"] = rp.material_path(3, cases.J2, eps)
    # PML matrices: distorted-free cells at two depths, face and corner normals
    p3 = np.array([2.0, 3.0, 1e-5, 1.0, 1.0, 0.0, 0.0, 0.0, -1.0])
    X3 = cube + [0.0, 0.0, -2.0]
    mm = rp.matrices(3, X3, 1, cases.PMLMAT, p3, want="MCKG")
    for k in "MCKG":
        G[f"pml3_face_{k}"] = mm[k]
    G["pml3_face_par"], G["pml3_face_X"] = p3, X3
    p3c = np.array([2.0, 3.0, 1e-5, 0.0, 0.0, 0.0, -1 / np.sqrt(3), -1 / np.sqrt(3), -1 / np.sqrt(3)])
    X3c = cube + [-2.0, -1.0, -3.0]
    mm = rp.matrices(3, X3c, 1, cases.PMLMAT, p3c, want="MCKG")
    for k in "MCKG":
        G[f"pml3_corner_{k}"] = mm[k]
    G["pml3_corner_par"], G["pml3_corner_X"] = p3c, X3c
    p2 = np.array([0.9, 2.0, 3.0, 1e-5, 0.0, 0.0, -1 / np.sqrt(2), -1 / np.sqrt(2)])
    X2 = sq + [-2.0, -1.0]
    mm = rp.matrices(4, X2, 2, cases.PMLMAT, p2, want="MCK")
    for k in "MCK":
        G[f"pml2_{k}"] = mm[k]
    G["pml2_par"], G["pml2_X"] = p2, X2
    # DRM element forces
    ext_h = np.array([0, 0, 1, 1, 0, 1, 1, 0], np.uint8)
    fld_h = 1e-3 * rng.uniform(-1, 1, (8, 9))
    G["drm_hex_ext"], G["drm_hex_field"] = ext_h, fld_h
    G["drm_hex_f"] = rp.drm_force(1, Xh, 1, cases.SOIL, ext_h, fld_h)
    ext_q = np.array([0, 1, 1, 0], np.uint8)
    fld_q = 1e-3 * rng.uniform(-1, 1, (4, 6))
    G["drm_quad_ext"], G["drm_quad_field"] = ext_q, fld_q
    G["drm_quad_f"] = rp.drm_force(2, Xq, 2, cases.SOIL, ext_q, fld_q, qa)
    np.savez_compressed(os.path.join(HERE, "element_kat.npz"), **G)
    print("element_kat:", len(G), "arrays")


if __name__ == "__main__":
    if not os.path.exists(EXE) or not RefProbe.available():
        raise SystemExit("oracle/_ref is missing: run `make -C oracle ref` in the build container first")
    args = sys.argv[1:]
    if args and args[0] == "newmark":
        newmark_cases(args[1:] or list(cases.NEWMARK_CASES))
        extended_newmark_cases()
        newton_cases()
        raise SystemExit(0)
    if args and args[0] == "mid":
        import time
        for name in args[1:] or list(cases.MID_CASES):
            m = cases.MID_CASES[name]()
            t0 = time.time()
            out = run_reference(m, ("disp",))
            np.savez_compressed(os.path.join(HERE, f"{name}.npz"), fingerprint=cases.fingerprint(m), rec_nodes=m.rec_nodes, dt=m.dt, nt=m.nt, **out)
            print(f"{name}: {out['disp'].shape} peak |u| = {np.abs(out['disp']).max():.6e}  ({time.time() - t0:.0f} s in the reference executable)", flush=True)
        raise SystemExit(0)
    if args and args[0] == "reaction":
        reaction_cases(args[1:] or list(cases.REACTION_CASES))
        raise SystemExit(0)
    names = args or list(cases.CASES)
    if not args:
        reaction_cases(list(cases.REACTION_CASES))
        element_kat()
        newmark_cases(list(cases.NEWMARK_CASES))
        extended_newmark_cases()
        newton_cases()
    history_cases(names)
