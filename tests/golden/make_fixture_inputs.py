#!/usr/bin/env python
"""Regenerates tests/golden/fixtures/<name>/ FROM THE REFERENCE (build container only):

  Partition/<file>.1.0.json   written by the reference's own pre-processor (01-Pre_Process) run on the fixture's script
                              03-Validations/01-Debugging/<fixture>.zip:<fixture>.py, single partition (no METIS needed);
                              matplotlib, which the pre-processor imports at module scope, is replaced by an empty stub
  *.txt / *.in                the fixture's load time series
  opensees.npz                the fixture's OpenSees golden histories (displacement / velocity / acceleration .out)
  opensees_gauss.npz          J02 / F02: the OpenSees element recorder files strain.out / stress.out (Gauss points of the one element)
  reference.npz               fixtures without shipped numbers (F11, J12: the reference validates them by a plot) and the
                              Newton-Raphson fixtures (F03, F07): NODE
                              recorder histories written by the unmodified reference executable oracle/_ref/SeismoVLAB.exe
                              run on exactly these input files (PARAVIEW recorder removed, 17 digits)

Usage: python tests/golden/make_fixture_inputs.py [fixture names, default all]
"""
import os
import shutil
import subprocess
import sys
import tempfile
import zipfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
FIXTURES = {"F02": "F02-DY_Lin_2DPointLoad_ElasticPStrain_Quad4", "F06": "F06-DY_Lin_2DSoilColumn_ElasticPStrain_Quad4",
            "J02": "J02-DY_Lin_3DPointLoad_Elastic_Hexa8", "F11": "F11-DY_Lin_2DPMLSoilColumn_ElasticPStrain_Quad4",
            "J12": "J12-DY_Axial_Load_Long_Rod_PML3D",
            "F03": "F03-DY_Lin_2DPointLoad_J2PStrain_Quad4", "F07": "F07-DY_Lin_2DSoilColumn_J2PStrain_Quad4"}
# NewmarkBeta + NewtonRaphson on PlasticPlaneStrainJ2: OpenSees histories AND the reference executable's (the reference's
# Newton iteration updates the LIVE material state, so it is not OpenSees' algorithm: SURVEY.md App. C q9)
NEWTON_FIXTURES = ("F03", "F07")
METIS_FIXTURES = ("F11", "J12", "F06")          # keep the reference's METIS input file (EQUAL ties collapse slave onto master)
EXE = os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "SeismoVLAB.exe")


def run_reference_on(src, dst):
    """NODE recorder histories of the unmodified reference executable on the fixture's own input files"""
    import json
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from svl_b200 import model as M
    jp = [os.path.join(src, "Partition", f) for f in os.listdir(os.path.join(src, "Partition")) if f.endswith(".json")][0]
    J = json.load(open(jp))
    J["Recorders"] = {k: dict(v, ndps=17) for k, v in J["Recorders"].items() if v["name"] == "NODE"}
    run_json = jp.replace(".1.0.json", "R.1.0.json")
    json.dump(J, open(run_json, "w"))
    combo = J["Combinations"][str(J["Simulations"]["combo"])]["attributes"]["folder"]
    os.makedirs(os.path.join(src, "Solution", combo), exist_ok=True)
    subprocess.run([EXE, "-dir", os.path.join(src, "Partition"), "-file", os.path.basename(run_json).replace(".0.json", ".$.json")],
                   cwd=src, check=True, stdout=subprocess.DEVNULL)
    os.remove(run_json)
    out = {}
    for r in J["Recorders"].values():
        out[r["resp"]] = M.read_node_recorder(os.path.join(src, "Solution", combo, r["file"]))
    np.savez_compressed(os.path.join(dst, "reference.npz"), **out)


def j2_plane_strain_paths(tmp):
    """Gauss-point strain / stress histories (points 1 and 3 of element 1) of the two PlasticPlaneStrainJ2 fixtures: the
    OpenSees element recorders the fixtures ship.  Material-level golden: the integrator (Newmark + Newton) does not matter."""
    out = {}
    for name, fx in (("F03", "F03-DY_Lin_2DPointLoad_J2PStrain_Quad4"), ("F07", "F07-DY_Lin_2DSoilColumn_J2PStrain_Quad4")):
        zipfile.ZipFile(os.path.join(REF, "03-Validations", "01-Debugging", fx + ".zip")).extractall(tmp)
        e = np.loadtxt(os.path.join(tmp, fx, "OpenSees", "strain.out"))
        s = np.loadtxt(os.path.join(tmp, fx, "OpenSees", "stress.out"))
        for g in (0, 2):
            out[f"{name}_strain_gp{g}"] = e[:, 1 + 3 * g:4 + 3 * g]      # e11, e22, gamma12 (engineering)
            out[f"{name}_stress_gp{g}"] = s[:, 1 + 3 * g:4 + 3 * g]
    os.makedirs(os.path.join(HERE, "fixtures", "J2PS"), exist_ok=True)
    np.savez_compressed(os.path.join(HERE, "fixtures", "J2PS", "opensees_stress_strain.npz"), **out)
    print("J2PS ->", os.path.join(HERE, "fixtures", "J2PS"))


def main():
    tmp = tempfile.mkdtemp(prefix="svlfix_")
    j2_plane_strain_paths(tmp)
    stub = os.path.join(tmp, "stub", "matplotlib")
    os.makedirs(stub)
    open(os.path.join(stub, "__init__.py"), "w").close()
    for mod in ("pyplot", "pylab"):
        with open(os.path.join(stub, mod + ".py"), "w") as f:
            f.write("def __getattr__(n):\n    raise AttributeError(n)\n")
    env = dict(os.environ, PYTHONPATH=os.path.join(REF, "01-Pre_Process") + os.pathsep + os.path.join(tmp, "stub"))
    only = set(sys.argv[1:])
    for name, fx in FIXTURES.items():
        if only and name not in only:
            continue
        zipfile.ZipFile(os.path.join(REF, "03-Validations", "01-Debugging", fx + ".zip")).extractall(tmp)
        src = os.path.join(tmp, fx)
        subprocess.run([sys.executable, fx + ".py"], cwd=src, env=env, check=True, stdout=subprocess.DEVNULL)
        if name in METIS_FIXTURES:
            # the METIS mesh file the reference's pre-processor writes for this model (Core/Partition.py:87-144 SetMetisInputFile;
            # createPartitions deletes it again, so it is rebuilt here from the script's entities with the cluster map cleared)
            code = ("import runpy, sys\nsys.argv=['x']\nrunpy.run_path(%r, run_name='__main__')\n"
                    "from Core.Definitions import Options\nfrom Core.Partition import SetMetisInputFile\n"
                    "Options['clustermap'] = {}\nSetMetisInputFile()\n" % (fx + ".py"))
            subprocess.run([sys.executable, "-c", code], cwd=src, env=env, check=True, stdout=subprocess.DEVNULL)
        dst = os.path.join(HERE, "fixtures", name)
        shutil.rmtree(dst, ignore_errors=True)
        os.makedirs(os.path.join(dst, "Partition"))
        for fn in os.listdir(os.path.join(src, "Partition")):
            if fn.endswith(".json"):
                shutil.copy(os.path.join(src, "Partition", fn), os.path.join(dst, "Partition", fn))
        if name in METIS_FIXTURES:
            shutil.copy(os.path.join(src, "Partition", "Graph.out"), os.path.join(dst, "Graph.out"))
        for fn in os.listdir(src):
            if fn.endswith((".txt", ".in")):
                shutil.copy(os.path.join(src, fn), os.path.join(dst, fn))
        o = os.path.join(src, "OpenSees")
        if os.path.isdir(o):
            np.savez_compressed(os.path.join(dst, "opensees.npz"), disp=np.loadtxt(os.path.join(o, "displacement.out")),
                                vel=np.loadtxt(os.path.join(o, "velocity.out")), accel=np.loadtxt(os.path.join(o, "acceleration.out")))
        if os.path.exists(os.path.join(o, "stress.out")) and name in ("J02", "F02"):
            # Gauss-point histories of the single element (OpenSees element recorder: time, then ngp x ncomp columns)
            np.savez_compressed(os.path.join(dst, "opensees_gauss.npz"), strain=np.loadtxt(os.path.join(o, "strain.out")),
                                stress=np.loadtxt(os.path.join(o, "stress.out")))
        if not os.path.isdir(o) or name in NEWTON_FIXTURES:
            run_reference_on(src, dst)
        print(name, "->", dst)


if __name__ == "__main__":
    main()
