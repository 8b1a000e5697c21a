#!/usr/bin/env python
"""Regenerates tests/golden/fixtures/<name>/ FROM THE REFERENCE (build container only):

  Partition/<file>.1.0.json   written by the reference's own pre-processor (01-Pre_Process) run on the fixture's script
                              03-Validations/01-Debugging/<fixture>.zip:<fixture>.py, single partition (no METIS needed);
                              matplotlib, which the pre-processor imports at module scope, is replaced by an empty stub
  *.txt / *.in                the fixture's load time series
  opensees.npz                the fixture's OpenSees golden histories (displacement / velocity / acceleration .out)

Usage: python tests/golden/make_fixture_inputs.py
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
            "J02": "J02-DY_Lin_3DPointLoad_Elastic_Hexa8"}


def main():
    tmp = tempfile.mkdtemp(prefix="svlfix_")
    stub = os.path.join(tmp, "stub", "matplotlib")
    os.makedirs(stub)
    open(os.path.join(stub, "__init__.py"), "w").close()
    for mod in ("pyplot", "pylab"):
        with open(os.path.join(stub, mod + ".py"), "w") as f:
            f.write("def __getattr__(n):\n    raise AttributeError(n)\n")
    env = dict(os.environ, PYTHONPATH=os.path.join(REF, "01-Pre_Process") + os.pathsep + os.path.join(tmp, "stub"))
    for name, fx in FIXTURES.items():
        zipfile.ZipFile(os.path.join(REF, "03-Validations", "01-Debugging", fx + ".zip")).extractall(tmp)
        src = os.path.join(tmp, fx)
        subprocess.run([sys.executable, fx + ".py"], cwd=src, env=env, check=True, stdout=subprocess.DEVNULL)
        dst = os.path.join(HERE, "fixtures", name)
        shutil.rmtree(dst, ignore_errors=True)
        os.makedirs(os.path.join(dst, "Partition"))
        for fn in os.listdir(os.path.join(src, "Partition")):
            if fn.endswith(".json"):
                shutil.copy(os.path.join(src, "Partition", fn), os.path.join(dst, "Partition", fn))
        for fn in os.listdir(src):
            if fn.endswith((".txt", ".in")):
                shutil.copy(os.path.join(src, fn), os.path.join(dst, fn))
        o = os.path.join(src, "OpenSees")
        np.savez_compressed(os.path.join(dst, "opensees.npz"), disp=np.loadtxt(os.path.join(o, "displacement.out")),
                            vel=np.loadtxt(os.path.join(o, "velocity.out")), accel=np.loadtxt(os.path.join(o, "acceleration.out")))
        print(name, "->", dst)


if __name__ == "__main__":
    main()
