"""What reaches the C ABI, checked on CPU: the C++ host driver and the ctypes binding run against a RECORDING stand-in for
libsvlgpu.so (tests/trace_shim/svlgpu_trace.c: every entry point of include/svlgpu.h, one trace line per builder call with
digests of the array arguments, no GPU and no physics).  The same model must produce the same calls
  * from a plain partition JSON and from its binary-table twin,
  * on several ranks: from the C++ driver reading per-rank files in the reference's schema (global tags / dof numbers, tie
    closure and renumbering done by the driver) and from the Python partitioner + binding that the GPU parity runs use."""
import os
import subprocess
import sys

import numpy as np
import pytest

import cases
from svl_b200 import model as M, partition as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "svl_b200", "SeismoVLAB_gpu.exe")


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("shim") / "libsvlgpu_trace.so")
    subprocess.run(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-o", so, os.path.join(ROOT, "tests", "trace_shim", "svlgpu_trace.c")],
                   check=True)
    return so


def run_driver(shim, part, pattern, trace, nparts=1):
    env = dict(os.environ, LD_PRELOAD=shim, SVLGPU_TRACE=trace)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    args = [EXE] + (["-np", str(nparts)] if nparts > 1 else []) + ["-dir", part, "-file", pattern]
    r = subprocess.run(args, capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    return [open(trace + (f".{q}" if nparts > 1 else "")).read().splitlines() for q in range(nparts)]


@pytest.mark.parametrize("case", ["kat444", "kat444_masses", "drm_box", "pml2d", "pml3d", "lysmer_column", "hex8_layered_rayleigh", "j2ps_area"])
def test_driver_hands_the_same_calls_from_json_and_from_binary_tables(shim, tmp_path, case):
    m = cases.CASES[case]()
    part = M.write_reference_json(m, str(tmp_path), "Case", "Run")
    M.pack_partition_tables(os.path.join(part, "Case.1.0.json"))
    a = run_driver(shim, part, "Case.1.$.json", str(tmp_path / "a.trace"))[0]
    b = run_driver(shim, part, "Case.1.$.bin.json", str(tmp_path / "b.trace"))[0]
    assert a == b
    kinds = {l.split()[0] for l in a}
    assert {"create", "set_nodes", "add_material", "add_elements", "add_node_recorder", "finalize"} <= kinds
    if case.startswith("pml"):
        assert sum(l.startswith("add_constraint") for l in a) == len(m.constraints)
    if case == "drm_box":
        assert "add_drm_load" in kinds
    if case == "kat444_masses":
        assert sum(l.startswith("add_nodal_mass") for l in a) == 2
    # the run completed against the stand-in: a recorder file of the right shape exists
    rec = M.read_node_recorder(os.path.join(str(tmp_path), "Solution", "Run", "disp.0.out"))
    assert rec.shape == (m.nt - 1, len(m.rec_dofs()))


_PY_RANK = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, cases
from svl_b200 import capi, partition as P
case, nparts, rank, how = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
capi.LIB_PATH = {shim!r}                       # the recording stand-in instead of svl_b200/libsvlgpu.so (this test only)
m = (cases.CASES.get(case) or cases.REACTION_CASE_FUNCS[case])()
ep = eval(how)
s = P.split_model(m, ep, nparts)[rank]
d = capi.DeviceModel(s, comm=(rank, nparts, b"\x5a" * 128), fields=(0, 3) if case in cases.REACTION_CASE_FUNCS else (0,))
d.close()
"""


def _kv(line):
    tok = line.split()
    return tok[0], dict(t.split("=", 1) for t in tok[1:] if "=" in t)


def _canon(lines):
    """the order-independent content of a trace that both front ends must agree on"""
    out = {"constraints": set(), "rayleigh": set(), "halos": [], "elements": [], "loads": set(), "recorders": [], "options": set()}
    for l in lines:
        name, kv = _kv(l)
        if name in ("create", "set_nodes", "finalize", "comm_init"):
            out[name] = kv
        elif name == "add_material":
            out.setdefault("materials", []).append(kv)
        elif name == "add_elements":
            out["elements"].append(kv)
        elif name == "add_constraint":
            out["constraints"].add(tuple(sorted(kv.items())))
        elif name == "set_rayleigh":
            out["rayleigh"].add(tuple(sorted(kv.items())))
        elif name == "add_halo":
            out["halos"].append(kv)
        elif name == "add_point_load":
            out["loads"].add((kv["nnodes"], kv["nt"], kv["factor"], kv["nodes"], kv["series"]))
        elif name == "add_node_recorder":
            out["recorders"].append((kv["field"], kv["nnodes"], kv["nodes"]))
        elif name == "add_drm_load":
            out["drm"] = kv
        elif name == "set_option" and ("pml_collective" in l or "reaction_collective" in l):
            out["options"].add(l)
        elif name == "add_support_motion":
            out.setdefault("supports", set()).add((kv["node"], kv["dof"], kv["nt"], kv["factor"], kv["series"]))
    return out


@pytest.mark.parametrize("case,nparts,how", [
    ("kat444", 2, "P.block_epart((4, 4, 4), (1, 1, 2))"),
    ("lysmer_column", 4, "P.block_epart((3, 3, 6), (2, 1, 2))"),
    ("hex8_layered_rayleigh", 2, "P.block_epart((4, 3, 6), (1, 1, 2))"),
    ("drm_box", 4, "P.block_epart((6, 6, 5), (2, 2, 1))"),
    ("pml3d", 2, "P.centroid_epart(m, (1, 1, 2))"),
    ("pml2d", 3, "np.random.default_rng(5).integers(0, 3, m.n_elem).astype(np.int32)"),
    # moving supports on every partition that holds the node, REACTION recorders and the collective reaction pass
    ("support_column", 2, "P.block_epart((3, 3, 6), (2, 1, 1))"),
    ("reaction_box", 4, "P.block_epart((4, 3, 6), (2, 2, 1))"),
])
def test_driver_on_reference_rank_files_matches_python_partitioner_at_the_c_abi(shim, tmp_path, case, nparts, how):
    """Several ranks.  Left: write_reference_partitions (global numbering, masters travel with slaves only, as createPartitions
    writes them) -> `SeismoVLAB_gpu.exe -np N` (PlanPartitions: tie closure, local numbering, halos; NCCL id through the
    partition directory; svlgpu_comm_init).  Right: partition.split_model -> capi.DeviceModel, the pair the GPU multi-rank
    parity runs use.  Both must hand every rank the same nodes, numbering, elements, constraints, halo lists, loads,
    recorders, options and communicator arguments."""
    reac = case in cases.REACTION_CASE_FUNCS
    m = (cases.CASES.get(case) or cases.REACTION_CASE_FUNCS[case])()
    ep = eval(how)
    resp = ("disp", "reaction") if reac else ("disp",)
    part = M.write_reference_partitions(m, ep, nparts, str(tmp_path), "Case", "Run", resp=resp)
    left = run_driver(shim, part, "Case.1.$.json", str(tmp_path / "cpp.trace"), nparts)
    assert not [f for f in os.listdir(part) if f.startswith(".svlgpu_nccl_id")]          # rank 0 removed the id file
    # the same rank files with their tables in binary sidecars, written straight from the arrays: identical calls
    M.write_reference_partitions(m, ep, nparts, str(tmp_path), "Case", "Run", resp=resp, binary=True)
    assert run_driver(shim, part, "Case.1.$.bin.json", str(tmp_path / "cppbin.trace"), nparts) == left
    for rank in range(nparts):
        env = dict(os.environ, SVLGPU_TRACE=str(tmp_path / "py.trace"), RANK=str(rank))
        r = subprocess.run([sys.executable, "-c", _PY_RANK.format(root=ROOT, shim=shim), case, str(nparts), str(rank), how],
                           capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        right = open(str(tmp_path / "py.trace") + f".{rank}").read().splitlines()
        a, b = _canon(left[rank]), _canon(right)
        for key in ("create", "set_nodes", "finalize", "comm_init", "materials", "elements", "constraints", "halos", "loads",
                    "options", "supports"):
            assert a.get(key) == b.get(key), (rank, key)
        if reac:
            assert any("reaction_collective" in o for o in a["options"])     # also on a rank that records no reaction itself
            assert {r[0] for r in a["recorders"]} <= {"0", "3"}
        if case == "support_column":
            assert a["supports"]
        # recorders differ by design: the reference's files list a recorded node in EVERY partition that holds it
        # (SeismoVLAB.py:237-246), split_model hands it to its owner only
        assert sum(int(r[1]) for r in a["recorders"]) >= sum(int(r[1]) for r in b["recorders"])
        if case == "hex8_layered_rayleigh":
            assert a["rayleigh"] == b["rayleigh"] and a["rayleigh"]
        if case == "drm_box":
            assert a.get("drm") == b.get("drm") and a.get("drm")            # every rank of this split holds DRM elements
        assert a["comm_init"]["rank"] == str(rank) and a["comm_init"]["nranks"] == str(nparts)
        if case.startswith("pml"):
            assert a["options"] and a["constraints"]


_PY_ONE = r"""
import os, sys
sys.path.insert(0, {root!r})
from svl_b200 import capi, model as M
capi.LIB_PATH = {shim!r}                       # the recording stand-in instead of svl_b200/libsvlgpu.so (this test only)
m = M.read_reference_json(sys.argv[1])
d = capi.DeviceModel(m, options={{"integrator": 1.0}} if m.integrator == "NEWMARK" else None)
d.close()
"""


@pytest.mark.parametrize("case", ["kat444", "kat444_masses", "pml2d", "pml3d", "lysmer_column", "hex8_layered_rayleigh", "j2ps_area",
                                  "drm_box", "drm_area", "support_column", "support_area", "F06", "F11"])
def test_cpp_driver_and_python_binding_hand_the_device_the_same_model(shim, tmp_path, case):
    """One partition file, two front ends: `SeismoVLAB_gpu.exe` (svl_host.cpp UpdateMesh + Initialize) and
    `model.read_reference_json` + `capi.DeviceModel` (what the GPU parity tests drive).  Nodes, numbering, materials, elements,
    constraints, Rayleigh groups, nodal masses, loads, recorder nodes and dt reach the C ABI identically -- also for the
    reference's own pre-processor output (fixtures F06: Rayleigh + dashpots under Newmark, F11: EQUAL ties)."""
    if case in cases.CASES or case in cases.REACTION_CASE_FUNCS:
        m = (cases.CASES.get(case) or cases.REACTION_CASE_FUNCS[case])()
        part = M.write_reference_json(m, str(tmp_path), "Case", "Run")
        jp, pattern = os.path.join(part, "Case.1.0.json"), "Case.1.$.json"
    else:
        import shutil
        src = cases.fixture_dir(case)
        shutil.copytree(src, str(tmp_path / "fx"))
        part = str(tmp_path / "fx" / "Partition")
        name = [f for f in os.listdir(part) if f.endswith(".json")][0]
        jp, pattern = os.path.join(part, name), name.replace(".0.json", ".$.json")
        import json
        J = json.load(open(jp))
        os.makedirs(os.path.join(str(tmp_path / "fx"), "Solution", J["Combinations"][str(J["Simulations"]["combo"])]["attributes"]["folder"]),
                    exist_ok=True)
    env = dict(os.environ, LD_PRELOAD=shim, SVLGPU_TRACE=str(tmp_path / "cpp.trace"))
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    cwd = os.path.dirname(part)                          # the reference resolves relative load files against the run directory
    r = subprocess.run([EXE, "-dir", part, "-file", pattern], capture_output=True, text=True, env=env, timeout=300, cwd=cwd)
    assert r.returncode == 0, r.stdout + r.stderr
    env = dict(os.environ, SVLGPU_TRACE=str(tmp_path / "py.trace"))
    env.pop("RANK", None)
    r = subprocess.run([sys.executable, "-c", _PY_ONE.format(root=ROOT, shim=shim), jp], capture_output=True, text=True, env=env,
                       timeout=300, cwd=cwd)
    assert r.returncode == 0, r.stdout + r.stderr
    a = _canon(open(str(tmp_path / "cpp.trace")).read().splitlines())
    b = _canon(open(str(tmp_path / "py.trace")).read().splitlines())
    for key in ("create", "set_nodes", "finalize", "materials", "elements", "constraints", "rayleigh", "loads", "drm", "supports"):
        assert a.get(key) == b.get(key), key
    if case.startswith("support"):
        assert len(a["supports"]) == len({(n, d) for n, d, _, _ in m.supports})      # Supports{} + SUPPORTMOTION load -> one call per dof
    assert a["recorders"][0] == b["recorders"][0]
    if case.startswith("drm"):
        assert a.get("drm")
    masses = lambda path: sorted(l for l in open(path).read().splitlines() if l.startswith("add_nodal_mass"))     # noqa: E731
    assert masses(str(tmp_path / "cpp.trace")) == masses(str(tmp_path / "py.trace"))


_PY_CASE = r"""
import sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
import cases
from svl_b200 import capi
capi.LIB_PATH = {shim!r}                       # the recording stand-in instead of svl_b200/libsvlgpu.so (this test only)
d = capi.DeviceModel(cases.CASES[sys.argv[1]]())
d.close()
"""


@pytest.mark.parametrize("case", ["drm_box", "drm_area"])
def test_drm_text_files_reach_the_c_abi_bit_for_bit(shim, tmp_path, case):
    """ELEMENTLOAD / GENERALWAVE: the per-node `.drm` text files (Driver.hpp:1689-1721: `nt nFields cond`, then nt rows) written
    from the model's tabulated field and read back by the C++ driver give the same svlgpu_add_drm_load call -- element list,
    node list, interior / exterior flags and every field value -- as the Python binding makes from the arrays."""
    m = cases.CASES[case]()
    part = M.write_reference_json(m, str(tmp_path), "Case", "Run")
    a = run_driver(shim, part, "Case.1.$.json", str(tmp_path / "cpp.trace"))[0]
    env = dict(os.environ, SVLGPU_TRACE=str(tmp_path / "py.trace"))
    env.pop("RANK", None)
    r = subprocess.run([sys.executable, "-c", _PY_CASE.format(root=ROOT, shim=shim), case], capture_output=True, text=True, env=env,
                       timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    b = open(str(tmp_path / "py.trace")).read().splitlines()
    da, db = [l for l in a if l.startswith("add_drm_load")], [l for l in b if l.startswith("add_drm_load")]
    assert len(da) == 1 and da == db


def test_launcher_takes_the_other_ranks_down_when_one_stops(shim, tmp_path):
    """`-np N`: a rank that stops (here: rank 1 cannot read its load file) must not leave the others waiting for it -- in a real run
    they would sit in an NCCL collective until a timeout kills the job.  The launcher returns promptly with a non-zero code."""
    import json
    import time
    m = cases.kat444()
    m.point_loads[0].nodes[:] = 124                                   # a node of the upper half: the load goes to rank 1
    part = M.write_reference_partitions(m, P.block_epart((4, 4, 4), (1, 1, 2)), 2, str(tmp_path), "Case", "Run")
    jp = os.path.join(part, "Case.1.1.json")
    J = json.load(open(jp))
    assert "Loads" in J and "Loads" not in json.load(open(os.path.join(part, "Case.1.0.json")))
    for L in J["Loads"].values():
        L["attributes"]["file"] = os.path.join(part, "missing_series.txt")
    json.dump(J, open(jp, "w"))
    env = dict(os.environ, LD_PRELOAD=shim, SVLGPU_TRACE=str(tmp_path / "t"))
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    t0 = time.time()
    r = subprocess.run([EXE, "-np", "2", "-dir", part, "-file", "Case.1.$.json"], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode != 0 and time.time() - t0 < 30.0
    assert "cannot read load file" in r.stdout


def test_element_recorder_files_have_the_reference_layout(shim, tmp_path):
    """ELEMENT recorders (responses STRAIN / STRESS) of the reference's own fixture F02, through the C++ driver against the
    stand-in: header `<n> <ntotal> <nt>`, one `<tag> <Gauss points> <components>` line per element, then nt - 1 rows
    (Recorder.cpp:106-160, :270-300) -- the layout the fixture's LaTeX/cmpResults.py reads with skiprows=2 -- and the driver asks
    the library to keep Gauss-point data and to leave the lattice fast path."""
    import json
    import shutil
    shutil.copytree(cases.fixture_dir("F02"), str(tmp_path / "fx"))
    part = str(tmp_path / "fx" / "Partition")
    J = json.load(open(os.path.join(part, "Debugging_F02.1.0.json")))
    folder = J["Combinations"][str(J["Simulations"]["combo"])]["attributes"]["folder"]
    os.makedirs(os.path.join(str(tmp_path / "fx"), "Solution", folder))
    env = dict(os.environ, LD_PRELOAD=shim, SVLGPU_TRACE=str(tmp_path / "t"), SVLGPU_ELEMENT_RECORDERS="1")
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    r = subprocess.run([EXE, "-dir", part, "-file", "Debugging_F02.1.$.json"], capture_output=True, text=True, env=env, timeout=120,
                       cwd=str(tmp_path / "fx"))
    assert r.returncode == 0, r.stdout + r.stderr
    nt = int(J["Simulations"]["attributes"]["analysis"]["nt"])
    for fn in ("Stress.0.out", "Strain.0.out"):
        lines = open(os.path.join(str(tmp_path / "fx"), "Solution", folder, fn)).read().splitlines()
        assert lines[0].split() == ["1", str(J["Global"]["ntotal"]), str(nt)] and lines[1].split() == ["1", "4", "3"]
        assert len(lines) == 2 + nt - 1 and all(len(l.split()) == 12 for l in lines[2:])
    opts = [l for l in open(str(tmp_path / "t")).read().splitlines() if l.startswith("set_option")]
    assert "set_option keep_gauss=1" in opts and "set_option lattice_guess=0" in opts
