"""Host-side multi-GPU logic on CPU: the element partitioner and the interface (halo) lists, exercised by two
`gloo` ranks.  Each rank evaluates the ORACLE's internal force / lumped mass on its own sub-model, the ranks
exchange the interface partial sums exactly the way libsvlgpu's NCCL exchange does (per-peer lists, ascending
rank summation), and the result must equal the single-domain oracle."""
import os
import socket

import numpy as np
import pytest

import cases
from svl_b200 import partition as P


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, q):
    import torch
    import torch.distributed as dist
    from oracle_lib import Oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = cases.CASES[case]()
        if case.startswith("pml"):
            # soil box + PML layer with EQUAL ties: geometric split by element centroid, cut through soil and PML alike
            subs = P.split_model(m, P.centroid_epart(m, (world, 1) if m.ndim == 2 else (1, 1, world)), world)
        else:
            ne = {"kat444": (4, 4, 4), "drm_box": (6, 6, 5), "quad4_area": (8, 6), "j2_column": (2, 2, 6)}[case]
            grid = P.proc_grid(world)[3 - len(ne):] if len(ne) == 3 else (1, world)
            subs = P.split_model(m, P.block_epart(ne, grid), world)
        s = subs[rank]
        o = Oracle()
        rng = np.random.default_rng(42)
        Ug = rng.uniform(-1e-3, 1e-3, m.n_total)

        def dofs_of(model, nodes):
            """node-major total dofs of the listed nodes (PML nodes carry 9 / 5 dofs, soil nodes ndim)"""
            return np.concatenate([np.arange(model.node_ptr[n], model.node_ptr[n + 1]) for n in nodes])

        # local restriction of the global state
        gd = dofs_of(m, s.global_nodes)
        out = {}
        for name, vec in (("F", o.internal_force(s, Ug[gd])), ("M", o.mass_diagonal(s))):
            vec = vec.copy()
            part = vec.copy()
            reqs, recv = [], {}
            for peer, nodes in sorted(s.halos.items()):
                dofs = dofs_of(s, nodes)
                recv[peer] = (dofs, torch.zeros(len(dofs), dtype=torch.float64))
                reqs.append(dist.isend(torch.from_numpy(part[dofs].copy()), peer))
                reqs.append(dist.irecv(recv[peer][1], peer))
            for r_ in reqs:
                r_.wait()
            # ascending-rank summation at the interface dofs
            if_dofs = np.unique(np.concatenate([d for d, _ in recv.values()])) if recv else np.array([], int)
            tot = np.zeros(len(if_dofs))
            for rk in range(world):
                if rk == rank:
                    tot += part[if_dofs]
                elif rk in recv:
                    d, t = recv[rk]
                    pos = np.searchsorted(if_dofs, d)
                    tot[pos] += t.numpy()
            vec[if_dofs] = tot
            out[name] = (gd, vec)
        q.put((rank, out, [len(v) for v in s.halos.values()], s.blocks))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", ["kat444", "drm_box", "quad4_area", "pml2d", "pml3d"])
def test_two_rank_interface_sum_matches_single_domain(oracle, case):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, case, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    m = cases.CASES[case]()
    rng = np.random.default_rng(42)
    Ug = rng.uniform(-1e-3, 1e-3, m.n_total)
    Fg, Mg = oracle.internal_force(m, Ug), oracle.mass_diagonal(m)
    for rank, out, halo_sizes, blocks in res:
        assert halo_sizes and halo_sizes[0] > 0 and (len(blocks) == 1 or case.startswith("pml"))
        gd, F = out["F"]
        assert np.abs(F - Fg[gd]).max() <= 1e-12 * np.abs(Fg).max()
        gd, Mv = out["M"]
        assert np.abs(Mv - Mg[gd]).max() <= 1e-12 * np.abs(Mg).max()


def test_split_hands_loads_and_recorders_to_one_partition():
    m = cases.drm_box()
    m.point_loads = cases.kat444().point_loads
    m.point_loads[0].nodes[:] = 7 * 7 * 3 + 10            # a node on the cut plane of the 2-way z split
    subs = P.split_model(m, P.block_epart((6, 6, 5), (1, 1, 2)), 2)
    assert sum(len(s.point_loads) for s in subs) == 1 and len(subs[0].point_loads) == 1
    assert sorted(np.concatenate([s.rec_global for s in subs])) == sorted(m.rec_nodes)
    assert sum(len(s.drm.elems) for s in subs) == len(m.drm.elems)
    # halo lists are mirror images in global numbering
    a = subs[0].global_nodes[subs[0].halos[1]]
    b = subs[1].global_nodes[subs[1].halos[0]]
    assert (a == b).all() and (np.diff(a) > 0).all()


def test_split_with_dashpots_counts_every_dashpot_once():
    """ZeroLength1D dashpots (appended after the lattice elements) follow the lowest rank that holds their soil node; the
    fixed twin nodes travel with them; the summed damping diagonal over the partitions equals the global one."""
    m = cases.lysmer_column()
    n_solid = int((m.elem_kind == 1).sum())
    subs = P.split_model(m, P.block_epart((3, 3, 6), (2, 1, 2)), 4)
    from svl_b200.model import ZEROLENGTH1D
    assert sum(int((s.elem_kind == ZEROLENGTH1D).sum()) for s in subs) == m.n_elem - n_solid
    assert sum(int((s.elem_kind == 1).sum()) for s in subs) == n_solid
    eta_global = np.zeros(m.n_total)
    for e in np.nonzero(m.elem_kind == ZEROLENGTH1D)[0]:
        node = m.elem_conn[e, 1]                                   # soil node (the twin comes first)
        eta_global[m.node_ptr[node] + int(m.elem_attr[e, 0])] += m.materials[m.elem_mat[e]][1][0]
    eta_sum = np.zeros(m.n_total)
    for s in subs:
        for e in np.nonzero(s.elem_kind == ZEROLENGTH1D)[0]:
            tw, node = s.elem_conn[e, 0], s.elem_conn[e, 1]
            assert (np.asarray(s.freedof[tw]) == -1).all()         # the twin stays fully restrained
            g = s.global_nodes[node]
            eta_sum[m.node_ptr[g] + int(s.elem_attr[e, 0])] += s.materials[s.elem_mat[e]][1][0]
    assert np.array_equal(eta_sum, eta_global) and eta_global.sum() > 0


def test_split_of_one_rank_equals_its_entry_in_the_full_split():
    m = cases.pml3d()
    ep = P.centroid_epart(m, (2, 2, 2))
    full = P.split_model(m, ep, 8)
    one = P.split_model(m, ep, 8, ranks=(5,))
    assert [s is None for s in one] == [r != 5 for r in range(8)]
    a, b = full[5], one[5]
    assert np.array_equal(a.global_nodes, b.global_nodes) and np.array_equal(a.elem_conn, b.elem_conn)
    assert a.constraints == b.constraints and sorted(a.halos) == sorted(b.halos)
    assert all(np.array_equal(a.halos[q], b.halos[q]) for q in a.halos)


def test_local_box_matches_split_of_global_box():
    from svl_b200 import model as M
    grid = (2, 2, 2)
    n = (3, 2, 2)
    g = M.make_box_model((6, 4, 4), 1.0, nt=4)
    subs = P.split_model(g, P.block_epart((6, 4, 4), grid), 8)
    for r in range(8):
        lb = P.local_box(n, grid, r, nt=4)
        assert np.allclose(lb.coords, subs[r].coords)
        assert (lb.elem_conn == subs[r].elem_conn).all()
        assert sorted(lb.halos) == sorted(subs[r].halos)
        for peer in lb.halos:
            assert (lb.halos[peer] == subs[r].halos[peer]).all()
        assert (np.asarray(lb.freedof_flat) < 0).sum() == (np.asarray(subs[r].freedof_flat) < 0).sum()


@pytest.mark.parametrize("case,grid", [("pml2d", (2, 1)), ("pml2d", (1, 2)), ("pml3d", (2, 2, 2))])
def test_split_carries_equal_constraints_and_mirrors_pml_halos(case, grid):
    """Soil-PML ties under partitioning: every partition that holds the slave or the master of an EQUAL constraint holds
    both and the constraint itself, renumbered to its own total / free dofs; the halo lists (9- / 5-dof PML nodes included)
    are mirror images in global numbering; every rank of a PML model is told to join the block solve's reductions."""
    m = cases.CASES[case]()
    n = int(np.prod(grid))
    subs = P.split_model(m, P.centroid_epart(m, grid), n)
    fd = np.asarray(m.freedof_flat)
    node_of_total = np.repeat(np.arange(m.n_nodes), np.diff(m.node_ptr))
    total_of_free = {int(f): q for q, f in enumerate(fd) if f > -1}
    glob = {t: (node_of_total[sl], sl - m.node_ptr[node_of_total[sl]], node_of_total[total_of_free[ms[0]]],
                total_of_free[ms[0]] - m.node_ptr[node_of_total[total_of_free[ms[0]]]]) for t, sl, ms, _ in m.constraints}
    seen = set()
    assert sum(s.n_elem for s in subs) == m.n_elem
    for r, s in enumerate(subs):
        assert s.pml_collective
        sfd = np.asarray(s.freedof_flat)
        s_node_of_total = np.repeat(np.arange(s.n_nodes), np.diff(s.node_ptr))
        s_total_of_free = {int(f): q for q, f in enumerate(sfd) if f > -1}
        for t, sl, ms, fac in s.constraints:
            assert sfd[sl] == t and fac == [1.0]
            sn, mq = s_node_of_total[sl], s_total_of_free[ms[0]]
            mn = s_node_of_total[mq]
            assert (s.global_nodes[sn], sl - s.node_ptr[sn], s.global_nodes[mn], mq - s.node_ptr[mn]) == glob[t]
            seen.add(t)
        # no constrained dof without its constraint, no master left behind
        assert set(int(v) for v in sfd[sfd < -1]) == {t for t, *_ in s.constraints}
        for peer, nodes in s.halos.items():
            a = s.global_nodes[nodes]
            b = subs[peer].global_nodes[subs[peer].halos[r]]
            assert (a == b).all() and (np.diff(a) > 0).all()
        assert any((s.node_ndof[nodes] > m.ndim).any() for nodes in s.halos.values())      # PML nodes on the cut
    assert seen == set(glob)


@pytest.mark.parametrize("case,grid", [("pml2d", (2, 2)), ("pml3d", (2, 2, 2)), ("pml3d", (1, 1, 3))])
def test_pml_block_unknowns_shared_between_ranks_are_mirror_images(case, grid):
    """What the multi-rank PML block solve relies on (planner.cu: hp.unk lists, halo.cu: pmlx_setup): walking a halo list
    in its declared order and keeping the dofs that CARRY a block unknown -- free dofs of 9- / 5-dof PML nodes and soil
    dofs that are the master of a tie present in the partition -- gives the same (global node, component) sequence on both
    sides of every pair of ranks; and every unknown is owned (lowest holder) by exactly one rank."""
    m = cases.CASES[case]()
    n = int(np.prod(grid))
    subs = P.split_model(m, P.centroid_epart(m, grid), n)
    want = 9 if m.ndim == 3 else 5

    def carriers(s):
        fd = np.asarray(s.freedof_flat)
        total_of_free = {int(f): q for q, f in enumerate(fd) if f > -1}
        car = np.zeros(s.n_total, dtype=bool)
        for nd_ in np.nonzero(s.node_ndof == want)[0]:
            q = np.arange(s.node_ptr[nd_], s.node_ptr[nd_ + 1])
            car[q[fd[q] > -1]] = True
        for _, _, ms, _ in s.constraints:
            car[total_of_free[ms[0]]] = True
        return car

    car = [carriers(s) for s in subs]
    seqs = {}
    for r, s in enumerate(subs):
        for peer, nodes in s.halos.items():
            seq = []
            for nd_ in nodes:
                for k, q in enumerate(range(s.node_ptr[nd_], s.node_ptr[nd_ + 1])):
                    if car[r][q]:
                        seq.append((int(s.global_nodes[nd_]), k))
            seqs[(r, peer)] = seq
    assert any(len(v) for v in seqs.values())
    for (r, peer), seq in seqs.items():
        assert seq == seqs[(peer, r)]
    # ownership: the global carrier set is covered exactly once by the lowest holders
    owned = {}
    for r, s in enumerate(subs):
        lower = set()
        for peer, nodes in s.halos.items():
            if peer < r:
                lower.update(x for x in seqs[(r, peer)])
        for nd_ in range(s.n_nodes):
            for k, q in enumerate(range(s.node_ptr[nd_], s.node_ptr[nd_ + 1])):
                if car[r][q] and (int(s.global_nodes[nd_]), k) not in lower:
                    key = (int(s.global_nodes[nd_]), k)
                    assert key not in owned, key
                    owned[key] = r
    gfd = np.asarray(m.freedof_flat)
    g_total_of_free = {int(f): q for q, f in enumerate(gfd) if f > -1}
    gcar = np.zeros(m.n_total, dtype=bool)
    for nd_ in np.nonzero(m.node_ndof == want)[0]:
        q = np.arange(m.node_ptr[nd_], m.node_ptr[nd_ + 1])
        gcar[q[gfd[q] > -1]] = True
    for _, _, ms, _ in m.constraints:
        gcar[g_total_of_free[ms[0]]] = True
    node_of_total = np.repeat(np.arange(m.n_nodes), np.diff(m.node_ptr))
    assert set(owned) == {(int(node_of_total[q]), int(q - m.node_ptr[node_of_total[q]])) for q in np.nonzero(gcar)[0]}


def _fnv(tags):
    h = 1469598103934665603
    for t in tags:
        h ^= int(t)
        h = (h * 1099511628211) % (1 << 64)
    return h


@pytest.mark.parametrize("case,nparts,how", [("kat444", 8, "block"), ("lysmer_column", 4, "block"), ("pml3d", 2, "centroid"),
                                             ("pml2d", 3, "random"), ("pml3d", 4, "random")])
def test_host_driver_plans_reference_partition_files_like_the_partitioner(tmp_path, case, nparts, how):
    """The C++ host driver in several-rank mode (`SeismoVLAB_gpu.exe -np N -plan`, no GPU touched): per-rank JSON files in the
    reference's schema (GLOBAL tags / dof numbers, masters travel with slaves only: write_reference_partitions mirrors
    createPartitions, SeismoVLAB.py:300-420) -> tie closure, compact local numbering, mirror-image shared-node lists.  Must
    equal what partition.split_model derives from the global model.  The random partitions put masters and slaves on
    different ranks, so the driver's own closure is exercised."""
    import subprocess
    from svl_b200 import model as M
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "svl_b200", "SeismoVLAB_gpu.exe")
    assert os.path.exists(exe), "host driver not built: python -c 'import __graft_entry__ as g; g.build()'"
    m = cases.CASES[case]()
    if how == "block":
        ne = {"kat444": (4, 4, 4), "lysmer_column": (3, 3, 6)}[case]
        ep = P.block_epart(ne, P.proc_grid(nparts) if nparts == 8 else (2, 1, 2))
    elif how == "centroid":
        ep = P.centroid_epart(m, (1, 1, nparts))
    else:
        ep = np.random.default_rng(5).integers(0, nparts, m.n_elem).astype(np.int32)
    part = M.write_reference_partitions(m, ep, nparts, str(tmp_path))
    out = subprocess.run([exe, "-np", str(nparts), "-plan", "-dir", part, "-file", "Model.1.$.json"],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    got = sorted(l for l in out.stdout.splitlines() if l.startswith("PLAN") and "digest" not in l)
    subs = P.split_model(m, ep, nparts)
    want = []
    for r, s in enumerate(subs):
        want.append(f"PLAN rank {r} nodes {s.n_nodes} ntotal {s.n_total} nfree {s.n_free} constraints {len(s.constraints)} "
                    f"pml_collective {int(s.pml_collective)}")
        for peer, nodes in sorted(s.halos.items()):
            want.append(f"PLAN rank {r} peer {peer} shared {len(nodes)} hash {_fnv(s.global_nodes[nodes] + 1)}")
    assert got == sorted(want)
    if how == "random":
        slave_only = P.split_model(m, ep, nparts, tie_closure="slave")
        assert sum(b.n_nodes - a.n_nodes for a, b in zip(slave_only, subs)) > 0      # the closure had work to do


@pytest.mark.parametrize("case", ["kat444", "pml3d", "pml2d", "lysmer_column", "hex8_layered_rayleigh", "j2ps_area"])
def test_binary_partition_tables_round_trip_and_host_driver_digest(tmp_path, case):
    """the Nodes / Elements / Constraints / Dampings tables of a partition file moved into flat binary sidecars
    (model.pack_partition_tables).  (i) The Python reader rebuilds the same Model from either form; (ii) the C++ host driver's
    UpdateMesh builds the same object graph from either form (`-plan` prints a digest of every table it filled); (iii) the
    same holds rank by rank for a partitioned model, and the partition plan does not change."""
    import subprocess
    from svl_b200 import model as M
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "svl_b200", "SeismoVLAB_gpu.exe")
    m = cases.CASES[case]()
    part = M.write_reference_json(m, str(tmp_path), "Case", "Run")
    jp = os.path.join(part, "Case.1.0.json")
    bp = M.pack_partition_tables(jp)
    assert bp.endswith("Case.1.0.bin.json")
    # the direct writer (arrays -> sidecars, no per-entity dicts) produces the same bytes
    import filecmp
    direct = M.write_reference_json(m, str(tmp_path / "direct"), "Case", "Run", binary=True)
    for suffix in ("nodes.bin", "elems.bin") + (("cons.bin",) if m.constraints else ()):
        assert filecmp.cmp(os.path.join(part, f"Case.1.0.{suffix}"), os.path.join(direct, f"Case.1.0.{suffix}"), shallow=False), suffix
    a, b = M.read_reference_json(jp), M.read_reference_json(bp)
    for k in ("coords", "node_ndof", "elem_conn", "elem_kind", "elem_mat", "elem_attr", "freedof_flat"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    assert a.constraints == b.constraints and a.masses == b.masses and a.dt == b.dt and a.nt == b.nt
    assert (a.elem_am is None) == (b.elem_am is None)
    if a.elem_am is not None:
        assert np.array_equal(a.elem_am, b.elem_am) and np.array_equal(a.elem_ak, b.elem_ak)
    sizes = (os.path.getsize(jp), sum(os.path.getsize(os.path.join(part, f)) for f in os.listdir(part) if ".bin" in f))
    assert sizes[1] < 0.5 * sizes[0]

    def plan(args):
        out = subprocess.run([exe, "-plan", "-dir", part] + args, capture_output=True, text=True, timeout=120)
        assert out.returncode == 0, out.stdout + out.stderr
        return sorted(l for l in out.stdout.splitlines() if l.startswith("PLAN"))

    one_json, one_bin = plan(["-file", "Case.1.$.json"]), plan(["-file", "Case.1.$.bin.json"])
    assert one_json == one_bin and any("digest" in l for l in one_json)
    # partitioned: 2 ranks
    ep = P.centroid_epart(m, (1, 2) if m.ndim == 2 else (1, 1, 2))
    M.write_reference_partitions(m, ep, 2, str(tmp_path), "Split", "Run")
    for r in range(2):
        M.pack_partition_tables(os.path.join(part, f"Split.1.{r}.json"))
    two_json, two_bin = plan(["-np", "2", "-file", "Split.1.$.json"]), plan(["-np", "2", "-file", "Split.1.$.bin.json"])
    assert two_json == two_bin and len([l for l in two_json if "digest" in l]) == 2


@pytest.mark.parametrize("name", ["F11", "J12", "F06"])
def test_metis_mesh_file_equals_the_reference_preprocessors(tmp_path, name):
    """partition.write_metis_graph on the model read back from the fixture's JSON reproduces, byte for byte, the `Graph.out`
    the reference's pre-processor writes for `mpmetis` (Core/Partition.py:87-144; golden kept by make_fixture_inputs.py):
    F11 / J12 carry EQUAL soil-PML ties (the slave takes its master's id), F06 mixes quads with 2-node ZeroLength1D
    dashpots."""
    m = cases.fixture_model(name)
    out = P.write_metis_graph(m, str(tmp_path / "Graph.out"))
    ref = os.path.join(cases.fixture_dir(name), "Graph.out")
    assert open(out).read() == open(ref).read()
    if name != "F06":
        assert len(m.constraints) > 0
        ids = set(int(v) for line in open(out).read().splitlines()[1:] for v in line.split())
        assert len(ids) < m.n_nodes                                 # tied nodes collapsed
    # and the way back: an epart file as mpmetis writes it (one rank per line) drives the splitter
    ep = np.arange(m.n_elem) % 2
    np.savetxt(str(tmp_path / "Graph.out.epart.2"), ep, fmt="%d")
    assert (P.read_epart(str(tmp_path / "Graph.out.epart.2")) == ep).all()


def test_rank_file_generator_tool_output_is_planned_by_the_driver(tmp_path):
    """tools/gen_rank_files.py (BASELINE-like configs as rank files with binary tables) -> `SeismoVLAB_gpu.exe -np N -plan`."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "gen_rank_files.py"), "c3", "--n", "8", "--np", "4", "--out",
                        str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    exe = os.path.join(root, "svl_b200", "SeismoVLAB_gpu.exe")
    p = subprocess.run([exe, "-np", "4", "-plan", "-dir", str(tmp_path / "Partition"), "-file", "C3.1.$.bin.json"],
                       capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    heads = [l for l in p.stdout.splitlines() if " nodes " in l]
    assert len(heads) == 4 and all("pml_collective 1" in l for l in heads)


def test_weighted_bisection_balances_the_pml_shell_with_planar_cuts():
    """partition.weighted_epart: recursive coordinate bisection with PML elements weighted 250x (what the block solve costs on the
    device): every rank gets (nearly) the same number of PML elements, every element exactly one rank, and the cuts are planes
    (all elements of one layer along the cut axis go to the same side) -- the geometric split leaves 37 % more PML on the bottom ranks."""
    import cases as _c
    from svl_b200 import model as M, partition as P
    m = M.make_pml_model((24, 24, 24), 6, 1.0, soil=(M.ELASTIC3DLINEAR, _c.SOIL), nt=10)
    pml = np.isin(m.elem_kind, (3, 4))
    for grid in ((1, 1, 2), (1, 2, 2), (2, 2, 2)):
        n = int(np.prod(grid))
        ep = P.weighted_epart(m, grid)
        assert ep.min() == 0 and ep.max() == n - 1 and len(ep) == m.n_elem
        w = np.array([(pml & (ep == r)).sum() for r in range(n)])
        g = np.array([(pml & (P.centroid_epart(m, grid) == r)).sum() for r in range(n)])
        assert w.max() / w.min() < 1.08 < g.max() / g.min()
        # planar cut along z: the z-slab index of an element is a function of its centroid height alone
        cz = np.round(m.coords[m.elem_conn[:, :8]].mean(axis=1)[:, 2], 6)
        slab = ep // (grid[0] * grid[1])
        for z in np.unique(cz):
            assert len(np.unique(slab[cz == z])) == 1
        subs = P.split_model(m, ep, n)
        assert sum(s.n_elem for s in subs) == m.n_elem
        for r, s in enumerate(subs):
            for p_, nodes in s.halos.items():
                assert (s.global_nodes[nodes] == subs[p_].global_nodes[subs[p_].halos[r]]).all()


def test_shuffled_numbering_is_the_same_model(oracle):
    """model.shuffle_numbering (random node ids and element order: the input of the neighbour-list node classes and of the
    planner's locality renumbering) describes the same physics: the oracle's histories agree to rounding."""
    import cases as _c
    from svl_b200 import model as M
    for name in ("hex8_layered_rayleigh", "drm_box", "quad4_area", "lysmer_column", "kat444_masses"):
        m = _c.CASES[name]()
        s = M.shuffle_numbering(m, 3)
        assert s.n_nodes == m.n_nodes and s.n_elem == m.n_elem and not s.blocks
        assert not np.array_equal(s.elem_conn, m.elem_conn)
        a, _ = oracle.run(m)
        b, _ = oracle.run(s)
        assert _c.rel_err(b, a) < 1e-11, name
