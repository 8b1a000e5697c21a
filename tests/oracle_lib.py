"""ctypes bindings of the CPU oracle (oracle/libsvl_oracle.so) and, when it exists in this
container, of the reference probe (oracle/_ref/libsvlref_probe.so).  TEST INFRASTRUCTURE."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_bp = C.POINTER(C.c_uint8)


def _d(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _i(a):
    return a.ctypes.data_as(_ip) if a is not None else None


def build_oracle():
    so = os.path.join(ORACLE_DIR, "libsvl_oracle.so")
    src = os.path.join(ORACLE_DIR, "svl_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "oracle"], stdout=subprocess.DEVNULL)
    return so


class SvloModel(C.Structure):
    _fields_ = [
        ("ndim", C.c_int32), ("lumped", C.c_int32),
        ("n_nodes", C.c_int32), ("n_total", C.c_int32), ("n_free", C.c_int32),
        ("node_ndof", _ip), ("node_ptr", _ip), ("totaldof", _ip), ("freedof", _ip), ("coords", _dp),
        ("n_cons", C.c_int32), ("cons_tag", _ip), ("cons_slave", _ip), ("cons_ptr", _ip),
        ("cons_master", _ip), ("cons_factor", _dp),
        ("n_mass", C.c_int32), ("mass_node", _ip), ("mass_val", _dp),
        ("n_mat", C.c_int32), ("mat_kind", _ip), ("mat_par", _dp),
        ("n_elem", C.c_int32), ("elem_kind", _ip), ("elem_conn", _ip), ("elem_mat", _ip),
        ("elem_attr", _dp), ("elem_am", _dp), ("elem_ak", _dp),
        ("n_pload", C.c_int32), ("pl_ptr", _ip), ("pl_nodes", _ip), ("pl_dir", _dp), ("pl_nt", _ip),
        ("pl_sptr", _ip), ("pl_series", _dp), ("pl_factor", _dp),
        ("n_drm_elem", C.c_int32), ("n_drm_node", C.c_int32), ("drm_nt", C.c_int32),
        ("drm_elem", _ip), ("drm_node", _ip), ("drm_ext", _bp), ("drm_field", _dp),
        ("drm_factor", C.c_double),
        ("dt", C.c_double), ("ftol", C.c_double), ("mtol", C.c_double),
        ("U0", _dp), ("V0", _dp), ("A0", _dp),
        ("n_sup", C.c_int32), ("sup_dof", _ip), ("sup_ptr", _ip), ("sup_series", _dp), ("sup_factor", _dp),
    ]


class Oracle:
    def __init__(self):
        self.lib = C.CDLL(build_oracle())
        L = self.lib
        L.svlo_run_central_difference.restype = C.c_int
        L.svlo_run_central_difference.argtypes = [C.POINTER(SvloModel), C.c_int, C.c_int, C.c_int, _ip, _dp,
                                                  _dp, C.c_int]
        L.svlo_run_newmark.restype = C.c_int
        L.svlo_run_newmark.argtypes = L.svlo_run_central_difference.argtypes
        L.svlo_run_newmark_newton.restype = C.c_int
        L.svlo_run_newmark_newton.argtypes = [C.POINTER(SvloModel), C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _ip,
                                              _dp, _dp, C.c_int]
        L.svlo_run_extended_newmark.restype = C.c_int
        L.svlo_run_extended_newmark.argtypes = L.svlo_run_central_difference.argtypes
        L.svlo_internal_force.argtypes = [C.POINTER(SvloModel), _dp, _dp]
        L.svlo_mass_diagonal.argtypes = [C.POINTER(SvloModel), _dp]
        L.svlo_elastic3d_C.argtypes = [C.c_double, C.c_double, _dp]
        L.svlo_planestrain_C.argtypes = [C.c_double, C.c_double, _dp]
        L.svlo_hex8_strain.argtypes = [_dp, _dp, _dp]
        L.svlo_hex8_force.argtypes = [_dp, _dp, _dp]
        L.svlo_hex8_mass.argtypes = [_dp, C.c_double, C.c_int, _dp]
        L.svlo_hex8_stiffness.argtypes = [_dp, _dp, _dp]
        L.svlo_quad4_strain.argtypes = [_dp, _dp, _dp]
        L.svlo_quad4_force.argtypes = [_dp, C.c_double, _dp, _dp]
        L.svlo_quad4_mass.argtypes = [_dp, C.c_double, C.c_double, C.c_int, _dp]
        L.svlo_quad4_stiffness.argtypes = [_dp, C.c_double, _dp, _dp]
        L.svlo_j2_update.argtypes = [_dp, _dp, _dp, _dp]
        L.svlo_pml3d_matrices.argtypes = [_dp, C.c_double, C.c_double, C.c_double, _dp, _dp, _dp, _dp, _dp]
        L.svlo_pml2d_matrices.argtypes = [_dp, C.c_double, C.c_double, C.c_double, _dp, _dp, _dp, _dp]
        L.svlo_hex8_drm_force.argtypes = [_dp, _dp, C.c_double, C.c_int, _bp, _dp, _dp, _dp, _dp]
        L.svlo_quad4_drm_force.argtypes = [_dp, C.c_double, _dp, C.c_double, C.c_int, _bp, _dp, _dp, _dp, _dp]

    # ---- element level ----------------------------------------------------------
    def hex8_elastic_force(self, X, U, E, nu):
        X = np.ascontiguousarray(X, float); U = np.ascontiguousarray(U, float)
        Cm = np.zeros(36); eps = np.zeros((8, 6)); f = np.zeros(24)
        self.lib.svlo_elastic3d_C(E, nu, _d(Cm))
        self.lib.svlo_hex8_strain(_d(X), _d(U), _d(eps))
        sig = np.ascontiguousarray(eps @ Cm.reshape(6, 6).T)
        self.lib.svlo_hex8_force(_d(X), _d(sig), _d(f))
        return f

    def quad4_elastic_force(self, X, U, th, E, nu):
        X = np.ascontiguousarray(X, float); U = np.ascontiguousarray(U, float)
        Cm = np.zeros(9); eps = np.zeros((4, 3)); f = np.zeros(8)
        self.lib.svlo_planestrain_C(E, nu, _d(Cm))
        self.lib.svlo_quad4_strain(_d(X), _d(U), _d(eps))
        sig = np.ascontiguousarray(eps @ Cm.reshape(3, 3).T)
        self.lib.svlo_quad4_force(_d(X), th, _d(sig), _d(f))
        return f

    def hex8_mass(self, X, rho, lumped):
        X = np.ascontiguousarray(X, float); M = np.zeros((24, 24))
        self.lib.svlo_hex8_mass(_d(X), rho, int(lumped), _d(M))
        return M

    def hex8_stiffness(self, X, E, nu):
        X = np.ascontiguousarray(X, float); Cm = np.zeros(36); K = np.zeros((24, 24))
        self.lib.svlo_elastic3d_C(E, nu, _d(Cm))
        self.lib.svlo_hex8_stiffness(_d(X), _d(Cm), _d(K))
        return K

    def quad4_mass(self, X, th, rho, lumped):
        X = np.ascontiguousarray(X, float); M = np.zeros((8, 8))
        self.lib.svlo_quad4_mass(_d(X), th, rho, int(lumped), _d(M))
        return M

    def quad4_stiffness(self, X, th, E, nu):
        X = np.ascontiguousarray(X, float); Cm = np.zeros(9); K = np.zeros((8, 8))
        self.lib.svlo_planestrain_C(E, nu, _d(Cm))
        self.lib.svlo_quad4_stiffness(_d(X), th, _d(Cm), _d(K))
        return K

    def j2_path(self, par, eps_path):
        par = np.ascontiguousarray(par, float)
        st = np.zeros(13); out = []
        for e in np.asarray(eps_path, float):
            e = np.ascontiguousarray(e); s = np.zeros(6)
            self.lib.svlo_j2_update(_d(par), _d(e), _d(st), _d(s))
            out.append(s)
        return np.array(out), st

    def pml3d(self, X, E, nu, rho, par):
        X = np.ascontiguousarray(X, float); par = np.ascontiguousarray(par, float)
        M, Cc, K, G = (np.zeros((72, 72)) for _ in range(4))
        self.lib.svlo_pml3d_matrices(_d(X), E, nu, rho, _d(par), _d(M), _d(Cc), _d(K), _d(G))
        return M, Cc, K, G

    def pml2d(self, X, E, nu, rho, par):
        X = np.ascontiguousarray(X, float); par = np.ascontiguousarray(par, float)
        M, Cc, K = (np.zeros((20, 20)) for _ in range(3))
        self.lib.svlo_pml2d_matrices(_d(X), E, nu, rho, _d(par), _d(M), _d(Cc), _d(K))
        return M, Cc, K

    def hex8_drm(self, X, E, nu, rho, lumped, ext, field):
        """field [8, 9] rows as in the .drm file (not yet sign-flipped)."""
        X = np.ascontiguousarray(X, float); Cm = np.zeros(36); f = np.zeros(24)
        self.lib.svlo_elastic3d_C(E, nu, _d(Cm))
        ext = np.ascontiguousarray(ext, np.uint8)
        sg = np.where(ext[:, None] != 0, -1.0, 1.0)
        fl = np.asarray(field, float) * sg
        Uo, Vo, Ao = (np.ascontiguousarray(fl[:, 3 * i:3 * i + 3].reshape(-1)) for i in range(3))
        self.lib.svlo_hex8_drm_force(_d(X), _d(Cm), rho, int(lumped), ext.ctypes.data_as(_bp), _d(Uo), _d(Vo),
                                     _d(Ao), _d(f))
        return f

    def quad4_drm(self, X, th, E, nu, rho, lumped, ext, field):
        X = np.ascontiguousarray(X, float); Cm = np.zeros(9); f = np.zeros(8)
        self.lib.svlo_planestrain_C(E, nu, _d(Cm))
        ext = np.ascontiguousarray(ext, np.uint8)
        sg = np.where(ext[:, None] != 0, -1.0, 1.0)
        fl = np.asarray(field, float) * sg
        Uo, Vo, Ao = (np.ascontiguousarray(fl[:, 2 * i:2 * i + 2].reshape(-1)) for i in range(3))
        self.lib.svlo_quad4_drm_force(_d(X), th, _d(Cm), rho, int(lumped), ext.ctypes.data_as(_bp), _d(Uo),
                                      _d(Vo), _d(Ao), _d(f))
        return f

    # ---- analysis level ------------------------------------------------------------
    def pack(self, m, U0=None, V0=None):
        """svl_b200.model.Model -> (SvloModel, keepalive list)."""
        keep = []

        def A(x, dt):
            a = np.ascontiguousarray(x, dtype=dt)
            keep.append(a)
            return a

        s = SvloModel()
        s.ndim, s.lumped = m.ndim, int(m.lumped)
        s.n_nodes, s.n_total, s.n_free = m.n_nodes, m.n_total, m.n_free
        s.node_ndof = _i(A(m.node_ndof, np.int32)); s.node_ptr = _i(A(m.node_ptr, np.int32))
        s.totaldof = _i(A(m.totaldof, np.int32)); s.freedof = _i(A(m.freedof_flat, np.int32))
        s.coords = _d(A(m.coords, np.float64))
        nc = len(m.constraints)
        s.n_cons = nc
        if nc:
            ptr = np.zeros(nc + 1, np.int32)
            ptr[1:] = np.cumsum([len(c[2]) for c in m.constraints])
            s.cons_tag = _i(A([c[0] for c in m.constraints], np.int32))
            s.cons_slave = _i(A([c[1] for c in m.constraints], np.int32))
            s.cons_ptr = _i(A(ptr, np.int32))
            s.cons_master = _i(A(np.concatenate([c[2] for c in m.constraints]), np.int32))
            s.cons_factor = _d(A(np.concatenate([c[3] for c in m.constraints]), np.float64))
        s.n_mass = len(m.masses)
        if m.masses:
            s.mass_node = _i(A([q[0] for q in m.masses], np.int32))
            s.mass_val = _d(A(np.concatenate([q[1] for q in m.masses]), np.float64))
        s.n_mat = len(m.materials)
        mp = np.zeros((len(m.materials), 8))
        for i, (_, par) in enumerate(m.materials):
            mp[i, :len(par)] = par
        s.mat_kind = _i(A([k for k, _ in m.materials], np.int32)); s.mat_par = _d(A(mp, np.float64))
        s.n_elem = m.n_elem
        s.elem_kind = _i(A(m.elem_kind, np.int32)); s.elem_conn = _i(A(m.elem_conn, np.int32))
        s.elem_mat = _i(A(m.elem_mat, np.int32)); s.elem_attr = _d(A(m.elem_attr if m.elem_attr is not None else np.zeros((m.n_elem, 10)), np.float64))
        if m.elem_am is not None:
            s.elem_am = _d(A(m.elem_am, np.float64)); s.elem_ak = _d(A(m.elem_ak, np.float64))
        npl = len(m.point_loads)
        s.n_pload = npl
        if npl:
            ptr = np.zeros(npl + 1, np.int32); sptr = np.zeros(npl + 1, np.int32)
            ptr[1:] = np.cumsum([len(p.nodes) for p in m.point_loads])
            sptr[1:] = np.cumsum([len(p.series) for p in m.point_loads])
            dirs = np.zeros((npl, 3))
            for i, p in enumerate(m.point_loads):
                dirs[i, :len(p.dir)] = p.dir
            s.pl_ptr = _i(A(ptr, np.int32)); s.pl_sptr = _i(A(sptr, np.int32))
            s.pl_nodes = _i(A(np.concatenate([p.nodes for p in m.point_loads]), np.int32))
            s.pl_dir = _d(A(dirs, np.float64))
            s.pl_nt = _i(A([len(p.series) for p in m.point_loads], np.int32))
            s.pl_series = _d(A(np.concatenate([p.series for p in m.point_loads]), np.float64))
            s.pl_factor = _d(A([p.factor for p in m.point_loads], np.float64))
        if m.drm is not None and m.drm.field is not None:
            d = m.drm
            s.n_drm_elem, s.n_drm_node, s.drm_nt = len(d.elems), len(d.nodes), d.field.shape[1]
            s.drm_elem = _i(A(d.elems, np.int32)); s.drm_node = _i(A(d.nodes, np.int32))
            e = A(d.exterior, np.uint8)
            s.drm_ext = e.ctypes.data_as(_bp)
            s.drm_field = _d(A(d.field, np.float64)); s.drm_factor = d.factor
        sup = getattr(m, "supports", None) or []
        s.n_sup = len(sup)
        if sup:                                 # (node, dof, series, factor): Model.supports
            ptr = np.zeros(len(sup) + 1, np.int32)
            ptr[1:] = np.cumsum([len(q[2]) for q in sup])
            s.sup_dof = _i(A([m.totaldof[m.node_ptr[q[0]] + q[1]] for q in sup], np.int32))
            s.sup_ptr = _i(A(ptr, np.int32))
            s.sup_series = _d(A(np.concatenate([np.asarray(q[2], float) for q in sup]), np.float64))
            s.sup_factor = _d(A([q[3] for q in sup], np.float64))
        s.dt, s.ftol, s.mtol = m.dt, 1e-12, 1e-12
        if U0 is not None:
            s.U0 = _d(A(U0, np.float64))
        if V0 is not None:
            s.V0 = _d(A(V0, np.float64))
        return s, keep

    def run(self, m, nt=None, field=0, rec_dofs=None, nthreads=1, U0=None, integrator="CENTRALDIFFERENCE", newton=None, V0=None):
        """newton = (cnvgtol, nstep, cnvgtest) selects NewtonRaphson (NewmarkBeta only); default Linear.
        Note: like the reference, the first step's internal force comes from the stresses stored by the
        previous UpdateState, i.e. it is ZERO whatever U0 is -- the device evaluates F_int(U0).  Parity runs that want every
        dof excited from the start therefore use an initial VELOCITY field (U0 = 0 keeps both sides identical)."""
        s, keep = self.pack(m, U0, V0)
        nt = nt or m.nt
        rd = np.ascontiguousarray(m.rec_dofs() if rec_dofs is None else rec_dofs, np.int32)
        out = np.zeros((nt - 1, len(rd)))
        Uf = np.zeros(m.n_total)
        if newton is not None:
            assert integrator.upper() == "NEWMARK"
            rc = self.lib.svlo_run_newmark_newton(C.byref(s), float(newton[0]), int(newton[1]), int(newton[2]), nt, field,
                                                  len(rd), _i(rd), _d(out), _d(Uf), nthreads)
        else:
            fn = {"NEWMARK": self.lib.svlo_run_newmark, "EXTENDEDNEWMARK": self.lib.svlo_run_extended_newmark}.get(
                integrator.upper(), self.lib.svlo_run_central_difference)
            rc = fn(C.byref(s), nt, field, len(rd), _i(rd), _d(out), _d(Uf), nthreads)
        if rc:
            raise RuntimeError(f"oracle stop code {rc}")
        return out, Uf

    def internal_force(self, m, U):
        s, keep = self.pack(m)
        U = np.ascontiguousarray(U, float); F = np.zeros(m.n_total)
        self.lib.svlo_internal_force(C.byref(s), _d(U), _d(F))
        return F

    def mass_diagonal(self, m):
        s, keep = self.pack(m)
        Md = np.zeros(m.n_total)
        self.lib.svlo_mass_diagonal(C.byref(s), _d(Md))
        return Md


class RefProbe:
    """The reference's own classes (only available in the build container)."""
    PATH = os.path.join(ORACLE_DIR, "_ref", "libsvlref_probe.so")

    @classmethod
    def available(cls):
        return os.path.exists(cls.PATH)

    def __init__(self):
        self.lib = C.CDLL(self.PATH)
        L = self.lib
        L.refprobe_internal_force.argtypes = [C.c_int, _dp, _dp, C.c_int, _dp, _dp, _dp]
        L.refprobe_matrices.argtypes = [C.c_int, _dp, C.c_int, _dp, _dp, C.c_int, _dp, _dp, _dp, _dp]
        L.refprobe_material_path.argtypes = [C.c_int, _dp, C.c_int, C.c_int, _dp, _dp]
        L.refprobe_drm_force.argtypes = [C.c_int, _dp, C.c_int, _dp, _dp, C.c_int, _bp, _dp, _dp]

    def internal_force(self, kind, X, U, matkind, mp, attr=None):
        X = np.ascontiguousarray(X, float); U = np.ascontiguousarray(U, float)
        mp = np.ascontiguousarray(mp, float)
        attr = np.ascontiguousarray(attr if attr is not None else np.zeros(10), float)
        f = np.zeros(72)
        n = self.lib.refprobe_internal_force(kind, _d(X), _d(U), matkind, _d(mp), _d(attr), _d(f))
        return f[:n]

    def matrices(self, kind, X, matkind, mp, attr=None, lumped=True, want="MCK"):
        X = np.ascontiguousarray(X, float); mp = np.ascontiguousarray(mp, float)
        attr = np.ascontiguousarray(attr if attr is not None else np.zeros(10), float)
        nd = {1: 24, 2: 8, 3: 72, 4: 20}[kind]
        mats = {k: np.zeros((nd, nd)) for k in want}
        self.lib.refprobe_matrices(kind, _d(X), matkind, _d(mp), _d(attr), int(lumped),
                                   _d(mats.get("M")), _d(mats.get("C")), _d(mats.get("K")), _d(mats.get("G")))
        return mats

    def material_path(self, matkind, mp, eps_path):
        mp = np.ascontiguousarray(mp, float)
        eps = np.ascontiguousarray(eps_path, float)
        sig = np.zeros_like(eps)
        self.lib.refprobe_material_path(matkind, _d(mp), eps.shape[0], eps.shape[1], _d(eps), _d(sig))
        return sig

    def drm_force(self, kind, X, matkind, mp, ext, field, attr=None, lumped=True):
        X = np.ascontiguousarray(X, float); mp = np.ascontiguousarray(mp, float)
        attr = np.ascontiguousarray(attr if attr is not None else np.zeros(10), float)
        ext = np.ascontiguousarray(ext, np.uint8); field = np.ascontiguousarray(field, float)
        f = np.zeros(24)
        n = self.lib.refprobe_drm_force(kind, _d(X), matkind, _d(mp), _d(attr), int(lumped),
                                        ext.ctypes.data_as(_bp), _d(field), _d(f))
        return f[:n]
