"""ctypes binding of the C ABI (include/svlgpu.h -> svl_b200/libsvlgpu.so).

This is the reference-facing call path used by the parity tests and bench.py: every compute
call goes through the `extern "C"` entry points, exactly what a cgo/JNI/ctypes stub on the
reference side would bind.  There is no fallback: if the CUDA library is
missing or no device is usable, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .model import Model

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsvlgpu.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_bp = C.POINTER(C.c_uint8)

DISP, VEL, ACCEL, REACTION = 0, 1, 2, 3

EXPORTS = [
    "svlgpu_create", "svlgpu_destroy", "svlgpu_last_error", "svlgpu_set_nodes", "svlgpu_add_nodal_mass",
    "svlgpu_add_constraint", "svlgpu_add_material", "svlgpu_add_elements", "svlgpu_set_rayleigh",
    "svlgpu_hint_structured_block", "svlgpu_set_option", "svlgpu_add_point_load", "svlgpu_add_drm_load",
    "svlgpu_add_drm_planewave", "svlgpu_add_node_recorder", "svlgpu_finalize", "svlgpu_set_initial_state",
    "svlgpu_step", "svlgpu_sync", "svlgpu_step_host", "svlgpu_get_state", "svlgpu_internal_force",
    "svlgpu_get_mass_diagonal", "svlgpu_get_gauss", "svlgpu_read_recorder", "svlgpu_recorder_rows",
    "svlgpu_recorder_width", "svlgpu_get_counters", "svlgpu_set_kernel_timing", "svlgpu_kernel_time",
    "svlgpu_device_ptr", "svlgpu_add_halo", "svlgpu_nccl_unique_id", "svlgpu_comm_init", "svlgpu_measure_peaks",
    "svlgpu_add_support_motion",
]


class Counters(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "n_elements", "n_nodes", "n_total_dofs", "n_block_nodes", "n_generic_nodes", "n_generic_elements",
        "n_elem_classes", "n_node_classes", "launches_per_step", "total_launches", "device_bytes")] + \
        [("last_step_ms", C.c_double), ("stencil_ms", C.c_double)] + \
        [(n, C.c_int64) for n in ("n_pml_elements", "n_pml_unknowns", "pml_solves", "pml_iterations", "n_nbr_nodes", "n_nbr_classes")]


_lib = None


def load_library():
    """Loads libsvlgpu.so (fails loudly if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    L.svlgpu_create.restype = C.c_void_p
    L.svlgpu_create.argtypes = [C.c_int, C.c_int]
    L.svlgpu_destroy.argtypes = [C.c_void_p]
    L.svlgpu_last_error.restype = C.c_char_p
    L.svlgpu_set_nodes.argtypes = [C.c_void_p, C.c_int, _ip, _dp, _ip, _ip, C.c_int, C.c_int]
    L.svlgpu_add_nodal_mass.argtypes = [C.c_void_p, C.c_int, _ip, _dp]
    L.svlgpu_add_constraint.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _ip, _dp]
    L.svlgpu_add_material.argtypes = [C.c_void_p, C.c_int, _dp, C.c_int]
    L.svlgpu_add_elements.argtypes = [C.c_void_p, C.c_int, C.c_int, _ip, _ip, _dp, C.c_int]
    L.svlgpu_set_rayleigh.argtypes = [C.c_void_p, C.c_int, _ip, C.c_double, C.c_double]
    L.svlgpu_hint_structured_block.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.svlgpu_add_point_load.argtypes = [C.c_void_p, C.c_int, _ip, C.c_int, _dp, C.c_int, _dp, C.c_double]
    L.svlgpu_add_drm_load.argtypes = [C.c_void_p, C.c_int, _ip, C.c_int, _ip, _bp, C.c_int, _dp, C.c_double]
    L.svlgpu_add_drm_planewave.argtypes = [C.c_void_p, C.c_int, _ip, C.c_int, _ip, _bp, _dp, _dp, _dp,
                                           C.c_double, C.c_double, C.c_double, C.c_double, C.c_double]
    L.svlgpu_add_node_recorder.argtypes = [C.c_void_p, C.c_int, C.c_int, _ip, C.c_int]
    L.svlgpu_finalize.argtypes = [C.c_void_p, C.c_double, C.c_int]
    L.svlgpu_set_initial_state.argtypes = [C.c_void_p, _dp, _dp, _dp]
    L.svlgpu_step.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.svlgpu_sync.argtypes = [C.c_void_p]
    L.svlgpu_step_host.argtypes = [C.c_void_p, C.c_int, _dp, C.c_int, C.c_int, _dp, C.c_int]
    L.svlgpu_get_state.argtypes = [C.c_void_p, C.c_int, _ip, C.c_int, _dp]
    L.svlgpu_internal_force.argtypes = [C.c_void_p, _dp]
    L.svlgpu_get_mass_diagonal.argtypes = [C.c_void_p, _dp]
    L.svlgpu_get_gauss.argtypes = [C.c_void_p, C.c_int, C.c_int, _ip, _dp]
    L.svlgpu_read_recorder.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _dp]
    L.svlgpu_recorder_rows.argtypes = [C.c_void_p, C.c_int]
    L.svlgpu_recorder_width.argtypes = [C.c_void_p, C.c_int]
    L.svlgpu_get_counters.argtypes = [C.c_void_p, C.POINTER(Counters)]
    L.svlgpu_set_kernel_timing.argtypes = [C.c_void_p, C.c_int]
    L.svlgpu_kernel_time.argtypes = [C.c_void_p, C.c_int, _dp, C.POINTER(C.c_int64), C.c_int]
    L.svlgpu_device_ptr.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    L.svlgpu_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
    L.svlgpu_add_halo.argtypes = [C.c_void_p, C.c_int, C.c_int, _ip]
    L.svlgpu_nccl_unique_id.argtypes = [C.c_void_p]
    L.svlgpu_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.svlgpu_measure_peaks.argtypes = [C.c_int, _dp, _dp]
    L.svlgpu_add_support_motion.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _dp, C.c_double]
    _lib = L
    return L


class SvlError(RuntimeError):
    pass


def nccl_unique_id() -> bytes:
    """ncclGetUniqueId through the C ABI (rank 0 calls it, the launcher broadcasts the 128 bytes)."""
    L = load_library()
    buf = C.create_string_buffer(128)
    if L.svlgpu_nccl_unique_id(buf):
        raise SvlError(L.svlgpu_last_error().decode())
    return buf.raw


def measure_peaks(device: int = 0):
    """(FP64 FMA TFLOP/s, streaming-copy GB/s) measured on the device by the library's micro-benchmarks."""
    L = load_library()
    tf, bw = C.c_double(0), C.c_double(0)
    if L.svlgpu_measure_peaks(device, C.byref(tf), C.byref(bw)):
        raise SvlError(L.svlgpu_last_error().decode())
    return tf.value, bw.value


def _d(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _i(a):
    return a.ctypes.data_as(_ip) if a is not None else None


class DeviceModel:
    """One model on one GPU, driven through the C ABI."""

    def __init__(self, m: Model, device: int = 0, max_rows: int | None = None, fields=(DISP,),
                 U0=None, V0=None, A0=None, comm=None, options=None):
        """comm = (rank, nranks, unique_id_bytes) joins the NCCL communicator after finalize; the
        model's `.halos` ({peer: local node indices}, svl_b200.partition) are registered before it."""
        if getattr(m, "newton", None) is not None:
            raise SvlError("the model asks for the NewtonRaphson algorithm: the device path implements Linear only "
                           "(09-Algorithms/01-Linear/Linear.cpp); refusing to run a different algorithm silently")
        self.L = load_library()
        self.m = m
        self.h = self.L.svlgpu_create(m.ndim, int(m.lumped))
        if not self.h:
            raise SvlError(self._err())
        self._keep = []
        A = self._arr
        self._ck(self.L.svlgpu_set_nodes(self.h, m.n_nodes, _i(A(m.node_ndof, np.int32)), _d(A(m.coords, np.float64)),
                                         _i(A(m.totaldof, np.int32)), _i(A(m.freedof_flat, np.int32)),
                                         m.n_total, m.n_free))
        for node, mass in m.masses:
            self._ck(self.L.svlgpu_add_nodal_mass(self.h, 1, _i(A([node], np.int32)), _d(A(mass, np.float64))))
        for tag, slave, masters, factors in m.constraints:
            self._ck(self.L.svlgpu_add_constraint(self.h, tag, slave, len(masters), _i(A(masters, np.int32)),
                                                  _d(A(factors, np.float64))))
        for kind, par in m.materials:
            if self.L.svlgpu_add_material(self.h, kind, _d(A(par, np.float64)), len(par)) < 0:
                raise SvlError(self._err())
        # elements in ascending order, grouped by runs of equal kind
        ek = np.asarray(m.elem_kind)
        start = 0
        while start < len(ek):
            end = start
            while end < len(ek) and ek[end] == ek[start]:
                end += 1
            kind = int(ek[start])
            npe = {1: 8, 2: 4, 3: 8, 4: 4, 5: 2}[kind]
            conn = A(m.elem_conn[start:end, :npe], np.int32)
            nattr = {1: 0, 2: 1, 3: 9, 4: 8, 5: 1}[kind]
            attrs = A(m.elem_attr[start:end, :nattr], np.float64) if nattr else None  # hex8: none
            if self.L.svlgpu_add_elements(self.h, kind, end - start, _i(conn), _i(A(m.elem_mat[start:end], np.int32)),
                                          _d(attrs), nattr) < 0:
                raise SvlError(self._err())
            start = end
        if m.elem_am is not None or m.elem_ak is not None:
            am_e = np.zeros(m.n_elem) if m.elem_am is None else np.asarray(m.elem_am, float)
            ak_e = np.zeros(m.n_elem) if m.elem_ak is None else np.asarray(m.elem_ak, float)
            # one call per distinct (am, ak) pair, ascending; a stiffness-proportional part without a mass-proportional one
            # is a pair of its own (it used to be dropped with am == 0: the library must get to refuse or honour it)
            for am, ak in sorted(set(zip(am_e.tolist(), ak_e.tolist()))):
                if am == 0.0 and ak == 0.0:
                    continue
                idx = A(np.nonzero((am_e == am) & (ak_e == ak))[0], np.int32)
                self._ck(self.L.svlgpu_set_rayleigh(self.h, len(idx), _i(idx), float(am), float(ak)))
        opts = {"lattice_guess": 0.0} if not m.blocks else {}      # a model without lattice hints asks for the generic path
        if getattr(m, "pml_collective", False) and (getattr(m, "halos", None) or comm is not None):
            opts["pml_collective"] = 1.0                            # partition of a PML model: every rank joins the block solve,
            #                                                         also one that shares no node with anybody
        if comm is not None and REACTION in tuple(fields):
            opts["reaction_collective"] = 1.0                       # also on a rank that holds none of the recorded nodes
        opts.update(options or {})
        for k, v in opts.items():
            self._ck(self.L.svlgpu_set_option(self.h, k.encode(), float(v)))
        for (n0, nx, ny, nz) in m.blocks:
            self._ck(self.L.svlgpu_hint_structured_block(self.h, n0, nx, ny, nz))
        for pl in m.point_loads:
            self._ck(self.L.svlgpu_add_point_load(self.h, len(pl.nodes), _i(A(pl.nodes, np.int32)), len(pl.dir),
                                                  _d(A(pl.dir, np.float64)), len(pl.series),
                                                  _d(A(pl.series, np.float64)), float(pl.factor)))
        if m.drm is not None:
            d = m.drm
            ext = A(d.exterior, np.uint8)
            if d.field is not None:
                self._ck(self.L.svlgpu_add_drm_load(self.h, len(d.elems), _i(A(d.elems, np.int32)), len(d.nodes),
                                                    _i(A(d.nodes, np.int32)), ext.ctypes.data_as(_bp),
                                                    d.field.shape[1], _d(A(d.field, np.float64)), float(d.factor)))
            else:
                pw = d.planewave
                self._ck(self.L.svlgpu_add_drm_planewave(
                    self.h, len(d.elems), _i(A(d.elems, np.int32)), len(d.nodes), _i(A(d.nodes, np.int32)),
                    ext.ctypes.data_as(_bp), _d(A(pw["dir"], np.float64)), _d(A(pw["pol"], np.float64)),
                    _d(A(pw["xref"], np.float64)), pw["c"], pw["f0"], pw["t0"], pw["amp"], float(d.factor)))
        for node, dof, series, fac in getattr(m, "supports", None) or []:
            sr = A(series, np.float64)
            self._ck(self.L.svlgpu_add_support_motion(self.h, int(node), int(dof), len(sr), _d(sr), float(fac)))
        self.recorders = {}
        rows = max_rows if max_rows is not None else max(m.nt, 1)
        if m.rec_nodes is not None and len(m.rec_nodes):
            for f in fields:
                r = self.L.svlgpu_add_node_recorder(self.h, f, len(m.rec_nodes), _i(A(m.rec_nodes, np.int32)), rows)
                if r < 0:
                    raise SvlError(self._err())
                self.recorders[f] = r
        if U0 is not None or V0 is not None or A0 is not None:
            self._ck(self.L.svlgpu_set_initial_state(self.h, _d(A(U0, np.float64)) if U0 is not None else None,
                                                     _d(A(V0, np.float64)) if V0 is not None else None,
                                                     _d(A(A0, np.float64)) if A0 is not None else None))
        for peer, nodes in sorted(getattr(m, "halos", {}).items()):
            self._ck(self.L.svlgpu_add_halo(self.h, int(peer), len(nodes), _i(A(nodes, np.int32))))
        self._ck(self.L.svlgpu_finalize(self.h, float(m.dt), device))
        self._keep = []
        if comm is not None:
            rank, nranks, uid = comm
            buf = C.create_string_buffer(bytes(uid), 128)
            self._ck(self.L.svlgpu_comm_init(self.h, buf, int(rank), int(nranks)))

    # ---- helpers -------------------------------------------------------------------
    def _arr(self, x, dt):
        a = np.ascontiguousarray(x, dtype=dt)
        self._keep.append(a)
        return a

    def _err(self):
        return self.L.svlgpu_last_error().decode()

    def _ck(self, rc):
        if rc:
            raise SvlError(self._err())

    def close(self):
        if self.h:
            self.L.svlgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- analysis --------------------------------------------------------------------
    def step(self, k0, k1, sync=True):
        self._ck(self.L.svlgpu_step(self.h, k0, k1, int(sync)))

    def sync(self):
        self._ck(self.L.svlgpu_sync(self.h))

    def step_host(self, k, amplitudes, rec=0, row=None):
        # per-step call: keep the marshalling off the path (one persistent amplitude buffer, the row pointer cached per array)
        n = len(amplitudes)
        buf = getattr(self, "_amp_buf", None)
        if buf is None or len(buf) < max(n, 1):
            buf = self._amp_buf = np.zeros(max(n, 1))
            self._amp_ptr = buf.ctypes.data_as(_dp)
        if n:
            buf[:n] = amplitudes
        if row is None:
            rp, rl = None, 0
        else:
            cached = getattr(self, "_row_cache", None)
            if cached is None or cached[0] is not row:
                assert row.dtype == np.float64 and row.flags.c_contiguous
                cached = self._row_cache = (row, row.ctypes.data_as(_dp), len(row))
            rp, rl = cached[1], cached[2]
        if self.L.svlgpu_step_host(self.h, k, self._amp_ptr, n, rec, rp, rl):
            raise SvlError(self._err())

    def run(self, nt=None):
        nt = nt or self.m.nt
        self.step(1, nt, True)
        return {f: self.read_recorder(r) for f, r in self.recorders.items()}

    def read_recorder(self, rec=0):
        rows = self.L.svlgpu_recorder_rows(self.h, rec)
        w = self.L.svlgpu_recorder_width(self.h, rec)
        out = np.zeros((rows, w))
        if rows:
            self._ck(self.L.svlgpu_read_recorder(self.h, rec, 0, rows, _d(out)))
        return out

    def get_state(self, field=DISP, dofs=None):
        if dofs is None:
            out = np.zeros(self.m.n_total)
            self._ck(self.L.svlgpu_get_state(self.h, field, None, self.m.n_total, _d(out)))
        else:
            dofs = np.ascontiguousarray(dofs, np.int32)
            out = np.zeros(len(dofs))
            self._ck(self.L.svlgpu_get_state(self.h, field, _i(dofs), len(dofs), _d(out)))
        return out

    def internal_force(self):
        F = np.zeros(self.m.n_total)
        self._ck(self.L.svlgpu_internal_force(self.h, _d(F)))
        return F

    def mass_diagonal(self):
        M = np.zeros(self.m.n_total)
        self._ck(self.L.svlgpu_get_mass_diagonal(self.h, _d(M)))
        return M

    def gauss(self, field, elems):
        elems = np.ascontiguousarray(elems, np.int32)
        ng, nc = (8, 6) if self.m.ndim == 3 else (4, 3)
        out = np.zeros((len(elems), ng, nc))
        self._ck(self.L.svlgpu_get_gauss(self.h, field, len(elems), _i(elems), _d(out)))
        return out

    def counters(self):
        c = Counters()
        self._ck(self.L.svlgpu_get_counters(self.h, C.byref(c)))
        return {n: getattr(c, n) for n, _ in Counters._fields_}

    def set_kernel_timing(self, on=True):
        self._ck(self.L.svlgpu_set_kernel_timing(self.h, int(on)))

    def kernel_time(self, which=0, reset=False):
        ms = C.c_double(0)
        n = C.c_int64(0)
        self._ck(self.L.svlgpu_kernel_time(self.h, which, C.byref(ms), C.byref(n), int(reset)))
        return ms.value, n.value
