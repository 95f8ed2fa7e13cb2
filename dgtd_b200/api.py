"""ctypes binding of include/dgtd_b200.h.

Class/method names mirror the reference's host interface so that tests read like the reference's:
  Evolution.SetTime / Evolution.Mult   <-> mfem::TimeDependentOperator (GlobalEvolution.h:19-22)
  Evolution.Step                        <-> mfem::ODESolver::Step (ode.hpp:72; RK4Solver, ode.cpp:109-136)
  Evolution.run                         <-> maxwell::Solver::run (Solver.cpp:483-533)
"""
from __future__ import annotations

import ctypes as C
import os
import re
from dataclasses import dataclass

import numpy as np

BC_NONE, BC_PEC, BC_PMC, BC_SMA = 0, 1, 2, 3
_HERE = os.path.dirname(os.path.abspath(__file__))
lib_path = os.environ.get("DGTD_B200_LIB") or os.path.join(_HERE, "libdgtd_b200.so")   # override: A/B builds of the same ABI
_HEADER = os.path.join(os.path.dirname(_HERE), "include", "dgtd_b200.h")


def _header_symbols():
    txt = open(_HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dgtd_[a-z0-9_]+)\s*\(", txt)))


HEADER_SYMBOLS = _header_symbols()


class DgtdError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"dgtd_b200 error {code}: {msg}")
        self.code = code


class _PlaneWaveC(C.Structure):
    _fields_ = [("enabled", C.c_int), ("spread", C.c_double), ("mean1d", C.c_double), ("freq", C.c_double),
                ("pol", C.c_double * 3), ("dir", C.c_double * 3), ("fieldtype", C.c_int)]


class _OptionsC(C.Structure):
    _fields_ = [("order", C.c_int), ("alpha", C.c_double),
                ("n_bdr", C.c_int), ("bdr_attr", C.POINTER(C.c_int)), ("bdr_cond", C.POINTER(C.c_int)),
                ("n_tfsf", C.c_int), ("tfsf_attr", C.POINTER(C.c_int)),
                ("n_mat", C.c_int), ("mat_attr", C.POINTER(C.c_int)), ("mat_eps_mu_sigma", C.POINTER(C.c_double)),
                ("pw", _PlaneWaveC), ("tfsf_gate", C.c_int), ("device", C.c_int),
                ("rank", C.c_int), ("nranks", C.c_int), ("partitioning", C.POINTER(C.c_int))]


def _load():
    if not os.path.exists(lib_path):
        raise ImportError(f"{lib_path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(dgtd_b200 has no CPU fallback)")
    L = C.CDLL(lib_path, mode=C.RTLD_GLOBAL)
    L.dgtd_last_error.restype = C.c_char_p
    L.dgtd_version.restype = C.c_char_p
    L.dgtd_launch_count.restype = C.c_longlong
    L.dgtd_launch_count.argtypes = [C.c_void_p]
    L.dgtd_kernel_info.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    L.dgtd_halo_mode.argtypes = [C.c_void_p]
    L.dgtd_gather_destroy.restype = None
    L.dgtd_gather_destroy.argtypes = [C.c_void_p]
    L.dgtd_gather_create.argtypes = [C.c_void_p, C.c_longlong, C.POINTER(C.c_longlong), C.POINTER(C.c_void_p), C.POINTER(C.c_longlong)]
    L.dgtd_gather_dofs.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
    L.dgtd_gather_launch.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
    L.dgtd_gather_wait.argtypes = [C.c_void_p, C.c_void_p]
    L.dgtd_mesh_destroy.restype = None
    L.dgtd_mesh_destroy.argtypes = [C.c_void_p]
    L.dgtd_destroy.restype = None
    L.dgtd_destroy.argtypes = [C.c_void_p]
    return L


lib = _load()


def _ck(rc):
    if rc != 0:
        raise DgtdError(rc, lib.dgtd_last_error().decode())


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


@dataclass
class PlaneWave:
    spread: float
    mean1d: float
    pol: tuple
    dir: tuple
    freq: float = 0.0
    fieldtype: int = 0


class Mesh:
    def __init__(self, handle):
        self._h = handle
        d, nv, ne, nbe = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _ck(lib.dgtd_mesh_info(self._h, C.byref(d), C.byref(nv), C.byref(ne), C.byref(nbe)))
        self.dim, self.nv, self.ne, self.nbe = d.value, nv.value, ne.value, nbe.value

    @classmethod
    def from_arrays(cls, dim, verts, elems, elem_attr, bdr, bdr_attr):
        verts = np.ascontiguousarray(verts, np.float64).reshape(-1, 3)
        elems = np.ascontiguousarray(elems, np.int32).reshape(-1, dim + 1)
        ea = np.ascontiguousarray(elem_attr, np.int32)
        bdr = np.ascontiguousarray(bdr, np.int32).reshape(-1, dim)
        ba = np.ascontiguousarray(bdr_attr, np.int32)
        h = C.c_void_p()
        _ck(lib.dgtd_mesh_from_arrays(dim, len(verts), _dp(verts), len(elems), _ip(elems), _ip(ea), len(bdr),
                                      _ip(bdr) if len(bdr) else None, _ip(ba) if len(bdr) else None, C.byref(h)))
        return cls(h)

    @classmethod
    def load(cls, path):
        h = C.c_void_p()
        _ck(lib.dgtd_mesh_load(str(path).encode(), C.byref(h)))
        return cls(h)

    @classmethod
    def cartesian3d(cls, nx, ny=None, nz=None, sx=1.0, sy=1.0, sz=1.0):
        h = C.c_void_p()
        _ck(lib.dgtd_mesh_cartesian3d(nx, ny or nx, nz or nx, C.c_double(sx), C.c_double(sy), C.c_double(sz), C.byref(h)))
        return cls(h)

    def arrays(self):
        v = np.zeros((self.nv, 3)); e = np.zeros((self.ne, self.dim + 1), np.int32); ea = np.zeros(self.ne, np.int32)
        b = np.zeros((self.nbe, self.dim), np.int32); ba = np.zeros(self.nbe, np.int32)
        _ck(lib.dgtd_mesh_get_arrays(self._h, _dp(v), _ip(e), _ip(ea), _ip(b), _ip(ba)))
        return v, e, ea, b, ba

    def boundary_elements(self, attrs):
        """(element, local face) pairs on boundary elements tagged with `attrs` (both sides of interior surfaces)."""
        a = np.ascontiguousarray(list(attrs), np.int32)
        n = C.c_longlong()
        _ck(lib.dgtd_mesh_boundary_elements(self._h, len(a), _ip(a), C.c_longlong(0), None, C.byref(n)))
        pairs = np.zeros((n.value, 2), np.int32)
        _ck(lib.dgtd_mesh_boundary_elements(self._h, len(a), _ip(a), C.c_longlong(n.value), _ip(pairs), C.byref(n)))
        return pairs

    def partition(self, nranks, method="rcb"):
        """element -> rank: "rcb" (coordinate bisection, the default of Evolution) or "metis" (k-way on the dual graph)."""
        p = np.zeros(self.ne, np.int32)
        _ck((lib.dgtd_mesh_partition_metis if method == "metis" else lib.dgtd_mesh_partition)(self._h, nranks, _ip(p)))
        return p

    def __del__(self):
        if getattr(self, "_h", None):
            lib.dgtd_mesh_destroy(self._h)
            self._h = None


def _options(order, alpha, bdr, tfsf, materials, planewave, tfsf_gate, device, rank, nranks, partitioning):
    o = _OptionsC()
    keep = []
    o.order, o.alpha = int(order), float(alpha)
    ba = np.array(list(bdr.keys()), np.int32); bc = np.array(list(bdr.values()), np.int32)
    tf = np.array(list(tfsf), np.int32)
    ma = np.array(list(materials.keys()), np.int32)
    mv = np.array([materials[k] for k in materials], np.float64).reshape(-1, 3)
    keep += [ba, bc, tf, ma, mv]
    o.n_bdr, o.bdr_attr, o.bdr_cond = len(ba), _ip(ba), _ip(bc)
    o.n_tfsf, o.tfsf_attr = len(tf), _ip(tf)
    o.n_mat, o.mat_attr, o.mat_eps_mu_sigma = len(ma), _ip(ma), _dp(mv)
    if planewave is not None:
        o.pw.enabled = 1
        o.pw.spread, o.pw.mean1d, o.pw.freq = planewave.spread, planewave.mean1d, planewave.freq
        o.pw.pol = (C.c_double * 3)(*planewave.pol); o.pw.dir = (C.c_double * 3)(*planewave.dir)
        o.pw.fieldtype = planewave.fieldtype
    o.tfsf_gate, o.device, o.rank, o.nranks = int(tfsf_gate), int(device), int(rank), int(nranks)
    if partitioning is not None:
        pa = np.ascontiguousarray(partitioning, np.int32); keep.append(pa)
        o.partitioning = _ip(pa)
    return o, keep


def setup_query(mesh: Mesh, name: str, dtype, *, order, alpha=1.0, bdr=None, tfsf=(), materials=None, planewave=None,
                tfsf_gate=True, rank=0, nranks=1, partitioning=None):
    """Host-only: one of the flat operator tables a rank would upload (no GPU needed)."""
    o, keep = _options(order, alpha, bdr or {}, tfsf, materials or {}, planewave, tfsf_gate, 0, rank, nranks, partitioning)
    n = C.c_longlong()
    _ck(lib.dgtd_setup_query(mesh._h, C.byref(o), name.encode(), None, C.c_longlong(0), C.byref(n)))
    buf = np.zeros(n.value // np.dtype(dtype).itemsize, dtype)
    _ck(lib.dgtd_setup_query(mesh._h, C.byref(o), name.encode(), buf.ctypes.data_as(C.c_void_p), C.c_longlong(buf.nbytes), C.byref(n)))
    return buf


class Gather:
    """Asynchronous snapshot of a fixed dof list (probes, RCS surface export): launch() queues a gather kernel on the compute
    stream and a device-to-host copy on a side stream, wait() makes `out` readable.  dgtd_gather_* in the C ABI."""

    def __init__(self, ev, dofs):
        dofs = np.ascontiguousarray(dofs, np.int64)
        self._ev, self._h = ev, C.c_void_p()
        n = C.c_longlong()
        _ck(lib.dgtd_gather_create(ev._h, C.c_longlong(len(dofs)), dofs.ctypes.data_as(C.POINTER(C.c_longlong)), C.byref(self._h), C.byref(n)))
        self.n_local = n.value
        self.dofs = np.zeros(self.n_local, np.int64)
        if self.n_local:
            _ck(lib.dgtd_gather_dofs(self._h, self.dofs.ctypes.data_as(C.POINTER(C.c_longlong))))

    def launch(self, out):
        """out: float64 array [6, n_local] (pinned memory makes the copy asynchronous); not readable before wait()."""
        assert out.dtype == np.float64 and out.size == 6 * self.n_local and out.flags["C_CONTIGUOUS"]
        _ck(lib.dgtd_gather_launch(self._ev._h, self._h, _dp(out)))

    def wait(self):
        _ck(lib.dgtd_gather_wait(self._ev._h, self._h))

    def close(self):
        if getattr(self, "_h", None):
            lib.dgtd_gather_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()


class Evolution:
    """The evolution operator + fused RK4 on one GPU (one rank)."""

    def __init__(self, mesh: Mesh, *, order, alpha=1.0, bdr=None, tfsf=(), materials=None, planewave=None,
                 tfsf_gate=True, device=0, rank=0, nranks=1, partitioning=None):
        o, keep = _options(order, alpha, bdr or {}, tfsf, materials or {}, planewave, tfsf_gate, device, rank, nranks, partitioning)
        self._h = C.c_void_p()
        self.mesh = mesh
        _ck(lib.dgtd_create(mesh._h, C.byref(o), C.byref(self._h)))
        ng, np_, nel, nl = C.c_longlong(), C.c_int(), C.c_longlong(), C.c_longlong()
        _ck(lib.dgtd_sizes(self._h, C.byref(ng), C.byref(np_), C.byref(nel), C.byref(nl)))
        self.N, self.Np, self.ne_local, self.n_local = ng.value, np_.value, nel.value, nl.value
        self.rank, self.nranks = rank, nranks
        self._t = 0.0

    # --- mfem::TimeDependentOperator ---
    def Height(self):
        return 6 * self.N

    def SetTime(self, t):
        self._t = float(t)

    def GetTime(self):
        return self._t

    def Mult(self, x, out=None):
        x = np.ascontiguousarray(x, np.float64)
        if x.size != 6 * self.N:
            raise DgtdError(-1, f"Mult: input has {x.size} entries, operator height is {6 * self.N}")
        if out is None:
            out = np.zeros(6 * self.N)
        _ck(lib.dgtd_mult(self._h, C.c_double(self._t), _dp(x), _dp(out), 0))
        return out

    # --- state + mfem::ODESolver ---
    def set_state(self, x):
        x = np.ascontiguousarray(x, np.float64)
        if x.size != 6 * self.N:
            raise DgtdError(-1, "set_state: wrong size")
        _ck(lib.dgtd_set_state(self._h, _dp(x)))

    def get_state(self, out=None):
        if out is None:
            out = np.zeros(6 * self.N)
        _ck(lib.dgtd_get_state(self._h, _dp(out)))
        return out

    def set_state_local(self, x):
        x = np.ascontiguousarray(x, np.float64)
        if x.size != 6 * self.n_local:
            raise DgtdError(-1, "set_state_local: wrong size")
        _ck(lib.dgtd_set_state_local(self._h, _dp(x)))

    def get_state_local(self, out=None):
        if out is None:
            out = np.zeros(6 * self.n_local)
        _ck(lib.dgtd_get_state_local(self._h, _dp(out)))
        return out

    def set_state_parlocal(self, x):
        """local vector [6][n_local], owned elements by ascending global id (the rank's mfem::ParMesh order)"""
        x = np.ascontiguousarray(x, np.float64)
        if x.size != 6 * self.n_local:
            raise DgtdError(-1, "set_state_parlocal: wrong size")
        _ck(lib.dgtd_set_state_parlocal(self._h, _dp(x)))

    def get_state_parlocal(self, out=None):
        if out is None:
            out = np.zeros(6 * self.n_local)
        _ck(lib.dgtd_get_state_parlocal(self._h, _dp(out)))
        return out

    def Mult_parlocal(self, x, out=None):
        x = np.ascontiguousarray(x, np.float64)
        if x.size != 6 * self.n_local:
            raise DgtdError(-1, "Mult_parlocal: wrong size")
        if out is None:
            out = np.zeros(6 * self.n_local)
        _ck(lib.dgtd_mult_parlocal(self._h, C.c_double(self._t), _dp(x), _dp(out)))
        return out

    def Step(self, t, dt):
        _ck(lib.dgtd_rk4_step(self._h, C.c_double(t), C.c_double(dt)))
        return t + dt

    def run(self, t0, dt, nsteps):
        _ck(lib.dgtd_rk4_run(self._h, C.c_double(t0), C.c_double(dt), int(nsteps)))
        return t0 + nsteps * dt

    def run_until(self, t0, dt, t_final, check_every=0):
        """maxwell::Solver::run: steps of min(dt, t_final - t) until t_final -> (t, nsteps, unstable)."""
        t, n, bad = C.c_double(t0), C.c_longlong(), C.c_int()
        _ck(lib.dgtd_run_until(self._h, C.byref(t), C.c_double(dt), C.c_double(t_final), int(check_every), C.byref(n), C.byref(bad)))
        return t.value, n.value, bool(bad.value)

    def synchronize(self):
        _ck(lib.dgtd_synchronize(self._h))

    def set_stream(self, cuda_stream_ptr):
        _ck(lib.dgtd_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def norm2_local(self):
        s = C.c_double()
        _ck(lib.dgtd_norm2_local(self._h, C.byref(s)))
        return s.value

    def node_coords(self):
        xyz = np.zeros((self.N, 3))
        _ck(lib.dgtd_node_coords(self._h, _dp(xyz)))
        return xyz

    def local_elements(self):
        ids = np.zeros(self.ne_local, np.int32)
        _ck(lib.dgtd_local_elements(self._h, _ip(ids)))
        return ids

    def sample(self, local_elem, shape):
        le = np.ascontiguousarray(local_elem, np.int32); sh = np.ascontiguousarray(shape, np.float64)
        out = np.zeros((len(le), 6))
        _ck(lib.dgtd_sample(self._h, len(le), _ip(le), _dp(sh), _dp(out)))
        return out

    def state_device_ptr(self):
        p = C.POINTER(C.c_double)()
        _ck(lib.dgtd_state_device_ptr(self._h, C.byref(p)))
        return C.cast(p, C.c_void_p).value

    def launch_count(self):
        return lib.dgtd_launch_count(self._h)

    def kernel_info(self):
        buf = C.create_string_buffer(320)
        _ck(lib.dgtd_kernel_info(self._h, buf, 320))
        return buf.value.decode()

    def halo_bytes(self):
        b = C.c_longlong()
        _ck(lib.dgtd_halo_bytes(self._h, C.byref(b)))
        return b.value

    def halo_mode(self):
        """0 single rank, 1 NCCL send/recv, 2 peer-memory stores fused into the stage kernel (DGTD_HALO_*)."""
        return lib.dgtd_halo_mode(self._h)

    def comm_init(self, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        _ck(lib.dgtd_comm_init(self._h, buf))

    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        _ck(lib.dgtd_comm_unique_id(buf))
        return buf.raw

    def close(self):
        if getattr(self, "_h", None):
            lib.dgtd_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()
