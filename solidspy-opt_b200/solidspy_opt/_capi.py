"""ctypes binding of libtopo_b200.so (C ABI declared in include/topo_b200.h).

There is no CPU fallback: if the library is missing or no CUDA device is usable, every compute
call raises ``TopoError``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TOPO_B200_LIB", os.path.join(_HERE, "..", "lib", "libtopo_b200.so"))

OK, EINVAL, ENODEV, ECUDA, ENOCONV, ESTATE = 0, -1, -2, -3, -4, -5
PRECOND_JACOBI, PRECOND_GMG = 0, 1
F_U, F_LOAD, F_RHO, F_DC, F_SCALE, F_CONS, F_SENS, F_KEEP, F_RESID, F_SENS_PREV, F_NODAL = range(11)

# every symbol include/topo_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "topo_abi_version", "topo_last_error", "topo_device_count", "topo_create", "topo_destroy",
    "topo_set_cons", "topo_set_loads", "topo_set", "topo_get", "topo_fill", "topo_matvec",
    "topo_matvec_device", "topo_solve", "topo_compliance", "topo_equilibrium_check", "topo_simp_scale",
    "topo_simp_sensitivity", "topo_filter_sensitivity", "topo_oc_update", "topo_simp_step",
    "topo_simp_step_host", "topo_energy_sensitivity", "topo_beso_filter", "topo_beso_nodes", "topo_beso_elements", "topo_beso_elements_radius", "topo_floating_elements", "topo_rank_select",
    "topo_beso_apply", "topo_eso_apply", "topo_beso_save_history", "topo_vonmises_sensitivity", "topo_strain_nodes",
    "topo_get_timing", "topo_gmg_info", "topo_set_gmg", "topo_timer_start", "topo_timer_stop",
    "topo_profile_matvec", "topo_profile_read", "topo_profile_report", "topo_set_stream", "topo_matvec_dev",
    "topo_diag_dev", "topo_vec_axpby_dev", "topo_vec_mul_dev", "topo_vec_dot_dev",
    "topo_dist_init", "topo_dist_info", "topo_dist_xchg_ptrs", "topo_dist_xchg_finish", "topo_dist_gather_ptrs",
    "topo_dist_scalar_ptr", "topo_dist_read_scalars", "topo_dist_setup", "topo_dist_solve_begin",
    "topo_dist_pcg_matvec", "topo_dist_pcg_update", "topo_dist_vcycle_down", "topo_dist_vcycle_tail",
    "topo_dist_vcycle_up", "topo_dist_pcg_smooth", "topo_dist_pcg_finish", "topo_dist_solve_poll",
    "topo_dist_solve_end", "topo_dist_compliance_partial", "topo_dist_filter_pack", "topo_dist_oc",
    "topo_dist_oc_state", "topo_dist_p2p_export", "topo_dist_p2p_attach", "topo_dist_p2p_xchg", "topo_dist_p2p_gather",
    "topo_dist_p2p_allreduce", "topo_dist_p2p_error", "topo_dist_p2p_fused", "topo_dist_p2p_gather_tail", "topo_dist_p2p_mode",
]

# row-slab mode: exchange kinds, gather kinds, scalar blocks (include/topo_b200.h)
X_AP, X_R0, X_Z, X_LEVEL_B, X_LEVEL_X2, X_DIAG, X_FILTER = range(7)
G_CELLS, G_BK = range(2)
S_PAP, S_RR_RZ, S_BB, S_RED = range(4)


class TopoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libtopo_b200 error %d: %s" % (code, msg))
        self.code = code


class SimpParams(C.Structure):
    _fields_ = [("penal", C.c_double), ("emin", C.c_double), ("emax", C.c_double), ("rmin", C.c_double),
                ("dx", C.c_double), ("dy", C.c_double), ("move", C.c_double), ("rtol", C.c_double),
                ("precond", C.c_int), ("maxit", C.c_int), ("warm", C.c_int)]


class SimpResult(C.Structure):
    _fields_ = [("compliance", C.c_double), ("objective", C.c_double), ("g", C.c_double),
                ("change", C.c_double), ("relres", C.c_double), ("cg_iters", C.c_int), ("oc_steps", C.c_int),
                ("equilibrium_ok", C.c_int), ("status", C.c_int)]


class Timing(C.Structure):
    _fields_ = [("solve_ms", C.c_float), ("matvec_ms", C.c_float), ("sens_ms", C.c_float),
                ("filter_ms", C.c_float), ("oc_ms", C.c_float), ("setup_ms", C.c_float),
                ("total_ms", C.c_float), ("kernel_launches", C.c_longlong)]


_lib = None


def load_library():
    """dlopen the C-ABI library; raises (loudly) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.path.abspath(LIB_PATH)
    if not os.path.exists(path):
        raise TopoError(ENODEV, "CUDA extension not built: %s is missing (run `python __graft_entry__.py build` "
                                "or `python solidspy-opt_b200/build.py`); there is no CPU fallback" % path)
    lib = C.CDLL(path)
    lib.topo_last_error.restype = C.c_char_p
    lib.topo_abi_version.restype = C.c_int
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    vp, i64p = C.c_void_p, C.POINTER(C.c_int64)
    sig = {
        "topo_device_count": [ip],
        "topo_create": [C.POINTER(vp), C.c_int, C.c_int, dp, C.c_int],
        "topo_destroy": [vp],
        "topo_set_cons": [vp, C.c_void_p],
        "topo_set_loads": [vp, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int],
        "topo_set": [vp, C.c_int, C.c_void_p, C.c_size_t],
        "topo_get": [vp, C.c_int, C.c_void_p, C.c_size_t],
        "topo_fill": [vp, C.c_int, C.c_double],
        "topo_matvec": [vp, C.c_void_p, C.c_void_p],
        "topo_matvec_device": [vp, C.c_int, C.POINTER(C.c_float)],
        "topo_solve": [vp, C.c_int, C.c_double, C.c_int, C.c_int, ip, dp],
        "topo_compliance": [vp, dp],
        "topo_equilibrium_check": [vp, ip],
        "topo_simp_scale": [vp, C.c_double, C.c_double, C.c_double],
        "topo_simp_sensitivity": [vp, C.c_double, C.c_double, C.c_double, dp],
        "topo_filter_sensitivity": [vp, C.c_double, C.c_double, C.c_double],
        "topo_oc_update": [vp, C.c_double, dp, dp, ip],
        "topo_simp_step": [vp, C.POINTER(SimpParams), dp, C.POINTER(SimpResult)],
        "topo_simp_step_host": [vp, C.POINTER(SimpParams), C.c_void_p, C.c_void_p, dp, C.POINTER(SimpResult)],
        "topo_energy_sensitivity": [vp, C.c_int],
        "topo_beso_filter": [vp, dp, dp, dp, dp, C.c_int],
        "topo_beso_nodes": [vp, dp, dp, dp],
        "topo_beso_elements": [vp, dp],
        "topo_beso_elements_radius": [vp, C.c_double, C.c_double, C.c_double],
        "topo_rank_select": [vp, C.c_int64, dp],
        "topo_beso_apply": [vp, C.c_double, C.c_void_p, C.c_int, i64p, i64p],
        "topo_eso_apply": [vp, C.c_double, C.c_void_p, C.c_int, i64p],
        "topo_beso_save_history": [vp],
        "topo_floating_elements": [vp, C.c_void_p, C.c_int, i64p],
        "topo_vonmises_sensitivity": [vp, C.c_double, C.c_double, C.c_double, C.c_double],
        "topo_strain_nodes": [vp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p],
        "topo_get_timing": [vp, C.POINTER(Timing)],
        "topo_gmg_info": [vp, ip, ip, ip],
        "topo_set_gmg": [vp, C.c_double, C.c_int, C.c_int],
        "topo_timer_start": [vp],
        "topo_timer_stop": [vp, C.POINTER(C.c_float)],
        "topo_profile_matvec": [vp, C.c_int],
        "topo_profile_read": [vp, ip, C.POINTER(C.c_float)],
        "topo_profile_report": [vp, C.c_char_p, C.c_size_t],
        "topo_set_stream": [vp, C.c_void_p],
        "topo_matvec_dev": [vp, C.c_void_p, C.c_void_p],
        "topo_diag_dev": [vp, C.c_void_p],
        "topo_vec_axpby_dev": [vp, C.c_longlong, C.c_double, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p],
        "topo_vec_mul_dev": [vp, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int],
        "topo_vec_dot_dev": [vp, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p],
        "topo_dist_init": [vp, C.c_int, C.c_int, C.c_int],
        "topo_dist_info": [vp, ip, ip, ip, ip],
        "topo_dist_xchg_ptrs": [vp, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp),
                                C.POINTER(C.c_longlong)],
        "topo_dist_xchg_finish": [vp, C.c_int, C.c_int],
        "topo_dist_gather_ptrs": [vp, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_longlong)],
        "topo_dist_scalar_ptr": [vp, C.c_int, C.POINTER(vp), ip],
        "topo_dist_read_scalars": [vp, dp, C.c_int],
        "topo_dist_setup": [vp, C.c_int],
        "topo_dist_solve_begin": [vp, C.c_double, C.c_int, C.c_int],
        "topo_dist_pcg_matvec": [vp, C.c_int],
        "topo_dist_pcg_update": [vp, C.c_int, C.c_int],
        "topo_dist_vcycle_down": [vp, C.c_int],
        "topo_dist_vcycle_tail": [vp, C.c_int],
        "topo_dist_p2p_export": [vp, C.c_void_p, C.POINTER(vp)],
        "topo_dist_p2p_attach": [vp, C.c_void_p, C.POINTER(vp)],
        "topo_dist_p2p_xchg": [vp, C.c_int, C.c_int, C.c_int],
        "topo_dist_p2p_gather": [vp, C.c_int],
        "topo_dist_p2p_allreduce": [vp, C.c_int, C.c_int, C.c_int, C.c_int],
        "topo_dist_p2p_error": [vp, ip],
        "topo_dist_p2p_fused": [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int],
        "topo_dist_p2p_gather_tail": [vp],
        "topo_dist_p2p_mode": [vp, C.c_int],
        "topo_dist_vcycle_up": [vp, C.c_int],
        "topo_dist_pcg_smooth": [vp, C.c_int, C.c_int],
        "topo_dist_pcg_finish": [vp, C.c_int],
        "topo_dist_solve_poll": [vp, ip, ip, dp],
        "topo_dist_solve_end": [vp, C.c_int],
        "topo_dist_compliance_partial": [vp],
        "topo_dist_filter_pack": [vp, C.c_double, C.c_double],
        "topo_dist_oc": [vp, C.c_int, C.c_double, C.c_double, C.c_longlong],
        "topo_dist_oc_state": [vp, ip, ip, ip, dp, dp, dp],
    }
    for name, args in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    _lib = lib
    return lib


def _check(code, allow=()):
    if code != OK and code not in allow:
        raise TopoError(code, load_library().topo_last_error().decode("utf-8", "replace"))
    return code


def device_count():
    n = C.c_int(0)
    _check(load_library().topo_device_count(C.byref(n)))
    return n.value


def _dbl(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if n is not None and a.size != n:
        raise ValueError("expected %d doubles, got %d" % (n, a.size))
    return a


class Topo:
    """One structured quad4 mesh on one GPU (owns all device memory)."""

    def __init__(self, nx, ny, ke, device=0):
        self.lib = load_library()
        self.nx, self.ny = int(nx), int(ny)
        self.nn = (self.nx + 1)*(self.ny + 1)
        self.ne = self.nx*self.ny
        ke = _dbl(ke, 64)
        h = C.c_void_p()
        _check(self.lib.topo_create(C.byref(h), self.nx, self.ny, ke.ctypes.data_as(C.POINTER(C.c_double)), int(device)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.topo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- state ----
    def set_cons(self, cons):
        """cons: (N_n, 2) BC flags as in ``nodes[:, 3:5]`` (-1 constrained, 0 free)."""
        c = np.ascontiguousarray(np.asarray(cons).reshape(self.nn, 2), dtype=np.int8)
        _check(self.lib.topo_set_cons(self.h, c.ctypes.data))

    def get_cons(self):
        out = np.empty((self.nn, 2), dtype=np.int8)
        _check(self.lib.topo_get(self.h, F_CONS, out.ctypes.data, out.nbytes))
        return out

    def set_loads(self, loads):
        """loads: (n, 3) rows [node, fx, fy] (``loadasem`` semantics)."""
        loads = np.asarray(loads)
        node = np.ascontiguousarray(loads[:, 0], dtype=np.int64)
        fx = _dbl(loads[:, 1])
        fy = _dbl(loads[:, 2])
        _check(self.lib.topo_set_loads(self.h, node.ctypes.data, fx.ctypes.data, fy.ctypes.data, node.size))

    _SIZES = {F_U: ("nn2", np.float64), F_LOAD: ("nn2", np.float64), F_RESID: ("nn2", np.float64),
              F_RHO: ("ne", np.float64), F_DC: ("ne", np.float64), F_SCALE: ("ne", np.float64),
              F_SENS: ("ne", np.float64), F_SENS_PREV: ("ne", np.float64), F_KEEP: ("ne", np.uint8),
              F_NODAL: ("nn", np.float64)}

    def _shape(self, field):
        kind, dt = self._SIZES[field]
        return {"nn2": (self.nn, 2), "nn": (self.nn,), "ne": (self.ne,)}[kind], dt

    def get(self, field):
        shape, dt = self._shape(field)
        out = np.empty(shape, dtype=dt)
        _check(self.lib.topo_get(self.h, field, out.ctypes.data, out.nbytes))
        return out

    def set(self, field, arr):
        shape, dt = self._shape(field)
        a = np.ascontiguousarray(np.asarray(arr).reshape(shape), dtype=dt)
        _check(self.lib.topo_set(self.h, field, a.ctypes.data, a.nbytes))

    def fill(self, field, value):
        _check(self.lib.topo_fill(self.h, field, float(value)))

    # ---- stages ----
    def matvec(self, u):
        u = _dbl(u, 2*self.nn)
        y = np.empty((self.nn, 2))
        _check(self.lib.topo_matvec(self.h, u.ctypes.data, y.ctypes.data))
        return y

    def matvec_device(self, repeats=10):
        ms = C.c_float(0)
        _check(self.lib.topo_matvec_device(self.h, int(repeats), C.byref(ms)))
        return ms.value

    def solve(self, precond=PRECOND_GMG, rtol=1e-10, maxit=20000, warm=True, raise_on_noconv=True):
        it, rel = C.c_int(0), C.c_double(0)
        code = _check(self.lib.topo_solve(self.h, precond, rtol, maxit, 1 if warm else 0, C.byref(it), C.byref(rel)),
                      allow=() if raise_on_noconv else (ENOCONV,))
        return it.value, rel.value, code

    def compliance(self):
        c = C.c_double(0)
        _check(self.lib.topo_compliance(self.h, C.byref(c)))
        return c.value

    def equilibrium_ok(self):
        ok = C.c_int(0)
        _check(self.lib.topo_equilibrium_check(self.h, C.byref(ok)))
        return bool(ok.value)

    def simp_scale(self, penal, emin, emax):
        _check(self.lib.topo_simp_scale(self.h, penal, emin, emax))

    def simp_sensitivity_partial(self, penal, emin, emax):
        """Sensitivity kernel only; the objective stays on the device (TOPO_S_RED[0])."""
        _check(self.lib.topo_simp_sensitivity(self.h, penal, emin, emax, None))

    def simp_sensitivity(self, penal, emin, emax):
        obj = C.c_double(0)
        _check(self.lib.topo_simp_sensitivity(self.h, penal, emin, emax, C.byref(obj)))
        return obj.value

    def filter_sensitivity(self, rmin, dx, dy):
        _check(self.lib.topo_filter_sensitivity(self.h, rmin, dx, dy))

    def oc_update(self, g, move=0.2):
        gg, ch, ns = C.c_double(g), C.c_double(0), C.c_int(0)
        _check(self.lib.topo_oc_update(self.h, move, C.byref(gg), C.byref(ch), C.byref(ns)))
        return gg.value, ch.value, ns.value

    def simp_step(self, params, g, rho_in=None, rho_out=None):
        gg, res = C.c_double(g), SimpResult()
        if rho_in is None:
            code = self.lib.topo_simp_step(self.h, C.byref(params), C.byref(gg), C.byref(res))
        else:
            code = self.lib.topo_simp_step_host(self.h, C.byref(params), rho_in.ctypes.data, rho_out.ctypes.data,
                                                C.byref(gg), C.byref(res))
        _check(code, allow=(ENOCONV,))
        return gg.value, res

    def energy_sensitivity(self):
        _check(self.lib.topo_energy_sensitivity(self.h, 1))

    def beso_filter(self, w4, w2x, w2y, we, average_with_prev):
        dp = C.POINTER(C.c_double)
        a, b, c, d = _dbl(w4, 4), _dbl(w2x, 2), _dbl(w2y, 2), _dbl(we, 4)
        _check(self.lib.topo_beso_filter(self.h, a.ctypes.data_as(dp), b.ctypes.data_as(dp), c.ctypes.data_as(dp),
                                         d.ctypes.data_as(dp), 1 if average_with_prev else 0))

    def beso_nodes(self, w4, w2x, w2y):
        dp = C.POINTER(C.c_double)
        a, b, c = _dbl(w4, 4), _dbl(w2x, 2), _dbl(w2y, 2)
        _check(self.lib.topo_beso_nodes(self.h, a.ctypes.data_as(dp), b.ctypes.data_as(dp), c.ctypes.data_as(dp)))

    def beso_elements(self, we):
        d = _dbl(we, 4)
        _check(self.lib.topo_beso_elements(self.h, d.ctypes.data_as(C.POINTER(C.c_double))))

    def beso_elements_radius(self, r_min, dx, dy):
        _check(self.lib.topo_beso_elements_radius(self.h, float(r_min), float(dx), float(dy)))

    def rank_select(self, k):
        a = C.c_double(0)
        _check(self.lib.topo_rank_select(self.h, int(k), C.byref(a)))
        return a.value

    def beso_apply(self, alpha, protect_nodes):
        p = np.ascontiguousarray(protect_nodes, dtype=np.int64)
        nk, npz = C.c_int64(0), C.c_int64(0)
        _check(self.lib.topo_beso_apply(self.h, float(alpha), p.ctypes.data, p.size, C.byref(nk), C.byref(npz)))
        return nk.value, npz.value

    def eso_apply(self, rr, protect_nodes):
        p = np.ascontiguousarray(protect_nodes, dtype=np.int64)
        nk = C.c_int64(0)
        _check(self.lib.topo_eso_apply(self.h, float(rr), p.ctypes.data, p.size, C.byref(nk)))
        return nk.value

    def floating_elements(self, support_nodes):
        ids = np.ascontiguousarray(support_nodes, dtype=np.int64)
        n = C.c_int64(0)
        _check(self.lib.topo_floating_elements(self.h, ids.ctypes.data, int(ids.size), C.byref(n)))
        return int(n.value)

    def beso_save_history(self):
        _check(self.lib.topo_beso_save_history(self.h))

    def vonmises_sensitivity(self, E, nu, dx, dy):
        _check(self.lib.topo_vonmises_sensitivity(self.h, E, nu, dx, dy))

    def strain_nodes(self, E, nu, dx, dy, present_only=True):
        """(E_nodes, S_nodes), each (N_n, 3): nodal strains / stresses of the current displacements (the
        ``pos.strain_nodes`` pair of the reference, optimize.py:63,86,163,183,262,300)."""
        eps = np.empty((3, self.nn))
        sig = np.empty((3, self.nn))
        _check(self.lib.topo_strain_nodes(self.h, E, nu, dx, dy, 1 if present_only else 0, eps.ctypes.data,
                                          sig.ctypes.data))
        return np.ascontiguousarray(eps.T), np.ascontiguousarray(sig.T)

    def timing(self):
        t = Timing()
        _check(self.lib.topo_get_timing(self.h, C.byref(t)))
        return t

    def gmg_levels(self):
        n = C.c_int(0)
        nxs = (C.c_int*32)()
        nys = (C.c_int*32)()
        _check(self.lib.topo_gmg_info(self.h, C.byref(n), nxs, nys))
        return [(nxs[i], nys[i]) for i in range(n.value)]

    def set_gmg(self, omega=0.7, nu_pre=1, nu_post=1):
        _check(self.lib.topo_set_gmg(self.h, omega, nu_pre, nu_post))

    # ---- row-slab mode (include/topo_b200.h, "row-slab SIMP iteration") ----
    def dist_init(self, rank, world, K=0):
        _check(self.lib.topo_dist_init(self.h, int(rank), int(world), int(K)))
        r, w, k, n = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
        _check(self.lib.topo_dist_info(self.h, C.byref(r), C.byref(w), C.byref(k), C.byref(n)))
        return k.value, n.value

    def dist_xchg_ptrs(self, what, level=0):
        a, b, c, d = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        n = C.c_longlong(0)
        _check(self.lib.topo_dist_xchg_ptrs(self.h, what, level, C.byref(a), C.byref(b), C.byref(c), C.byref(d), C.byref(n)))
        return a.value, b.value, c.value, d.value, n.value

    def dist_xchg_finish(self, what, level=0):
        _check(self.lib.topo_dist_xchg_finish(self.h, what, level))

    def dist_gather_ptrs(self, what):
        a, b = C.c_void_p(), C.c_void_p()
        n = C.c_longlong(0)
        _check(self.lib.topo_dist_gather_ptrs(self.h, what, C.byref(a), C.byref(b), C.byref(n)))
        return a.value, b.value, n.value

    def dist_scalar_ptr(self, what):
        a, n = C.c_void_p(), C.c_int(0)
        _check(self.lib.topo_dist_scalar_ptr(self.h, what, C.byref(a), C.byref(n)))
        return a.value, n.value

    def dist_read_scalars(self, n):
        out = np.zeros(n)
        _check(self.lib.topo_dist_read_scalars(self.h, out.ctypes.data_as(C.POINTER(C.c_double)), n))
        return out

    def dist_call(self, name, *args):
        """Thin passthrough for the segment calls (topo_dist_<name>)."""
        _check(getattr(self.lib, "topo_dist_" + name)(self.h, *args))

    def dist_p2p_export(self):
        """(64-byte CUDA IPC handle, device address) of this slab's peer-memory region."""
        buf = (C.c_ubyte*64)()
        ptr = C.c_void_p()
        _check(self.lib.topo_dist_p2p_export(self.h, buf, C.byref(ptr)))
        return bytes(buf), ptr.value

    def dist_p2p_attach(self, ipc_handles, local_ptrs):
        """ipc_handles: list of 64-byte handles by rank (or None); local_ptrs: list of addresses by rank (0 = use IPC)."""
        world = len(local_ptrs)
        blob = b"".join(hh if hh is not None else bytes(64) for hh in (ipc_handles or [None]*world))
        arr = (C.c_void_p*world)(*[C.c_void_p(p or None) for p in local_ptrs])
        _check(self.lib.topo_dist_p2p_attach(self.h, blob, arr))

    def dist_p2p_error(self):
        e = C.c_int(0)
        _check(self.lib.topo_dist_p2p_error(self.h, C.byref(e)))
        return e.value

    def dist_solve_poll(self):
        d, it, rel = C.c_int(0), C.c_int(0), C.c_double(0)
        _check(self.lib.topo_dist_solve_poll(self.h, C.byref(d), C.byref(it), C.byref(rel)))
        return d.value, it.value, rel.value

    def dist_oc_state(self):
        d, ne, st = C.c_int(0), C.c_int(0), C.c_int(0)
        gt, ch, lm = C.c_double(0), C.c_double(0), C.c_double(0)
        _check(self.lib.topo_dist_oc_state(self.h, C.byref(d), C.byref(ne), C.byref(st), C.byref(gt), C.byref(ch),
                                           C.byref(lm)))
        return d.value, ne.value, st.value, gt.value, ch.value, lm.value

    def timer_start(self):
        _check(self.lib.topo_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float(0)
        _check(self.lib.topo_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def profile_matvec(self, enable):
        _check(self.lib.topo_profile_matvec(self.h, 1 if enable else 0))

    def profile_read(self):
        n, ms = C.c_int(0), C.c_float(0)
        _check(self.lib.topo_profile_read(self.h, C.byref(n), C.byref(ms)))
        return n.value, ms.value

    def profile_report(self):
        """[(kernel, launches, total_ms)] for everything launched since profile_matvec(True)."""
        buf = C.create_string_buffer(1 << 16)
        _check(self.lib.topo_profile_report(self.h, buf, len(buf)))
        out = []
        for line in buf.value.decode().splitlines():
            name, n, ms = line.rsplit(" ", 2)
            out.append((name, int(n), float(ms)))
        return out

    # ---- device-pointer entry points (multi-GPU building blocks) ----
    def set_stream(self, cuda_stream):
        _check(self.lib.topo_set_stream(self.h, C.c_void_p(int(cuda_stream))))

    def matvec_dev(self, u_ptr, y_ptr):
        _check(self.lib.topo_matvec_dev(self.h, C.c_void_p(int(u_ptr)), C.c_void_p(int(y_ptr))))

    def diag_dev(self, d_ptr):
        _check(self.lib.topo_diag_dev(self.h, C.c_void_p(int(d_ptr))))

    def axpby_dev(self, n_nodes, alpha, a_ptr, beta, b_ptr, c_ptr):
        _check(self.lib.topo_vec_axpby_dev(self.h, int(n_nodes), float(alpha), C.c_void_p(int(a_ptr)), float(beta),
                                           C.c_void_p(int(b_ptr)), C.c_void_p(int(c_ptr))))

    def mul_dev(self, n_nodes, a_ptr, b_ptr, c_ptr, divide=False):
        _check(self.lib.topo_vec_mul_dev(self.h, int(n_nodes), C.c_void_p(int(a_ptr)), C.c_void_p(int(b_ptr)),
                                         C.c_void_p(int(c_ptr)), 1 if divide else 0))

    def dot_dev(self, a_ptr, b_ptr, row_lo, row_hi, out_ptr):
        _check(self.lib.topo_vec_dot_dev(self.h, C.c_void_p(int(a_ptr)), C.c_void_p(int(b_ptr)), int(row_lo), int(row_hi),
                                         C.c_void_p(int(out_ptr))))
