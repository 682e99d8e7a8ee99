import os, sys, ctypes as C
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "solidspy-opt_b200"))
import numpy as np
from solidspy_opt import _capi
from solidspy_opt.Utils.solver import kloc_simp
nx, ny = int(sys.argv[1]), int(sys.argv[2])
t = _capi.Topo(nx, ny, kloc_simp(206.8e9, 0.28))
cons = np.zeros((t.nn, 2), np.int8); cons[::nx + 1] = -1
t.set_cons(cons); t.fill(_capi.F_RHO, 0.5); t.simp_scale(3.0, 1e-9, 1.0)
t.set_loads(np.array([[(nx+1)*(ny//4) - 1, 0.0, -1.0]]))
lib = t.lib
lib.topo_debug_tail_stamps.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
buf = (C.c_ulonglong*512)(); n = C.c_int(0)
lib.topo_debug_tail_stamps(t.h, buf, 511, C.byref(n))     # enables
print(t.solve(_capi.PRECOND_GMG, 1e-10, 2000, warm=False))
lib.topo_debug_tail_stamps(t.h, buf, 511, C.byref(n))
st = np.array(buf[:n.value], dtype=np.int64)
print("levels", t.gmg_levels())
print("stages", n.value, "total %.1f us" % ((st[-1]-st[0])/1e3))
print("stage durations (us):", np.round(np.diff(st)/1e3, 2).tolist())
