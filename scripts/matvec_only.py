"""Launch the stage-1 kernel a few times on the 2048x1024 mesh (ncu target)."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "solidspy-opt_b200"))
import numpy as np
from solidspy_opt import _capi
from solidspy_opt.Utils.solver import kloc_simp
nx, ny = int(sys.argv[1]) if len(sys.argv) > 1 else 2048, int(sys.argv[2]) if len(sys.argv) > 2 else 1024
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
t = _capi.Topo(nx, ny, kloc_simp(206.8e9, 0.28))
cons = np.zeros((t.nn, 2), np.int8); cons[::nx + 1] = -1
t.set_cons(cons); t.fill(_capi.F_RHO, 0.5); t.simp_scale(3.0, 1e-9, 1.0)
t.set(_capi.F_U, np.random.default_rng(0).standard_normal(2*t.nn))
t.matvec_device(5)
ms = t.matvec_device(reps)
print("matvec %.2f us  %.0f GB/s algorithmic" % (ms*1e3, (32*t.nn + 8*t.ne)/ms/1e6))
