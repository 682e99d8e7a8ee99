"""Three SIMP design iterations on the headline mesh (ncu target: run with TOPO_NO_GRAPH=1)."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "solidspy-opt_b200"))
import numpy as np
from solidspy_opt import _capi
from solidspy_opt.Utils.solver import kloc_simp
nx, ny = 2048, 1024
n_it = int(sys.argv[1]) if len(sys.argv) > 1 else 3
t = _capi.Topo(nx, ny, kloc_simp(206.8e9, 0.28))
cons = np.zeros((t.nn, 2), np.int8); cons[::nx + 1] = -1
t.set_cons(cons); t.set_loads(np.array([[(nx+1)*(ny//4) - 1, 0.0, -1.0]]))
t.fill(_capi.F_RHO, 0.5)
p = _capi.SimpParams(penal=3.0, emin=1e-9, emax=1.0, rmin=4.0, dx=1.0, dy=1.0, move=0.2, rtol=1e-9,
                     precond=_capi.PRECOND_GMG, maxit=3000, warm=1)
g = 0.0
for i in range(n_it):
    l0 = t.timing().kernel_launches
    g, res = t.simp_step(p, g)
    print("iteration", i, "launches", t.timing().kernel_launches - l0, "cg", res.cg_iters, "compliance", res.compliance, flush=True)
