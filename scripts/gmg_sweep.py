"""Sweep multigrid smoother parameters / tolerance on the headline mesh: PCG iterations and time per design iteration."""
import os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "solidspy-opt_b200"))
import numpy as np
from solidspy_opt import _capi
from solidspy_opt.Utils.solver import kloc_simp
nx, ny = 2048, 1024
nit = int(sys.argv[1]) if len(sys.argv) > 1 else 30
t = _capi.Topo(nx, ny, kloc_simp(206.8e9, 0.28))
cons = np.zeros((t.nn, 2), np.int8); cons[::nx + 1] = -1
t.set_cons(cons); t.set_loads(np.array([[(nx+1)*(ny//4) - 1, 0.0, -1.0]]))
for omega, nu1, nu2, rtol in [(0.7, 1, 1, 1e-9)]:
    t.set_gmg(omega, nu1, nu2)
    t.fill(_capi.F_RHO, 0.5); t.set(_capi.F_U, np.zeros(2*t.nn))
    p = _capi.SimpParams(penal=3.0, emin=1e-9, emax=1.0, rmin=4.0, dx=1.0, dy=1.0, move=0.2, rtol=rtol,
                         precond=_capi.PRECOND_GMG, maxit=3000, warm=1)
    g = 0.0; its = []; t0 = time.time(); comp = []
    for i in range(nit):
        g, res = t.simp_step(p, g); its.append(res.cg_iters); comp.append(res.compliance)
    dt = time.time() - t0
    print("omega %.1f nu (%d,%d) rtol %.0e: %.1f ms/iter, cg total %d, its %s, c_last %.8e" % (omega, nu1, nu2, rtol, dt/nit*1e3, sum(its), its[::3], comp[-1]), flush=True)
