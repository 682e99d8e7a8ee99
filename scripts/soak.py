"""Long SIMP trajectories (robustness of the multigrid-PCG late in the optimisation):
    python scripts/soak.py nx ny n_iter"""
import os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "solidspy-opt_b200"))
import numpy as np
from solidspy_opt import _capi
from solidspy_opt.Utils.solver import kloc_simp
nx, ny, n_it = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
t = _capi.Topo(nx, ny, kloc_simp(206.8e9, 0.28))
cons = np.zeros((t.nn, 2), np.int8); cons[::nx + 1] = -1
t.set_cons(cons); t.set_loads(np.array([[(nx+1)*(ny//4) - 1, 0.0, -1.0]]))
t.fill(_capi.F_RHO, 0.5)
p = _capi.SimpParams(penal=3.0, emin=1e-9, emax=1.0, rmin=4.0, dx=1.0, dy=1.0, move=0.2, rtol=1e-9,
                     precond=_capi.PRECOND_GMG, maxit=5000, warm=1)
g = 0.0
t0 = time.perf_counter()
its, cs = [], []
for i in range(n_it):
    g, res = t.simp_step(p, g)
    its.append(res.cg_iters); cs.append(res.compliance)
    if res.status != 0 or not np.isfinite(res.compliance):
        print("FAILED at design iteration", i, "status", res.status, "relres", res.relres); break
dt = time.perf_counter() - t0
rho = t.get(_capi.F_RHO)
print("%dx%d: %d design iterations in %.2f s (%.1f it/s); PCG iterations min/max %d/%d, last 10: %s" %
      (nx, ny, len(its), dt, len(its)/dt, min(its), max(its), its[-10:]))
print("compliance first/last %.6e / %.6e, change %.4f, rho in [%.3g, %.3g], mean %.6f, oc steps %d" %
      (cs[0], cs[-1], res.change, rho.min(), rho.max(), rho.mean(), res.oc_steps))
