"""Quick GPU check of the row-slab path on one GPU (several slabs in one process):
    python scripts/slab_check.py [nx ny world K]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "solidspy-opt_b200"))
import torch  # noqa: E402
from solidspy_opt import _capi  # noqa: E402
from solidspy_opt.Utils.beams import beam  # noqa: E402
from solidspy_opt.distributed import SlabSimp  # noqa: E402
from solidspy_opt.optimize import kloc_simp  # noqa: E402

nx, ny, world, K = (int(a) for a in (sys.argv[1:5] + ["256", "128", "4", "0"][len(sys.argv) - 1:]))
nodes, mats, els, loads, _ = beam(L=float(nx), H=float(ny), nx=nx, ny=ny, dirs=np.array([[0, -1]]),
                                  positions=np.array([[ny//2 + 1, 1]]), n=1)
ke = kloc_simp(mats[0, 0], mats[0, 1])
rho = np.random.default_rng(3).uniform(0.05, 1.0, nx*ny)
sim = SlabSimp(nx, ny, ke, nodes[:, 3:5], loads, rmin=2.5, world=world, K=K, rtol=1e-10)
print("K =", sim.K, "levels =", sim.n_levels, [t.gmg_levels() for t in sim.slabs[:1]])
sim.set_rho(rho)
topo = _capi.Topo(nx, ny, ke)
topo.set_cons(nodes[:, 3:5])
topo.set_loads(loads)
topo.set(_capi.F_RHO, rho)
topo.simp_scale(3.0, 1e-9, 1.0)
it1, rel1, _ = topo.solve(_capi.PRECOND_GMG, 1e-10, 2000, warm=False)
u1 = topo.get(_capi.F_U).reshape(ny + 1, nx + 1, 2)
with torch.cuda.stream(sim.stream):
    for t in sim.slabs:
        t.simp_scale(3.0, 1e-9, 1.0)
    sim.setup()
    it2, rel2 = sim.solve(warm=False)
u2 = sim.local_u()
print("single: %d its relres %.2e | slabs: %d its relres %.2e | max|du|/max|u| = %.2e" %
      (it1, rel1, it2, rel2, np.abs(u1 - u2).max()/np.abs(u1).max()))
params = _capi.SimpParams(penal=3.0, emin=1e-9, emax=1.0, rmin=2.5, dx=1.0, dy=1.0, move=0.2, rtol=1e-10,
                          precond=_capi.PRECOND_GMG, maxit=2000, warm=1)
topo.fill(_capi.F_RHO, 0.5)
sim.set_rho(np.full(nx*ny, 0.5))
g1 = g2 = 0.0
for k in range(4):
    g1, res = topo.simp_step(params, g1)
    t0 = time.perf_counter()
    g2, r2 = sim.step(g2)
    dt = time.perf_counter() - t0
    print("step %d: c %.10e / %.10e  its %d/%d  oc %d/%d  g %.3e/%.3e  max|drho| %.2e  (%.1f ms)" %
          (k, res.compliance, r2["compliance"], res.cg_iters, r2["cg_iters"], res.oc_steps, r2["oc_steps"], g1, g2,
           np.abs(sim.local_rho().reshape(-1) - topo.get(_capi.F_RHO)).max(), dt*1e3))
