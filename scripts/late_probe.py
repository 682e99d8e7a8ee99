"""PCG iteration counts and time per design iteration deep into a SIMP run (2048x1024, iterations 100..119 by default):
the experiment behind bench.py's `late` block.  Knobs come from the environment (TOPO_OMEGA, TOPO_NU_COARSE, ...).
    python scripts/late_probe.py [first] [count]"""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "solidspy-opt_b200"))
import numpy as np
from solidspy_opt import _capi
from solidspy_opt.Utils.solver import kloc_simp
nx, ny = 2048, 1024
first = int(sys.argv[1]) if len(sys.argv) > 1 else 100
count = int(sys.argv[2]) if len(sys.argv) > 2 else 20
t = _capi.Topo(nx, ny, kloc_simp(206.8e9, 0.28))
cons = np.zeros((t.nn, 2), np.int8); cons[::nx + 1] = -1
t.set_cons(cons); t.set_loads(np.array([[(nx+1)*(ny//4) - 1, 0.0, -1.0]]))
t.fill(_capi.F_RHO, 0.5)
p = _capi.SimpParams(penal=3.0, emin=1e-9, emax=1.0, rmin=4.0, dx=1.0, dy=1.0, move=0.2, rtol=1e-9,
                     precond=_capi.PRECOND_GMG, maxit=3000, warm=1)
g, its = 0.0, []
for i in range(first + count):
    if i == first:
        t.timer_start()
    g, res = t.simp_step(p, g)
    its.append(res.cg_iters)
ms = t.timer_stop()/count
knobs = {k: v for k, v in os.environ.items() if k.startswith("TOPO_")}
print(knobs, "ms/step %.2f" % ms, "it/s %.1f" % (1e3/ms), "cg", its[first:first + 3], "...", its[-3:], "max", max(its),
      "at 20:", its[20], "compliance %.10e" % res.compliance, "change %.4f" % res.change, flush=True)
