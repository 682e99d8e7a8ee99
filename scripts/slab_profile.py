"""Where does a row-slab PCG iteration spend its time BEFORE any communication?  One GPU, one mesh, two paths:
  (a) the ordinary single-handle SIMP step,
  (b) SlabSimp(world=1, dist=None): the same mesh as ONE slab in row-slab mode (segment calls, no neighbours).
Both run the same arithmetic; the per-kernel event profile of each is printed side by side.
    python scripts/slab_profile.py 8192 2048 [steps]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "solidspy-opt_b200"))
from solidspy_opt import _capi                                   # noqa: E402
from solidspy_opt.distributed import SlabSimp                     # noqa: E402
from solidspy_opt.optimize import kloc_simp                       # noqa: E402


def model(nx, ny):
    cons = np.zeros((ny + 1, nx + 1, 2), dtype=np.int8)
    cons[:, 0, :] = -1
    loads = np.array([[(nx + 1)*(ny//2) + nx, 0.0, -1.0]])
    return cons.reshape(-1, 2), loads, kloc_simp(206.8e9, 0.28)


def table(report, iters):
    rows = sorted(report, key=lambda r: -r[2])
    tot = sum(r[2] for r in rows)
    out = ["    total %.3f ms over %d PCG iterations = %.1f us / iteration" % (tot, iters, tot*1e3/max(iters, 1))]
    for nm, n, ms in rows[:22]:
        out.append("    %-28s n %5d  avg %8.2f us  per PCG it %8.2f us" % (nm, n, ms*1e3/n, ms*1e3/max(iters, 1)))
    return "\n".join(out)


def main():
    nx, ny = int(sys.argv[1]), int(sys.argv[2])
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    warm = 3
    cons, loads, ke = model(nx, ny)
    params = _capi.SimpParams(penal=3.0, emin=1e-9, emax=1.0, rmin=4.0, dx=1.0, dy=1.0, move=0.2, rtol=1e-9,
                              precond=_capi.PRECOND_GMG, maxit=2000, warm=1)
    with _capi.Topo(nx, ny, ke) as topo:
        topo.set_cons(cons)
        topo.set_loads(loads)
        topo.fill(_capi.F_RHO, 0.5)
        g, its = 0.0, 0
        for k in range(warm + steps):
            if k == warm:
                topo.timer_start()
            g, r = topo.simp_step(params, g)
            if k >= warm:
                its += r.cg_iters
        ms = topo.timer_stop()
        print("(a) single handle %dx%d: %.2f ms per design iteration (graph replay), %d PCG iterations in %d steps" % (nx, ny, ms/steps, its, steps))
        topo.fill(_capi.F_RHO, 0.5)
        topo.set(_capi.F_U, np.zeros(2*(nx + 1)*(ny + 1)))
        g, its = 0.0, 0
        for k in range(warm + steps):
            if k == warm:
                topo.profile_matvec(True)
            g, r = topo.simp_step(params, g)
            if k >= warm:
                its += r.cg_iters
        print(table(topo.profile_report(), its))
        topo.profile_matvec(False)
    sim = SlabSimp(nx, ny, ke, cons, loads, rmin=4.0, dist=None, world=1, device=0, rtol=1e-9)
    t = sim.slabs[0]
    g, its = 0.0, 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for k in range(warm + steps):
        if k == warm:
            with torch.cuda.stream(sim.stream):
                ev0.record()
        g, r = sim.step(g)
        if k >= warm:
            its += r["cg_iters"]
    with torch.cuda.stream(sim.stream):
        ev1.record()
    ev1.synchronize()
    print("(b) one slab in row-slab mode (eager launches, K = %d): %.2f ms per design iteration, %d PCG iterations, stages %s"
          % (sim.K, ev0.elapsed_time(ev1)/steps, its, {k: round(v, 2) for k, v in r["stage_ms"].items()}))
    sim.set_rho(np.full(nx*ny, 0.5))
    t.set(_capi.F_U, np.zeros(2*(nx + 1)*(ny + 1)))
    g, its = 0.0, 0
    for k in range(warm + steps):
        if k == warm:
            t.profile_matvec(True)
        g, r = sim.step(g)
        if k >= warm:
            its += r["cg_iters"]
    print(table(t.profile_report(), its))
    sim.close()


if __name__ == "__main__":
    main()
