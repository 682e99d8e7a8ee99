"""Diagnostic: multigrid-PCG on the 96x40 ESO state 3 from a tiny random initial guess, under several knobs."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "solidspy-opt_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import t1
from solidspy_opt import _capi
nx, ny = 96, 40
dirs, pos = np.array([[0, -1]]), np.array([[ny//4 + 1, 1]])
ref = {}
with np.errstate(all="ignore"):
    t1.eso_stress(nx, ny, nx, ny, dirs, pos, 12, 0.005, 0.05, 0.5, history=ref)
nodes, mats, els, loads, BC = t1.beam(L=nx, H=ny, nx=nx, ny=ny, dirs=dirs, positions=pos)
ke = t1.ke_gauss(nodes[els[0, 3:7], 1:3], mats[0, 0], mats[0, 1])
k = int(os.environ.get("STATE", "3"))
scale = np.zeros(nx*ny); scale[ref["els_ids"][k]] = 1.0
cons = np.asarray(ref["cons"][k]).reshape(-1, 2)
free = (cons.reshape(-1) == 0).astype(float)
x0 = (1e-12*np.random.default_rng(0).standard_normal(free.size)*free).reshape(-1, 2)
cfgs = [{}, {"TOPO_NO_GRAPH": "1"}, {"TOPO_TAIL_MAX": "100"}, {"TOPO_TAIL_MAX": "0"}, {"TOPO_COARSEST_NODES": "10"}]
for cfg in cfgs:
    for kk in ("TOPO_NO_GRAPH", "TOPO_TAIL_MAX", "TOPO_COARSEST_NODES"):
        os.environ.pop(kk, None)
    os.environ.update(cfg)
    t = _capi.Topo(nx, ny, ke)
    t.set_cons(cons); t.set_loads(loads); t.set(_capi.F_SCALE, scale)
    out = []
    for n in (20, 40, 60, 80, 100, 140, 200):
        t.set(_capi.F_U, x0)
        it, rel, code = t.solve(_capi.PRECOND_GMG, 1e-12, n, warm=True, raise_on_noconv=False)
        out.append("%d:%.1e" % (it, rel))
    t.set(_capi.F_U, x0)
    itj, relj, _ = t.solve(_capi.PRECOND_JACOBI, 1e-12, 3000, warm=True, raise_on_noconv=False)
    print(cfg, "levels", len(t.gmg_levels()), " ".join(out), "| jacobi %d:%.1e" % (itj, relj))
    t.close()
