"""Diagnostic: the fifth ESO_stress solve on 96x40, cold versus warm-started from the previous structure's solution."""
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
def state(k):
    scale = np.zeros(nx*ny); scale[ref["els_ids"][k]] = 1.0
    return scale, np.asarray(ref["cons"][k]).reshape(-1, 2)
t = _capi.Topo(nx, ny, ke)
t.set_loads(loads)
for k in range(4):
    scale, cons = state(k)
    t.set_cons(cons); t.set(_capi.F_SCALE, scale)
    it, rel, code = t.solve(_capi.PRECOND_GMG, 1e-12, 1000, warm=True, raise_on_noconv=False)
    print("warm chain, state", k, "present", int(scale.sum()), "its", it, "relres %.2e" % rel, "code", code)
t.close()
t = _capi.Topo(nx, ny, ke)
t.set_loads(loads)
scale, cons = state(3)
t.set_cons(cons); t.set(_capi.F_SCALE, scale)
it, rel, code = t.solve(_capi.PRECOND_GMG, 1e-12, 1000, warm=False, raise_on_noconv=False)
print("cold, state 3: its", it, "relres %.2e" % rel)
# --- which part of the warm start hurts? ---
t.close()
def fresh(k):
    tt = _capi.Topo(nx, ny, ke); tt.set_loads(loads)
    scale, cons = state(k); tt.set_cons(cons); tt.set(_capi.F_SCALE, scale)
    return tt, scale, cons
t2, s2, c2 = fresh(2)
t2.solve(_capi.PRECOND_GMG, 1e-12, 1000, warm=False)
u2 = t2.get(_capi.F_U).copy(); t2.close()
t3, s3, c3 = fresh(3)
free3 = (c3.reshape(-1) == 0).astype(float)
for name, x0 in (("u2 masked by the new BC flags", (u2.reshape(-1)*free3)), ("1e-3 * u2 masked", 1e-3*(u2.reshape(-1)*free3)),
                 ("random 1e-12 noise masked", 1e-12*np.random.default_rng(0).standard_normal(u2.size)*free3)):
    t3.set(_capi.F_U, x0.reshape(-1, 2))
    it, rel, code = t3.solve(_capi.PRECOND_GMG, 1e-12, 600, warm=True, raise_on_noconv=False)
    print("x0 =", name, ": its", it, "relres %.2e" % rel)
for n in (1, 5, 20, 60, 100, 150):
    t3.set(_capi.F_U, (u2.reshape(-1)*free3).reshape(-1, 2))
    it, rel, code = t3.solve(_capi.PRECOND_GMG, 1e-12, n, warm=True, raise_on_noconv=False)
    print("warm, maxit", n, "relres %.3e" % rel)
