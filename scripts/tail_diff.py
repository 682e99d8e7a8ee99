"""Diagnostic: one PCG iteration (u = alpha * M^-1 r0) with the multigrid tail starting at different levels must agree
when every level does V(1,1); shows where they differ on a hard 0/1 ESO design."""
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
k = 3
scale = np.zeros(nx*ny); scale[ref["els_ids"][k]] = 1.0
cons = np.asarray(ref["cons"][k]).reshape(-1, 2)
os.environ["TOPO_NU_TAIL_SMALL"] = "1"; os.environ["TOPO_NU_TAIL_BIG"] = "1"
out = {}
for tm in ("100", "500", "10000"):
    os.environ["TOPO_TAIL_MAX"] = tm
    t = _capi.Topo(nx, ny, ke)
    t.set_cons(cons); t.set_loads(loads); t.set(_capi.F_SCALE, scale)
    it, rel, code = t.solve(_capi.PRECOND_GMG, 1e-12, int(os.environ.get("NIT", "1")), warm=False, raise_on_noconv=False)
    out[tm] = t.get(_capi.F_U).reshape(ny + 1, nx + 1, 2).copy()
    print("tail_max", tm, "levels", t.gmg_levels(), "relres after it: %.3e" % rel)
    t.close()
a, b, c = out["100"], out["500"], out["10000"]
for name, x in (("500 vs 100", b), ("10000 vs 100", c)):
    d = np.abs(x - a).max(axis=2)
    print(name, "max|du| %.3e of max|u| %.3e" % (d.max(), np.abs(a).max()))
    idx = np.argwhere(d > 1e-6*np.abs(a).max())
    print("  nodes differing:", len(idx), "rows", sorted(set(idx[:, 0]))[:20], "cols", sorted(set(idx[:, 1]))[:20])
