"""Diagnostic: direction of M^-1 r (one multigrid V(1,1) cycle) on the device versus a numpy restatement, for a smooth
and a rough right-hand side on a hard 0/1 ESO design."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "solidspy-opt_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, scipy.sparse as sp
import t1, gmg_proto as G
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
free = (cons.reshape(-1) == 0).astype(float)
K = G.full_K(nx, ny, scale, ke); M = sp.diags(free); A0 = (M @ K @ M).tocsr()
def p1(nf):
    """device transfer rule (gmg_cycle.cu prolong_taps): nc = ceil(nf/2) coarse cells, coarse node I at fine min(2I, nf)"""
    nc = (nf + 1)//2
    P = sp.lil_matrix((nf + 1, nc + 1))
    for i in range(nf + 1):
        if i % 2 == 0: P[i, i//2] = 1.0
        elif i == nf: P[i, (i + 1)//2] = 1.0
        else:
            P[i, (i - 1)//2] = 0.5; P[i, (i + 1)//2] = 0.5
    return P.tocsr(), nc
levels = []; A = A0; lx, ly = nx, ny
while True:
    d = A.diagonal(); dinv = np.zeros_like(d); dinv[d > 0] = 1/d[d > 0]
    lev = {"A": A, "dinv": dinv}; levels.append(lev)
    if (lx + 1)*(ly + 1) <= 48 or max(lx, ly) <= 1: break
    Px, cx = p1(lx); Py, cy = p1(ly)
    P = sp.kron(sp.kron(Py, Px), sp.identity(2)).tocsr()
    lev["P"] = P; A = (P.T @ A @ P).tocsr(); lx, ly = cx, cy
print("numpy levels end at", lx, ly, "n levels", len(levels))
Ac = levels[-1]["A"].toarray(); dd = np.diag(Ac).copy(); dead = dd <= 0
Ac[dead, :] = 0; Ac[:, dead] = 0; Ac[dead, dead] = 1
Ainv = np.linalg.pinv(Ac, rcond=1e-12)
def vc(b, l=0, om=0.7):
    lev = levels[l]
    if l == len(levels) - 1:
        x = Ainv @ b; x[dead] = 0; return x
    A, dinv = lev["A"], lev["dinv"]
    x = om*dinv*b; r = (b - A @ x)*(dinv != 0)
    xc = vc(lev["P"].T @ r, l + 1); x = x + (lev["P"] @ xc)*(dinv != 0)
    return x + om*dinv*(b - A @ x)
os.environ["TOPO_NU_TAIL_SMALL"] = "1"; os.environ["TOPO_NU_TAIL_BIG"] = "1"
rng = np.random.default_rng(1)
f_smooth = np.zeros(free.size)
for l in loads: f_smooth[2*int(l[0]) + 1] = l[2]
tests = {"smooth (the load)": f_smooth*free, "rough (random)": rng.standard_normal(free.size)*free}
for tm in ("10000", "100"):
    os.environ["TOPO_TAIL_MAX"] = tm
    t = _capi.Topo(nx, ny, ke)
    t.set_cons(cons); t.set(_capi.F_SCALE, scale)
    for name, r in tests.items():
        t.set_loads(np.zeros((0, 3)))
        t.set(_capi.F_LOAD, r.reshape(-1, 2))
        t.solve(_capi.PRECOND_GMG, 1e-30, 1, warm=False, raise_on_noconv=False)
        u1 = t.get(_capi.F_U).reshape(-1)
        z = vc(r)
        alpha = (u1 @ z)/(z @ z)
        diff = np.abs(u1 - alpha*z).reshape(ny + 1, nx + 1, 2).max(axis=2)
        rel = diff.max()/np.abs(alpha*z).max()
        idx = np.argwhere(diff > 0.01*np.abs(alpha*z).max())
        print("tail_max", tm, "|", name, ": max rel diff %.2e" % rel, "nodes > 1%%: %d" % len(idx),
              "rows", sorted(set(idx[:, 0]))[:12], "cols", sorted(set(idx[:, 1]))[:12])
    t.close()
