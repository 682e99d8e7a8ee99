"""CPU study behind DESIGN.md section 4 "Late in an optimisation" (ORACLE-SIDE EXPERIMENT, numpy/scipy only; nothing in the
product imports this).  Smoothers of the Galerkin multigrid (oracle/gmg_proto.py) on a converged design: damped / l1 Jacobi, Chebyshev, 4-colour Gauss-Seidel.
Run from the repo root:  python oracle/experiments/smoothers.py   (after gen_design.py 256 128 100)"""
import os
import sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, HERE)
OUT = os.environ.get("PROTO_DIR", "/tmp/proto")
os.makedirs(OUT, exist_ok=True)
import time
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import gmg_proto as G
import t1

def setup(nx, ny, rho):
    ke = t1.kloc_closed_form(206.8e9, 0.28)
    scale = 1e-9 + rho**3*(1 - 1e-9)
    fixed = np.zeros((ny + 1, nx + 1, 2), bool); fixed[:, 0, :] = True
    f = np.zeros((ny + 1, nx + 1, 2)); 
    node = (nx + 1)*(ny//4) - 1
    f.reshape(-1, 2)[node, 1] = -1.0
    return ke, scale, fixed.reshape(-1), f.reshape(-1)

class GMG2(G.GMG):
    """adds: coloured Gauss-Seidel (4 colours by node parity; forward going down, backward going up), l1-Jacobi"""
    def __init__(self, *a, gs_levels=99, **kw):
        sm = kw.get("smoother", "jacobi")
        kw["smoother"] = "jacobi" if sm in ("gs", "l1", "gsb") else sm
        super().__init__(*a, **kw)
        self.sm2 = sm
        self.gs_levels = gs_levels
        for lev in self.levels[:-1]:
            nxl, nyl = lev["nx"], lev["ny"]
            r, c = np.divmod(np.arange((nxl + 1)*(nyl + 1)), nxl + 1)
            col = (r % 2)*2 + (c % 2)
            lev["cidx"] = [np.repeat(2*np.nonzero(col == k)[0], 2) + np.tile([0, 1], (col == k).sum()) for k in range(4)]
            A = lev["A"]
            lev["Arows"] = [A[idx] for idx in lev["cidx"]]
            if sm == "gsb":     # 2x2 node-block inverse
                d = A.diagonal().reshape(-1, 2)
                n2 = np.arange(0, A.shape[0], 2)
                off = np.asarray(A[n2, n2 + 1]).ravel()
                det = d[:, 0]*d[:, 1] - off*off
                lev["binv"] = (d[:, 1]/det, -off/det, d[:, 0]/det)
            if sm == "l1":
                Aabs = abs(A)
                lev["l1inv"] = 1.0/np.asarray(Aabs.sum(1)).ravel()
    def smooth(self, lev, x, b, n, backward=False):
        li = self.levels.index(lev)
        if self.sm2 in ("gs", "gsb") and li < self.gs_levels:
            if x is None: x = np.zeros_like(b)
            order = [3, 2, 1, 0] if backward else [0, 1, 2, 3]
            for _ in range(n):
                for k in order:
                    idx = lev["cidx"][k]
                    res = b[idx] - lev["Arows"][k] @ x
                    if self.sm2 == "gsb":
                        a, o, d = lev["binv"]
                        nidx = idx[::2]//2
                        rx, ry = res[0::2], res[1::2]
                        x[idx[0::2]] += a[nidx]*rx + o[nidx]*ry
                        x[idx[1::2]] += o[nidx]*rx + d[nidx]*ry
                    else:
                        x[idx] += lev["dinv"][idx]*res
            return x
        if self.sm2 == "l1":
            for _ in range(n):
                x = x + lev["l1inv"]*(b - lev["A"] @ x) if x is not None else lev["l1inv"]*b
            return x
        return super().smooth(lev, x, b, n)
    def vcycle(self, b, l=0):
        lev = self.levels[l]
        if l == len(self.levels) - 1:
            return self.coarse.solve(b)
        if self.sm2 in ("gs", "gsb"):
            x = self.smooth(lev, None, b, self.nu_pre, backward=False)
        else:
            x = self.smooth(lev, None, b, self.nu_pre)
        r = b - lev["A"] @ x
        P = lev["P"]; nxt = self.levels[l + 1]
        bc = (P.T @ (r*lev["free"]))*nxt["free"]
        xc = self.vcycle(bc, l + 1)
        x = x + (P @ (xc*nxt["free"]))*lev["free"]
        if self.sm2 in ("gs", "gsb"):
            x = self.smooth(lev, x, b, self.nu_post, backward=True)
        else:
            x = self.smooth(lev, x, b, self.nu_post)
        return x

if __name__ == "__main__":
    nx, ny = 256, 128
    d = np.load(os.path.join(OUT, "design_256x128_100.npz"))
    ke, scale, fixed, f = setup(nx, ny, d["rho_cur"])
    _, scale_prev, _, _ = setup(nx, ny, d["rho_prev"])
    free = (~fixed).astype(float)
    def Aof(sc):
        K = G.full_K(nx, ny, sc, ke); M = sp.diags(free)
        return (M @ K @ M + sp.diags(1 - free)).tocsr()
    A, Ap = Aof(scale), Aof(scale_prev)
    fb = f*free
    u_prev = spla.spsolve(Ap.tocsc(), fb)
    for name, kw in [("jacobi 0.7", dict(smoother="jacobi", omega=0.7)),
                     ("jacobi 0.7 V(2,2)", dict(smoother="jacobi", omega=0.7, nu_pre=2, nu_post=2)),
                     ("cheby2", dict(smoother="cheby", cheb_deg=2)),
                     ("cheby3", dict(smoother="cheby", cheb_deg=3)),
                     ("l1-jacobi", dict(smoother="l1")),
                     ("gs 4-colour", dict(smoother="gs")),
                     ("gs node-block", dict(smoother="gsb")),
                     ("gs fine+L1 only", dict(smoother="gs", gs_levels=2)),
                     ("gs V(2,2)", dict(smoother="gs", nu_pre=2, nu_post=2)),
                     ]:
        t0 = time.time()
        mg = GMG2(nx, ny, scale, ke, fixed, min_coarse=4, **kw)
        _, it0, h0 = G.pcg(A, fb, mg.vcycle, rtol=1e-9, maxiter=400)
        _, it1, h1 = G.pcg(A, fb, mg.vcycle, x0=u_prev, rtol=1e-9, maxiter=400)
        print("%-22s cold %3d  warm %3d   (%.1fs)" % (name, it0, it1, time.time() - t0), flush=True)
