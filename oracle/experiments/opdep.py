"""CPU study behind DESIGN.md section 4 "Late in an optimisation" (ORACLE-SIDE EXPERIMENT, numpy/scipy only; nothing in the
product imports this).  Stiffness-weighted (operator-dependent) interpolation with Galerkin coarsening on the same design.
Run from the repo root:  python oracle/experiments/opdep.py   (after gen_design.py 256 128 100)"""
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
from smoothers import setup

def edge_weights(nx, ny, s):
    """s: (ny, nx) element 'stiffness' of this level. returns horizontal fine-edge stiffness kh (ny+1, nx) and vertical kv (ny, nx+1)"""
    sp_ = np.zeros((ny + 2, nx + 2)); sp_[1:-1, 1:-1] = s
    kh = sp_[:-1, 1:-1] + sp_[1:, 1:-1]      # edge between node (j,i) and (j,i+1): elements below (j-1,i) and above (j,i)
    kv = sp_[1:-1, :-1] + sp_[1:-1, 1:]      # edge between node (j,i) and (j+1,i): elements left (j,i-1), right (j,i)
    return kh, kv

def prolongation_opdep(nxc, nyc, s_fine, power=1.0):
    nx, ny = 2*nxc, 2*nyc
    kh, kv = edge_weights(nx, ny, s_fine**power)
    nf, nc = (nx + 1)*(ny + 1), (nxc + 1)*(nyc + 1)
    rows, cols, vals = [], [], []
    fid = lambda j, i: j*(nx + 1) + i
    cid = lambda J, I: J*(nxc + 1) + I
    W = {}
    # coarse points
    for J in range(nyc + 1):
        for I in range(nxc + 1):
            W[(2*J, 2*I)] = {cid(J, I): 1.0}
    # horizontal edge midpoints (even row, odd col)
    for J in range(nyc + 1):
        for I in range(nxc):
            j, i = 2*J, 2*I + 1
            kl, kr = kh[j, i - 1], kh[j, i]
            W[(j, i)] = {cid(J, I): kl/(kl + kr), cid(J, I + 1): kr/(kl + kr)}
    for J in range(nyc):
        for I in range(nxc + 1):
            j, i = 2*J + 1, 2*I
            kd, ku = kv[j - 1, i], kv[j, i]
            W[(j, i)] = {cid(J, I): kd/(kd + ku), cid(J + 1, I): ku/(kd + ku)}
    for J in range(nyc):
        for I in range(nxc):
            j, i = 2*J + 1, 2*I + 1
            kw, ke_, ks, kn = kh[j, i - 1], kh[j, i], kv[j - 1, i], kv[j, i]
            tot = kw + ke_ + ks + kn
            acc = {}
            for wgt, nb in ((kw, (j, i - 1)), (ke_, (j, i + 1)), (ks, (j - 1, i)), (kn, (j + 1, i))):
                for c, w in W[nb].items():
                    acc[c] = acc.get(c, 0.0) + wgt/tot*w
            W[(j, i)] = acc
    for (j, i), d in W.items():
        for c, w in d.items():
            rows.append(fid(j, i)); cols.append(c); vals.append(w)
    Pn = sp.coo_matrix((vals, (rows, cols)), shape=(nf, nc)).tocsr()
    return sp.kron(Pn, sp.identity(2)).tocsr()

class GMGop(G.GMG):
    def __init__(self, nx, ny, scale, ke, fixed, omega=0.7, min_coarse=4, power=1.0, opdep_levels=99, nu=1):
        self.omega, self.nu_pre, self.nu_post = omega, nu, nu
        self.smoother = "jacobi"
        self.levels = []
        K = G.full_K(nx, ny, scale, ke)
        free = (~fixed).astype(float)
        lx, ly = nx, ny
        s = scale.reshape(ny, nx).copy()
        li = 0
        while True:
            M = sp.diags(free)
            A = (M @ K @ M + sp.diags(1.0 - free)).tocsr()
            lev = {"A": A, "free": free, "nx": lx, "ny": ly, "dinv": 1.0/A.diagonal()}
            self.levels.append(lev)
            if lx % 2 or ly % 2 or min(lx, ly)//2 < min_coarse:
                break
            P = prolongation_opdep(lx//2, ly//2, s, power) if li < opdep_levels else G.prolongation(lx//2, ly//2)
            lev["P"] = P
            K = (P.T @ K @ P).tocsr()
            fn = free.reshape(ly + 1, lx + 1, 2)
            free = fn[::2, ::2, :].reshape(-1).copy()
            # coarse 'stiffness' per cell: mean of children (arithmetic)
            s = 0.25*(s[0::2, 0::2] + s[1::2, 0::2] + s[0::2, 1::2] + s[1::2, 1::2])
            lx, ly = lx//2, ly//2
            li += 1
        self.coarse = spla.splu(self.levels[-1]["A"].tocsc())

if __name__ == "__main__":
    nx, ny = 256, 128
    d = np.load(os.path.join(OUT, "design_256x128_100.npz"))
    ke, scale, fixed, f = setup(nx, ny, d["rho_cur"])
    free = (~fixed).astype(float)
    K = G.full_K(nx, ny, scale, ke); M = sp.diags(free)
    A = (M @ K @ M + sp.diags(1 - free)).tocsr()
    fb = f*free
    for name, kw in [("bilinear", dict(opdep_levels=0)), ("opdep p=1", dict()), ("opdep p=0.5", dict(power=0.5)), ("opdep p=1 L0 only", dict(opdep_levels=1)),
                     ("opdep p=1 L0-1", dict(opdep_levels=2)), ("opdep p=0.25", dict(power=0.25))]:
        t0 = time.time()
        mg = GMGop(nx, ny, scale, ke, fixed, **kw)
        _, it0, h0 = G.pcg(A, fb, mg.vcycle, rtol=1e-9, maxiter=400)
        print("%-22s cold %3d   (%.1fs)" % (name, it0, time.time() - t0), flush=True)
