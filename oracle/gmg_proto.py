"""Numpy/scipy prototype of the Galerkin geometric-multigrid PCG the device solver implements.

ORACLE / TEST INFRASTRUCTURE ONLY (algorithm exploration + a CPU cross-check of the device solver's
iteration counts).  Not part of the reference: the reference solves with SuperLU (optimize.py:437);
this is the replacement's algorithm restated with sparse matrices so it can be studied on the CPU.

Full-DOF-space formulation: vectors have 2*(nx+1)*(ny+1) entries; constrained DOFs are handled by a
0/1 mask (A = M K M + (I - M)).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def full_K(nx, ny, scale, ke):
    """Unconstrained K over all 2*N_n DOFs from per-element scale (N_e,) and ke (8,8)."""
    e = np.arange(nx*ny)
    n = e + e//nx
    con = np.stack([n, n + 1, n + nx + 2, n + nx + 1], 1)
    dof = np.stack([2*con, 2*con + 1], 2).reshape(-1, 8)
    rows = np.repeat(dof, 8, 1).ravel()
    cols = np.tile(dof, (1, 8)).ravel()
    vals = (scale[:, None]*ke.ravel()[None]).ravel()
    nd = 2*(nx + 1)*(ny + 1)
    return sp.coo_matrix((vals, (rows, cols)), shape=(nd, nd)).tocsr()


def prolongation(nxc, nyc):
    """Bilinear interpolation from a (nxc x nyc)-element grid to the (2nxc x 2nyc) grid, per DOF."""
    def p1(nc):
        nf = 2*nc
        P = sp.lil_matrix((nf + 1, nc + 1))
        for i in range(nc + 1):
            P[2*i, i] = 1.0
        for i in range(nc):
            P[2*i + 1, i] = 0.5
            P[2*i + 1, i + 1] = 0.5
        return P.tocsr()
    Pn = sp.kron(p1(nyc), p1(nxc))          # node ordering: row-major, x fastest
    return sp.kron(Pn, sp.identity(2)).tocsr()


class GMG:
    def __init__(self, nx, ny, scale, ke, fixed, nu_pre=1, nu_post=1, omega=0.7, min_coarse=2,
                 smoother="jacobi", cheb_deg=2):
        """fixed: bool (2*N_n,) mask of constrained DOFs on the fine grid."""
        self.omega, self.nu_pre, self.nu_post = omega, nu_pre, nu_post
        self.smoother, self.cheb_deg = smoother, cheb_deg
        self.levels = []
        K = full_K(nx, ny, scale, ke)
        free = (~fixed).astype(float)
        lx, ly = nx, ny
        while True:
            M = sp.diags(free)
            A = (M @ K @ M + sp.diags(1.0 - free)).tocsr()
            lev = {"A": A, "free": free, "nx": lx, "ny": ly, "dinv": 1.0/A.diagonal()}
            self.levels.append(lev)
            if lx % 2 or ly % 2 or min(lx, ly)//2 < min_coarse:
                break
            P = prolongation(lx//2, ly//2)
            lev["P"] = P
            K = (P.T @ K @ P).tocsr()         # Galerkin on the UNCONSTRAINED operator
            fn = free.reshape(ly + 1, lx + 1, 2)
            free = fn[::2, ::2, :].reshape(-1).copy()   # injection of the constraint mask
            lx, ly = lx//2, ly//2
        Ac = self.levels[-1]["A"]
        self.coarse = spla.splu(Ac.tocsc())
        if smoother == "cheby":
            for lev in self.levels[:-1]:
                # eigenvalue bound of D^-1 A by a few power iterations
                A, dinv = lev["A"], lev["dinv"]
                v = np.random.default_rng(0).standard_normal(A.shape[0])
                for _ in range(15):
                    v = dinv*(A @ v)
                    lam = np.linalg.norm(v)
                    v /= lam
                lev["lmax"] = 1.1*lam

    def smooth(self, lev, x, b, n):
        A, dinv = lev["A"], lev["dinv"]
        if self.smoother == "jacobi":
            for _ in range(n):
                x = x + self.omega*dinv*(b - A @ x) if x is not None else self.omega*dinv*b
            return x
        # Chebyshev on [lmax/a, lmax]
        lmax = lev["lmax"]
        lmin = lmax/6.0
        theta, delta = 0.5*(lmax + lmin), 0.5*(lmax - lmin)
        for _ in range(n):
            r = b - A @ x if x is not None else b.copy()
            if x is None:
                x = np.zeros_like(b)
            sigma = theta/delta
            rho = 1.0/sigma
            d = dinv*r/theta
            for k in range(self.cheb_deg):
                x = x + d
                if k + 1 < self.cheb_deg:
                    r = r - A @ d
                    rho_new = 1.0/(2*sigma - rho)
                    d = rho_new*rho*d + 2*rho_new/delta*(dinv*r)
                    rho = rho_new
        return x

    def vcycle(self, b, l=0):
        lev = self.levels[l]
        if l == len(self.levels) - 1:
            return self.coarse.solve(b)
        x = self.smooth(lev, None, b, self.nu_pre)
        r = b - lev["A"] @ x
        P = lev["P"]
        nxt = self.levels[l + 1]
        bc = (P.T @ (r*lev["free"]))*nxt["free"]
        xc = self.vcycle(bc, l + 1)
        x = x + (P @ (xc*nxt["free"]))*lev["free"]
        x = self.smooth(lev, x, b, self.nu_post)
        return x


def pcg(A, b, precond, x0=None, rtol=1e-10, maxiter=1000):
    x = np.zeros_like(b) if x0 is None else x0.copy()
    r = b - A @ x
    bn = np.linalg.norm(b)
    z = precond(r)
    p = z.copy()
    rz = r @ z
    hist = [np.linalg.norm(r)/bn]
    it = 0
    while hist[-1] > rtol and it < maxiter:
        Ap = A @ p
        alpha = rz/(p @ Ap)
        x += alpha*p
        r -= alpha*Ap
        z = precond(r)
        rz_new = r @ z
        p = z + (rz_new/rz)*p
        rz = rz_new
        it += 1
        hist.append(np.linalg.norm(r)/bn)
    return x, it, hist
