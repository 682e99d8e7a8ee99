"""Slab / halo / ownership logic of solidspy_opt.distributed on CPU: world_size-2 and -3 `gloo` process groups with a
numpy stand-in for the local kernels (the product backend is CUDA-only; this covers the host-side N>1 logic)."""
import os
import socket
import sys

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (os.path.join(ROOT, "solidspy-opt_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


class NumpyBackend:
    """Local kernels restated with scipy on CPU tensors (test double of distributed.TopoBackend)."""

    def __init__(self, nx, ny_local, ke, cons_local, scale_local):
        import gmg_proto
        self.nx, self.ny = nx, ny_local
        K = gmg_proto.full_K(nx, ny_local, scale_local, ke)
        free = (cons_local.reshape(-1) == 0).astype(float)
        M = sp.diags(free)
        self.A = (M @ K @ M).tocsr()
        self.free = free

    def empty(self):
        return torch.zeros((self.ny + 1, self.nx + 1, 2), dtype=torch.float64)

    zeros = empty

    def from_numpy(self, a):
        return torch.as_tensor(np.array(a, dtype=np.float64).reshape(self.ny + 1, self.nx + 1, 2))

    def scalar(self):
        return torch.zeros(1, dtype=torch.float64)

    def matvec(self, u, y):
        y.copy_(torch.as_tensor(self.A @ u.numpy().reshape(-1)).reshape(y.shape))

    def diag(self, d):
        d.copy_(torch.as_tensor(self.A.diagonal()*self.free).reshape(d.shape))

    def axpby(self, alpha, a, beta, b, c):
        c.copy_(alpha*a + beta*b)

    def mul(self, a, b, c, divide=False):
        c.copy_(torch.where(b != 0, a/b, torch.zeros_like(a)) if divide else a*b)

    def dot(self, a, b, row_lo, row_hi, out):
        out.copy_((a[row_lo:row_hi]*b[row_lo:row_hi]).sum().reshape(1))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nx, ny, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import gmg_proto
        import t1
        from solidspy_opt.distributed import SlabSolver, slab_rows
        rng = np.random.default_rng(7)                       # same stream on every rank: same global problem
        ke = t1.kloc_closed_form(3.0, 0.3)
        scale = 1e-3 + rng.uniform(0, 1, nx*ny)**3
        cons = np.zeros(((nx + 1)*(ny + 1), 2), dtype=np.int8)
        cons[::nx + 1] = -1
        cons[rng.choice(cons.shape[0], 5, replace=False), 1] = -1
        u = rng.standard_normal((ny + 1, nx + 1, 2))
        f = np.zeros((ny + 1, nx + 1, 2))
        f[ny//2, nx, 1] = -1.0
        f[ny, nx//2, 0] = 0.5
        free = (cons.reshape(-1) == 0).astype(float)
        f = (f.reshape(-1)*free).reshape(f.shape)
        # global reference
        K = gmg_proto.full_K(nx, ny, scale, ke)
        M = sp.diags(free)
        A = (M @ K @ M).tocsr()
        y_ref = (A @ (u.reshape(-1)*free)).reshape(u.shape)
        idx = np.flatnonzero(free)
        from scipy.sparse.linalg import spsolve
        u_ref = np.zeros(free.size)
        u_ref[idx] = spsolve(A[idx][:, idx].tocsc(), f.reshape(-1)[idx])
        u_ref = u_ref.reshape(u.shape)
        # this rank's slab
        e0, e1 = slab_rows(ny, world, rank)
        be = NumpyBackend(nx, e1 - e0, ke, cons.reshape(ny + 1, nx + 1, 2)[e0:e1 + 1].reshape(-1, 2),
                          scale.reshape(ny, nx)[e0:e1].reshape(-1))
        S = SlabSolver(be, rank, world, dist)
        ul = be.from_numpy((u.reshape(-1)*free).reshape(u.shape)[e0:e1 + 1])
        yl = be.empty()
        S.matvec(ul, yl)
        err_mv = np.abs(yl.numpy() - y_ref[e0:e1 + 1]).max()/np.abs(y_ref).max()
        dot = float(S.dot(ul, ul).item())
        err_dot = abs(dot - float(((u.reshape(-1)*free)**2).sum()))/dot
        us, it, rel = S.solve_jacobi_pcg(be.from_numpy(f[e0:e1 + 1]), rtol=1e-11, maxit=20000)
        err_u = np.abs(us.numpy() - u_ref[e0:e1 + 1]).max()/np.abs(u_ref).max()
        q.put((rank, (e0, e1), err_mv, err_dot, err_u, it, rel))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nx,ny", [(2, 12, 9), (3, 7, 10)])
def test_slab_matvec_dot_and_pcg_match_global(world, nx, ny):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nx, ny, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rows = sorted(r[1] for r in res)
    assert rows[0][0] == 0 and rows[-1][1] == ny and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
    for rank, _, err_mv, err_dot, err_u, it, rel in res:
        assert err_mv < 1e-13, (rank, err_mv)          # halo sum of the interface rows
        assert err_dot < 1e-13, (rank, err_dot)        # every node row counted exactly once
        assert err_u < 1e-8 and rel <= 1e-11, (rank, err_u, rel, it)
    assert len({r[5] for r in res}) == 1               # all ranks did the same number of iterations


def test_slab_rows_partition():
    from solidspy_opt.distributed import slab_rows
    for ny, world in [(1024, 8), (10, 3), (7, 7), (4096, 8), (5, 2)]:
        parts = [slab_rows(ny, world, r) for r in range(world)]
        assert parts[0][0] == 0 and parts[-1][1] == ny
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        slab_rows(3, 4, 0)


# ======================================================================================================================
# The row-slab multigrid-PCG ALGORITHM of solidspy_opt.distributed.SlabSimp (accumulated / distributed vectors,
# weighted interface rows, restriction without exchange, gathered level K, completion of locally smoothed interface
# rows) restated in numpy and executed over gloo: same segment order and the same movements as the CUDA path
# (csrc/dist.cu header), checked against the single-domain multigrid-PCG on the same hierarchy.
# ======================================================================================================================
def _halo_sum(v, rank, world):
    """v: (rows, npr, 2) torch tensor; add the neighbours' interface rows (both copies end up identical)."""
    ops, todo = [], []
    if rank > 0:
        lo = torch.empty_like(v[0])
        ops += [dist.P2POp(dist.isend, v[0].clone(), rank - 1), dist.P2POp(dist.irecv, lo, rank - 1)]
        todo.append((0, lo))
    if rank < world - 1:
        hi = torch.empty_like(v[-1])
        ops += [dist.P2POp(dist.isend, v[-1].clone(), rank + 1), dist.P2POp(dist.irecv, hi, rank + 1)]
        todo.append((-1, hi))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for row, buf in todo:
        v[row] += buf
    return todo


def _slab_gmg_worker(rank, world, port, nx, ny, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import gmg_proto as G
        import t1
        rng = np.random.default_rng(11)
        ke = t1.kloc_closed_form(3.0, 0.3)
        scale = 1e-6 + rng.uniform(0, 1, nx*ny)**3
        cons = np.zeros((ny + 1, nx + 1, 2), dtype=np.int8)
        cons[:, 0, :] = -1
        f = np.zeros((ny + 1, nx + 1, 2))
        nyl = ny//world
        f[nyl, nx, 1] = -1.0                                  # on the first interface row
        f[ny, nx//2, 0] = 0.5
        om = 0.7

        def T(a, rows, npr):                                  # flat numpy vector <-> (rows, npr, 2) torch view
            return torch.from_numpy(a).reshape(rows, npr, 2)

        # ---- single-domain reference: levels 0, 1 (V(1,1)), level 2 solved exactly ----------------------------------
        free = (cons.reshape(-1) == 0).astype(float)
        A0 = (sp.diags(free) @ G.full_K(nx, ny, scale, ke) @ sp.diags(free)).tocsr()
        P0, P1 = G.prolongation(nx//2, ny//2), G.prolongation(nx//4, ny//4)
        A1 = (P0.T @ A0 @ P0).tocsr()
        A2 = (P1.T @ A1 @ P1).toarray()
        dead2 = np.diag(A2) <= 0
        A2[dead2] = 0; A2[:, dead2] = 0; A2[dead2, dead2] = 1
        A2inv = np.linalg.inv(A2)

        def inv_diag(A):
            d = A.diagonal()
            return np.where(d > 0, 1.0/np.where(d > 0, d, 1.0), 0.0)
        d0, d1 = inv_diag(A0), inv_diag(A1)

        def vcycle_ref(r):
            x0 = om*d0*r
            r0 = (r - A0 @ x0)*(d0 != 0)
            b1 = P0.T @ r0
            x1 = om*d1*b1
            r1 = (b1 - A1 @ x1)*(d1 != 0)
            x2 = A2inv @ (P1.T @ r1)
            x2[dead2] = 0
            x1 = x1 + (P1 @ x2)*(d1 != 0)
            x1 = x1 + om*d1*(b1 - A1 @ x1)
            x0 = x0 + (P0 @ x1)*(d0 != 0)
            return x0 + om*d0*(r - A0 @ x0)
        fr = f.reshape(-1)*free
        u_ref, it_ref, _ = G.pcg((A0 + sp.diags(1 - free)).tocsr(), fr, lambda r: vcycle_ref(r*free), rtol=1e-10, maxiter=500)

        # ---- this rank's slab ------------------------------------------------------------------------------------------
        e0 = rank*nyl
        npr0, npr1, npr2 = nx + 1, nx//2 + 1, nx//4 + 1
        rows0, rows1, rows2 = nyl + 1, nyl//2 + 1, nyl//4 + 1
        cl = cons[e0:e0 + nyl + 1].reshape(-1)
        freel = (cl == 0).astype(float)
        A0l = (sp.diags(freel) @ G.full_K(nx, nyl, scale.reshape(ny, nx)[e0:e0 + nyl].reshape(-1), ke) @ sp.diags(freel)).tocsr()
        P0l, P1l = G.prolongation(nx//2, nyl//2), G.prolongation(nx//4, nyl//4)
        A1l = (P0l.T @ A0l @ P0l).tocsr()                    # slab-local Galerkin product: a DISTRIBUTED operator
        A2l = (P1l.T @ A1l @ P1l).toarray()                  # this slab's piece of the global level-2 operator
        has_lo, has_hi = rank > 0, rank < world - 1

        def W(rows, npr):                                     # 1/2 on interface rows: accumulated -> distributed
            w = np.ones((rows, npr, 2))
            if has_lo: w[0] = 0.5
            if has_hi: w[-1] = 0.5
            return w.reshape(-1)
        W0, W1 = W(rows0, npr0), W(rows1, npr1)

        def acc_diag(A, rows, npr):                           # interface diagonals are summed once per set-up
            d = A.diagonal().copy()
            _halo_sum(T(d, rows, npr), rank, world)
            return np.where(d > 0, 1.0/np.where(d > 0, d, 1.0), 0.0)
        d0l, d1l = acc_diag(A0l, rows0, npr0), acc_diag(A1l, rows1, npr1)
        # level 2 is global: gather the slab pieces of the operator (once) ...
        pieces = [None]*world
        dist.all_gather_object(pieces, A2l)
        n2 = (ny//4 + 1)*npr2*2
        A2g = np.zeros((n2, n2))
        for g, Ag in enumerate(pieces):
            off = g*(nyl//4)*npr2*2
            A2g[off:off + Ag.shape[0], off:off + Ag.shape[0]] += Ag
        dg = np.diag(A2g) <= 0
        A2g[dg] = 0; A2g[:, dg] = 0; A2g[dg, dg] = 1
        A2ginv = np.linalg.inv(A2g)

        def allsum(x):
            t = torch.tensor([x], dtype=torch.float64)
            dist.all_reduce(t)
            return float(t[0])

        def complete(x2, base, rows, npr):                    # x2 <- x2 + x2(neighbour) - base on interface rows
            for row, _ in _halo_sum(T(x2, rows, npr), rank, world):
                T(x2, rows, npr)[row] -= T(base, rows, npr)[row]

        def vcycle_slab(r):                                   # r accumulated; returns z accumulated, and r.z share
            xh = om*d0l*r                                     # pcg_update: x^ = w D^-1 r'
            res = (W0*r - A0l @ xh)*(d0l != 0)                #   res = W r' - K_loc x^  (distributed)
            b1 = P0l.T @ res                                  #   restriction: no exchange
            _halo_sum(T(b1, rows1, npr1), rank, world)        # exchange TOPO_X_LEVEL_B
            x1 = om*d1l*b1                                    # vcycle_down(1)
            r1 = (W1*b1 - A1l @ x1)*(d1l != 0)
            b2 = P1l.T @ r1                                   #   piece of the level-2 right-hand side
            allb = [torch.empty(b2.size, dtype=torch.float64) for _ in range(world)]
            dist.all_gather(allb, torch.from_numpy(b2))       # gather TOPO_G_BK
            b2g = np.zeros(n2)
            for g, bg in enumerate(allb):
                off = g*(nyl//4)*npr2*2
                b2g[off:off + bg.numel()] += bg.numpy()
            x2g = A2ginv @ b2g                                # vcycle_tail: every rank, redundantly
            x2g[dg] = 0
            off = rank*(nyl//4)*npr2*2
            x2 = x2g[off:off + b2.size]
            x1 = x1 + (P1l @ x2)*(d1l != 0)                   # vcycle_up(1): prolongation needs no exchange
            x1n = x1 + om*d1l*(W1*b1 - A1l @ x1)
            complete(x1n, x1, rows1, npr1)                    # exchange TOPO_X_LEVEL_X2
            xt = xh + (P0l @ x1n)*(d0l != 0)                  # pcg_smooth
            z = xt + om*d0l*(W0*r - A0l @ xt)
            rz = float(r @ (z - (1 - W0)*xt))                 #   r.z share (before the completion of z)
            complete(z, xt, rows0, npr0)                      # exchange TOPO_X_Z
            return z, rz

        fl = (f[e0:e0 + nyl + 1].reshape(-1)*freel)           # accumulated right-hand side
        u = np.zeros_like(fl)
        r = fl.copy()
        bb = allsum(float((W0*fl) @ fl))
        z, rz_part = vcycle_slab(r)
        rz = allsum(rz_part)
        p = z.copy()
        it = 0
        rr = allsum(float((W0*r) @ r))
        while rr > (1e-10)**2*bb and it < 500:
            Ap = A0l @ p                                      # pcg_matvec: distributed
            pAp = allsum(float(p @ Ap))                       #   p.Ap needs no weights
            _halo_sum(T(Ap, rows0, npr0), rank, world)        # exchange TOPO_X_AP
            alpha = rz/pAp
            u += alpha*p
            r -= alpha*Ap
            r *= (d0l != 0)
            rr = allsum(float((W0*r) @ r))
            z, rz_part = vcycle_slab(r)
            rz_new = allsum(rz_part)
            p = z + (rz_new/rz)*p
            rz = rz_new
            it += 1
        u_ref_l = u_ref.reshape(ny + 1, nx + 1, 2)[e0:e0 + nyl + 1].reshape(-1)
        q.put((rank, it, it_ref, float(np.abs(u - u_ref_l).max()/np.abs(u_ref).max()), (rr/bb)**0.5))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nx,ny", [(2, 16, 16), (4, 12, 32)])
def test_row_slab_multigrid_pcg_algorithm_over_gloo(world, nx, ny):
    """Iteration counts and solution of the slab algorithm equal the single-domain multigrid-PCG (same hierarchy)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_slab_gmg_worker, args=(r, world, port, nx, ny, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, it, it_ref, err, rel in res:
        assert it == it_ref, (rank, it, it_ref)
        assert err < 1e-8 and rel <= 1e-10, (rank, err, rel)
