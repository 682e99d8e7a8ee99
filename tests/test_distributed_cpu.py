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
