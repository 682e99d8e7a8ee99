"""Row-slab decomposition of stages 1-2 across GPUs -- first building block.

One process per GPU (``torch.distributed``, NCCL over NVLink).  Because the reference numbering is row-major
(``rect_grid``), a slab of element rows is a contiguous range of element ids and of node ids.  Rank ``g`` owns
element rows ``[e0, e1)`` as an ORDINARY handle of ``libtopo_b200`` on an ``nx x (e1-e0)`` mesh; the node rows on a
slab interface exist on both neighbours.  A local stiffness application therefore yields PARTIAL sums on the
interface rows (local elements only); the halo exchange adds the neighbour's partials (one node row =
``2*(nx+1)`` doubles each way per matvec, e.g. 131 KB at nx = 8192).  Scalars (dot products) are all-reduced;
each node row is counted by exactly one rank (the lower slab owns its top interface row).

What is distributed so far: K(rho)u with halo sum, the Jacobi diagonal, and a Jacobi-preconditioned CG on top of
them (host loop in Python, every vector kernel is the library's own, NCCL only moves halo rows and scalars).  The
multigrid preconditioner, filter and OC stages are single-GPU only.

The numerical kernels are injected through a small backend object so that the slab / halo / ownership logic is
testable on CPU with the ``gloo`` backend (``tests/test_distributed_cpu.py``) -- the product backend is
``TopoBackend`` (CUDA kernels through the C ABI; no CPU fallback).
"""
from __future__ import annotations

import numpy as np


def slab_rows(ny, world, rank):
    """Balanced contiguous partition of the ``ny`` element rows: rank -> [e0, e1)."""
    if world > ny:
        raise ValueError("more ranks (%d) than element rows (%d)" % (world, ny))
    base, rem = divmod(ny, world)
    e0 = rank*base + min(rank, rem)
    return e0, e0 + base + (1 if rank < rem else 0)


class TopoBackend:
    """Local kernels = libtopo_b200 on this rank's slab; vectors are torch CUDA tensors (device memory only)."""

    def __init__(self, nx, ny_local, ke, cons_local, scale_local, device):
        import torch
        from solidspy_opt import _capi
        self.torch, self._capi = torch, _capi
        self.nx, self.ny = nx, ny_local
        self.topo = _capi.Topo(nx, ny_local, ke, device=device)
        self.topo.set_stream(torch.cuda.current_stream().cuda_stream)     # NCCL ops are ordered against this stream
        self.topo.set_cons(cons_local)
        self.topo.set(_capi.F_SCALE, scale_local)
        self.device = torch.device("cuda", device)
        self.nn = (nx + 1)*(ny_local + 1)

    def empty(self):
        return self.torch.empty((self.ny + 1, self.nx + 1, 2), dtype=self.torch.float64, device=self.device)

    def zeros(self):
        return self.torch.zeros((self.ny + 1, self.nx + 1, 2), dtype=self.torch.float64, device=self.device)

    def from_numpy(self, a):
        return self.torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64).reshape(self.ny + 1, self.nx + 1, 2)).to(self.device)

    def scalar(self):
        return self.torch.zeros(1, dtype=self.torch.float64, device=self.device)

    def matvec(self, u, y):
        self.topo.matvec_dev(u.data_ptr(), y.data_ptr())

    def diag(self, d):
        self.topo.diag_dev(d.data_ptr())

    def axpby(self, alpha, a, beta, b, c):
        self.topo.axpby_dev(a.numel()//2, alpha, a.data_ptr(), beta, b.data_ptr(), c.data_ptr())

    def mul(self, a, b, c, divide=False):
        self.topo.mul_dev(a.numel()//2, a.data_ptr(), b.data_ptr(), c.data_ptr(), divide)

    def dot(self, a, b, row_lo, row_hi, out):
        self.topo.dot_dev(a.data_ptr(), b.data_ptr(), row_lo, row_hi, out.data_ptr())


class SlabSolver:
    """K(scale) u and Jacobi-PCG over row slabs.  ``backend`` supplies the local kernels, ``dist`` is
    ``torch.distributed`` (or None for a single rank)."""

    def __init__(self, backend, rank, world, dist=None):
        self.b, self.rank, self.world, self.dist = backend, rank, world, dist
        self.has_lo, self.has_hi = rank > 0, rank < world - 1
        n_rows = backend.ny + 1
        # ownership: the lower slab owns the shared interface row, so every rank but the first skips its first row
        self.own_lo, self.own_hi = (1 if self.has_lo else 0), n_rows
        self._lo_buf = self._hi_buf = None
        self.halo_bytes_per_exchange = 0

    # ---- communication -----------------------------------------------------------------------------------
    def halo_sum(self, v):
        """Add the neighbours' partial sums on the interface node rows of ``v`` (rows, nx+1, 2)."""
        if self.world == 1:
            return
        dist = self.dist
        ops, todo = [], []
        if self.has_lo:
            if self._lo_buf is None:
                self._lo_buf = v[0].clone()
            ops += [dist.P2POp(dist.isend, v[0], self.rank - 1), dist.P2POp(dist.irecv, self._lo_buf, self.rank - 1)]
            todo.append((v[0], self._lo_buf))
        if self.has_hi:
            if self._hi_buf is None:
                self._hi_buf = v[-1].clone()
            ops += [dist.P2POp(dist.isend, v[-1], self.rank + 1), dist.P2POp(dist.irecv, self._hi_buf, self.rank + 1)]
            todo.append((v[-1], self._hi_buf))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        for row, buf in todo:
            self.b.axpby(1.0, row, 1.0, buf, row)            # a + b == b + a: both copies stay bit-identical
        self.halo_bytes_per_exchange = sum(buf.numel()*8 for _, buf in todo)

    def allreduce(self, t):
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t

    # ---- operators ------------------------------------------------------------------------------------------
    def matvec(self, u, y):
        """y = K u on free DOFs, complete on every row of the slab (interface rows summed)."""
        self.b.matvec(u, y)
        self.halo_sum(y)

    def dot(self, a, b, out=None):
        out = self.b.scalar() if out is None else out
        self.b.dot(a, b, self.own_lo, self.own_hi, out)
        return self.allreduce(out)

    def jacobi_diag_inverse(self):
        d = self.b.empty()
        self.b.diag(d)
        self.halo_sum(d)
        ones = self.b.empty()
        ones.fill_(1.0)
        self.b.mul(ones, d, d, divide=True)              # 1/diag, 0 at constrained DOFs
        return d

    def solve_jacobi_pcg(self, f, rtol=1e-8, maxit=100000, x0=None):
        """Jacobi-preconditioned CG on K u = f (f, u: slab tensors, constrained DOFs zero).
        Returns (u, iterations, relative residual)."""
        b = self.b
        dinv = self.jacobi_diag_inverse()
        u = b.zeros() if x0 is None else x0.clone()
        r, z, p, Ap = b.empty(), b.empty(), b.empty(), b.empty()
        if x0 is None:
            r.copy_(f)
        else:
            self.matvec(u, Ap)
            b.axpby(1.0, f, -1.0, Ap, r)
        b.mul(r, dinv, z)
        p.copy_(z)
        bb = float(self.dot(f, f).item())
        rz = float(self.dot(r, z).item())
        rr = float(self.dot(r, r).item())
        it = 0
        two = b.scalar().repeat(2)
        while rr > (rtol*rtol)*bb and it < maxit:
            self.matvec(p, Ap)
            pAp = float(self.dot(p, Ap).item())
            alpha = rz/pAp
            b.axpby(1.0, u, alpha, p, u)
            b.axpby(1.0, r, -alpha, Ap, r)
            b.mul(r, dinv, z)
            b.dot(r, z, self.own_lo, self.own_hi, two[0:1])
            b.dot(r, r, self.own_lo, self.own_hi, two[1:2])
            self.allreduce(two)                          # both scalars in one all-reduce
            rz_new, rr = (float(v) for v in two.tolist())
            b.axpby(1.0, z, rz_new/rz, p, p)
            rz = rz_new
            it += 1
        return u, it, (rr/bb)**0.5 if bb > 0 else 0.0
