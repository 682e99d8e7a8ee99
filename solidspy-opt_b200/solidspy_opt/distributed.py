"""Row-slab decomposition of stages 1-2 across GPUs -- first building block.

One process per GPU (``torch.distributed``, NCCL over NVLink).  Because the reference numbering is row-major
(``rect_grid``), a slab of element rows is a contiguous range of element ids and of node ids.  Rank ``g`` owns
element rows ``[e0, e1)`` as an ORDINARY handle of ``libtopo_b200`` on an ``nx x (e1-e0)`` mesh; the node rows on a
slab interface exist on both neighbours.  A local stiffness application therefore yields PARTIAL sums on the
interface rows (local elements only); the halo exchange adds the neighbour's partials (one node row =
``2*(nx+1)`` doubles each way per matvec, e.g. 131 KB at nx = 8192).  Scalars (dot products) are all-reduced;
each node row is counted by exactly one rank (the lower slab owns its top interface row).

Two layers live here: ``SlabSolver`` (the first building block: K(rho)u with halo sum, the Jacobi diagonal and a
Jacobi-preconditioned CG on caller-owned tensors, NCCL only moves halo rows and scalars) and ``SlabSimp`` (the WHOLE
design iteration on row slabs: multigrid-preconditioned CG, sensitivity, filter and OC, with peer-memory or NCCL
transport and CUDA-graph capture -- DESIGN.md section 6).

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


# ======================================================================================================================
# The whole SIMP design iteration on row slabs (multigrid-preconditioned CG, sensitivity, filter, OC)
# ======================================================================================================================
class _DevMem:
    """Zero-copy view of library-owned device memory for torch (``__cuda_array_interface__``)."""

    def __init__(self, ptr, n_doubles):
        self.__cuda_array_interface__ = {"shape": (int(n_doubles),), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2}


class SlabSimp:
    """SIMP on an ``nx x ny`` mesh split into ``world`` equal slabs of element rows.

    Two deployments of the SAME host loop:

    * ``dist`` = ``torch.distributed`` (NCCL): this process owns ONE slab on its GPU, ``world`` = number of processes;
      interface rows travel with grouped send/recv, level-K blocks with all-gather, scalars with all-reduce, all on
      ``self.stream`` -- and one pair of PCG iterations (kernels + collectives) is captured in a CUDA graph.
    * ``dist`` = None: this process owns ALL ``world`` slabs on one GPU and moves the rows with device copies --
      the single-GPU test double of the multi-GPU path (same kernels, same segments, same arithmetic).

    The library (``libtopo_b200``) runs the segments; this class only moves bytes between them
    (``include/topo_b200.h``, section "row-slab SIMP iteration").  There is no CPU fallback.
    """

    def __init__(self, nx, ny, ke, cons, loads, volfrac=0.5, penal=3.0, emin=1e-9, emax=1.0, rmin=None, dx=1.0, dy=1.0,
                 move=0.2, rtol=1e-9, maxit=2000, world=None, dist=None, device=0, K=0, use_graph=True,
                 transport="p2p", fused=2):
        import torch
        from solidspy_opt import _capi
        self.torch, self._capi, self.dist = torch, _capi, dist
        if dist is not None:
            self.world, self.rank0, local = dist.get_world_size(), dist.get_rank(), 1
        else:
            self.world, self.rank0, local = int(world), 0, int(world)
        if ny % self.world:
            raise ValueError("ny = %d is not a multiple of the number of slabs %d" % (ny, self.world))
        self.nx, self.ny, self.nyl = int(nx), int(ny), ny//self.world
        self.ne_global = self.nx*self.ny
        self.penal, self.emin, self.emax, self.move = float(penal), float(emin), float(emax), float(move)
        self.rmin = float(rmin) if rmin is not None else None
        self.dx, self.dy, self.rtol, self.maxit, self.volfrac = float(dx), float(dy), float(rtol), int(maxit), float(volfrac)
        self.device = torch.device("cuda", device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.use_graph = bool(use_graph) and dist is not None
        self._graph = None
        self.graph_error = None
        npr = self.nx + 1
        cons = np.asarray(cons).reshape(self.ny + 1, npr, 2)
        loads = np.asarray(loads, dtype=np.float64).reshape(-1, 3)
        self.slabs, self.ranks = [], []
        with torch.cuda.stream(self.stream):
            for i in range(local):
                g = self.rank0 + i
                r0 = g*self.nyl
                t = _capi.Topo(self.nx, self.nyl, ke, device=device)
                t.set_stream(self.stream.cuda_stream)
                t.set_cons(cons[r0:r0 + self.nyl + 1].reshape(-1, 2))
                node = loads[:, 0].astype(np.int64)
                row = node//npr
                sel = (row >= r0) & (row <= r0 + self.nyl)            # interface loads exist on both slabs (accumulated f)
                ll = loads[sel].copy()
                ll[:, 0] = node[sel] - r0*npr
                t.set_loads(ll.reshape(-1, 3))
                self.K, self.n_levels = t.dist_init(g, self.world, K)
                t.fill(_capi.F_RHO, self.volfrac)
                self.slabs.append(t)
                self.ranks.append(g)
        self._views = {}
        self.p2p, self.p2p_error = False, None
        if transport == "p2p":
            self._attach_p2p()
        # one-launch exchanges (push + wait + completion [+ all-reduce + bookkeeping] in ONE kernel): only with one slab per
        # process -- several slabs of one process share a stream, where a kernel that waits for a peer would wait forever
        self.fused = bool(fused) and self.p2p and dist is not None and len(self.slabs) == 1
        # fused == 2 ("inkernel", default with fused=True): the restriction / up kernels of the slab-local levels swap their
        # interface rows themselves -- no exchange launch at all for those levels
        self.inkernel = self.fused and fused is not False and int(fused) != 1
        # fused == 3: additionally the SMALL slab-local levels and the level-K gather run inside one persistent kernel per
        # direction (k_slab_mid_down / k_slab_mid_up); the segment calls of those levels become no-ops inside the library
        self.mid = self.inkernel and int(fused) == 3
        if self.fused:
            self.slabs[0].dist_call("p2p_mode", (1 if self.inkernel else 0) | (2 if self.mid else 0))
        self._lam_prev = 0.0
        self._last_iters = 0
        self.halo_bytes = 0
        self.collectives = 0

    # ---- peer-memory transport -----------------------------------------------------------------------------------------
    def _attach_p2p(self):
        """Map the peers' regions (CUDA IPC between processes).  All ranks agree on success, else everyone stays on
        NCCL for the PCG-loop movements too."""
        ok = 1
        try:
            if self.dist is None:
                ptrs = [t.dist_p2p_export()[1] for t in self.slabs]
                for t in self.slabs:
                    t.dist_p2p_attach(None, ptrs)
            else:
                handle, _ = self.slabs[0].dist_p2p_export()
                handles = [None]*self.world
                self.dist.all_gather_object(handles, handle)
                self.slabs[0].dist_p2p_attach(handles, [0]*self.world)
        except Exception as exc:                                           # e.g. IPC not permitted in this container
            ok, self.p2p_error = 0, repr(exc)
        if self.dist is not None:
            flag = self.torch.tensor([ok], dtype=self.torch.int32, device=self.device)
            self.dist.all_reduce(flag, op=self.dist.ReduceOp.MIN)
            ok = int(flag.item())
        self.p2p = bool(ok)

    _ROW_KINDS = (0, 1, 2, 3, 4)          # X_AP, X_R0, X_Z, X_LEVEL_B, X_LEVEL_X2

    # ---- device views ------------------------------------------------------------------------------------------------
    def _view(self, ptr, n):
        return self.torch.as_tensor(_DevMem(ptr, n), device=self.device)

    def _xchg_views(self, i, what, level):
        key = ("x", i, what, level)
        v = self._views.get(key)
        if v is None or what == self._capi.X_FILTER:
            a, b, c, d, n = self.slabs[i].dist_xchg_ptrs(what, level)
            v = tuple(self._view(p, max(n, 1))[:n] for p in (a, b, c, d))
            self._views[key] = v
        return v

    def _gather_views(self, i, what):
        key = ("g", i, what)
        v = self._views.get(key)
        if v is None:
            a, b, n = self.slabs[i].dist_gather_ptrs(what)
            v = (self._view(a, n), self._view(b, n*self.world))
            self._views[key] = v
        return v

    def _scalar_view(self, i, what):
        key = ("s", i, what)
        v = self._views.get(key)
        if v is None:
            p, n = self.slabs[i].dist_scalar_ptr(what)
            v = self._view(p, n)
            self._views[key] = v
        return v

    # ---- movement ----------------------------------------------------------------------------------------------------
    def exchange(self, what, level=0):
        """Interface rows / blocks to both neighbours, then the library's finishing kernel."""
        if self.fused and what in self._ROW_KINDS:
            self.slabs[0].dist_call("p2p_fused", what, level, 0, 0, 0)
            self.collectives += 1
            return
        if self.p2p and what in self._ROW_KINDS:
            for t in self.slabs:                                           # every push before any wait
                t.dist_call("p2p_xchg", 0, what, level)
            for t in self.slabs:
                t.dist_call("p2p_xchg", 1, what, level)
            self.collectives += 1
            return
        vs = [self._xchg_views(i, what, level) for i in range(len(self.slabs))]
        if vs[0][0].numel() > 0:
            for i in range(len(self.slabs) - 1):                       # neighbours inside this process
                vs[i + 1][2].copy_(vs[i][1])                           # my last row -> upper neighbour's recv_lo
                vs[i][3].copy_(vs[i + 1][0])                           # its first row -> my recv_hi
            if self.dist is not None and self.world > 1:
                dist, ops = self.dist, []
                lo_rank, hi_rank = self.ranks[0] - 1, self.ranks[-1] + 1
                if lo_rank >= 0:
                    ops += [dist.P2POp(dist.isend, vs[0][0], lo_rank), dist.P2POp(dist.irecv, vs[0][2], lo_rank)]
                if hi_rank < self.world:
                    ops += [dist.P2POp(dist.isend, vs[-1][1], hi_rank), dist.P2POp(dist.irecv, vs[-1][3], hi_rank)]
                for req in dist.batch_isend_irecv(ops):
                    req.wait()
            self.halo_bytes += 8*vs[0][0].numel()*2
            self.collectives += 1
        for t in self.slabs:
            t.dist_xchg_finish(what, level)

    def gather(self, what):
        if self.p2p and what == self._capi.G_BK:
            self._all("p2p_gather", 0)
            self._all("p2p_gather", 1)
            self.collectives += 1
            return
        vs = [self._gather_views(i, what) for i in range(len(self.slabs))]
        if self.dist is not None:
            self.dist.all_gather_into_tensor(vs[0][1], vs[0][0])
        else:
            n = vs[0][0].numel()
            for send_i, (send, _) in enumerate(vs):
                for _, recv in vs:
                    recv[send_i*n:(send_i + 1)*n].copy_(send)
        self.collectives += 1

    def allreduce(self, what, n, op="sum"):
        if self.fused:                                                     # one launch, flag-in-data protocol
            self.slabs[0].dist_call("p2p_fused", -1, 0, what, n, 3 if op == "max" else 0)
            self.collectives += 1
            return
        if self.p2p:
            self._all("p2p_allreduce", 0, what, n, 1 if op == "max" else 0)
            self._all("p2p_allreduce", 1, what, n, 1 if op == "max" else 0)
            self.collectives += 1
            return
        vs = [self._scalar_view(i, what)[:n] for i in range(len(self.slabs))]
        if self.dist is not None:
            self.dist.all_reduce(vs[0], op=self.dist.ReduceOp.SUM if op == "sum" else self.dist.ReduceOp.MAX)
        elif len(vs) > 1:
            st = self.torch.stack(vs)
            tot = st.sum(0) if op == "sum" else st.max(0).values
            for v in vs:
                v.copy_(tot)
        self.collectives += 1

    def _all(self, name, *args):
        for t in self.slabs:
            t.dist_call(name, *args)

    # ---- stage 2: multigrid-preconditioned CG ----------------------------------------------------------------------------
    def setup(self):
        c = self._capi
        self._all("setup", 0)
        self.exchange(c.X_DIAG)
        self.gather(c.G_CELLS)
        self._all("setup", 1)

    def _cycle(self, par, init):
        """update + V-cycle + smoothing: everything of an iteration after the K p product."""
        c, K = self._capi, self.K
        if self.fused:
            t = self.slabs[0]
            t.dist_call("pcg_update", par, init)
            for l in range(1, K):
                if not self.inkernel:
                    t.dist_call("p2p_fused", c.X_LEVEL_B, l, 0, 0, 0)
                t.dist_call("vcycle_down", l)
            t.dist_call("p2p_gather_tail")
            for l in range(K - 1, 0, -1):
                t.dist_call("vcycle_up", l)
                if not self.inkernel:
                    t.dist_call("p2p_fused", c.X_LEVEL_X2, l, 0, 0, 0)
            t.dist_call("pcg_smooth", par, init)
            t.dist_call("p2p_fused", c.X_Z, 0, c.S_RR_RZ, 2, 2 if init else 1)
            self.collectives += 2*K + 1
            return
        self._all("pcg_update", par, init)
        for l in range(1, K):
            self.exchange(c.X_LEVEL_B, l)
            self._all("vcycle_down", l)
        self.gather(c.G_BK)
        self._all("vcycle_tail", 1 if self.p2p else 0)
        for l in range(K - 1, 0, -1):
            self._all("vcycle_up", l)
            self.exchange(c.X_LEVEL_X2, l)
        self._all("pcg_smooth", par, init)
        self.exchange(c.X_Z)
        self.allreduce(c.S_RR_RZ, 2)
        self._all("pcg_finish", init)

    def _iteration(self, par):
        c = self._capi
        self._all("pcg_matvec", par)
        if self.fused:
            self.slabs[0].dist_call("p2p_fused", c.X_AP, 0, c.S_PAP, 1, 0)
            self.collectives += 1
        else:
            self.exchange(c.X_AP)
            self.allreduce(c.S_PAP, 1)
        self._cycle(par, 0)

    def _pair(self):
        if not self.use_graph:
            self._iteration(1)
            self._iteration(0)
            return
        torch = self.torch
        if self._graph is None:
            try:
                g = torch.cuda.CUDAGraph()
                self.stream.synchronize()
                with torch.cuda.graph(g, stream=self.stream, capture_error_mode="relaxed"):
                    self._iteration(1)
                    self._iteration(0)
                self._graph = g
            except Exception as exc:                                   # keep running eagerly, but say so
                self.graph_error = repr(exc)
                self.use_graph = False
                self.stream.synchronize()
                self._iteration(1)
                self._iteration(0)
                return
        self._graph.replay()

    def solve(self, warm=True):
        c = self._capi
        self._all("solve_begin", self.rtol, self.maxit, 1 if warm else 0)
        self.exchange(c.X_R0)
        self.allreduce(c.S_BB, 1)
        self._cycle(0, 1)
        launched = 0
        while True:
            done, iters, relres = self.slabs[0].dist_solve_poll()
            if done or launched >= self.maxit:
                break
            if self.p2p and launched and any(t.dist_p2p_error() for t in self.slabs):
                break                                                  # a peer never answered: stop queueing work (raised below)
            # consecutive design iterations need almost the same number of PCG iterations: enqueue the expected
            # count (minus one) before the first poll, then poll after every pair
            if self._last_iters > 4:
                pairs = (self._last_iters - 1)//2 if launched == 0 else 1
            else:
                pairs = 2
            for _ in range(max(1, pairs)):
                self._pair()
            launched += 2*max(1, pairs)
        self._last_iters = iters
        self._all("solve_end", iters)
        if self.p2p and any(t.dist_p2p_error() for t in self.slabs):
            raise self._capi.TopoError(self._capi.ECUDA, "peer-memory transport: a wait for a neighbour's data timed out")
        if relres != relres:
            raise self._capi.TopoError(self._capi.ENOCONV, "row-slab PCG produced NaN (singular system?)")
        if done != 1:
            raise self._capi.TopoError(self._capi.ENOCONV, "row-slab PCG did not converge: %d iterations, relative "
                                                           "residual %.3e" % (iters, relres))
        return iters, relres

    # ---- stage 4: OC bisection ---------------------------------------------------------------------------------------------
    def oc_update(self, g):
        c, ne, mv = self._capi, self.ne_global, self.move
        self._all("oc", 0, mv, float(g), ne)
        probe = self._lam_prev > 0.0
        if probe:                                                      # bracket the root around the previous multiplier
            self._all("oc", 9, mv, self._lam_prev, ne)
            self.allreduce(c.S_RED, 16)
            self._all("oc", 10, mv, self._lam_prev, ne)
        rounds, batch = 0, (4 if probe else 8)
        while rounds < 700:
            for _ in range(batch):
                self._all("oc", 1, mv, 0.0, ne)
                self.allreduce(c.S_RED, 16)
                self._all("oc", 2, mv, 0.0, ne)
            rounds += batch
            done, need = self.slabs[0].dist_oc_state()[:2]
            if need:                                                   # a sum inside the rounding margin: exact step
                self._all("oc", 3, mv, 0.0, ne)
                self.allreduce(c.S_RED, 1)
                self._all("oc", 4, mv, 0.0, ne)
                continue
            if done:
                break
        self._all("oc", 5, mv, 0.0, ne)
        self.allreduce(c.S_RED, 1)
        self._all("oc", 6, mv, 0.0, ne)
        self._all("oc", 7, mv, 0.0, ne)
        self.allreduce(c.S_RED, 1, op="max")
        self._all("oc", 8, mv, 0.0, ne)
        _, _, steps, gt, change, lmid = self.slabs[0].dist_oc_state()
        self._lam_prev = lmid if (steps > 0 and 0.0 < lmid < 1e300) else 0.0
        return gt, change, steps

    # ---- one design iteration (reference optimize.py:428-456) ------------------------------------------------------------------
    def step(self, g, warm=True):
        """Returns (g_new, dict(compliance, objective, change, cg_iters, relres, oc_steps, stage_ms))."""
        c, torch = self._capi, self.torch
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        with torch.cuda.stream(self.stream):
            ev[0].record()
            for t in self.slabs:
                t.simp_scale(self.penal, self.emin, self.emax)
            self.setup()
            ev[1].record()
            iters, relres = self.solve(warm)
            ev[2].record()
            for t in self.slabs:
                t.simp_sensitivity_partial(self.penal, self.emin, self.emax)
                t.dist_call("compliance_partial")
            self.allreduce(c.S_RED, 2)
            obj, comp = self.slabs[0].dist_read_scalars(2)
            ev[3].record()
            if self.rmin is not None:
                for t in self.slabs:
                    t.dist_call("filter_pack", self.rmin, self.dy)
                self.exchange(c.X_FILTER)
                for t in self.slabs:
                    t.filter_sensitivity(self.rmin, self.dx, self.dy)
            ev[4].record()
            g, change, steps = self.oc_update(g)
            ev[5].record()
            ev[5].synchronize()
        names = ("setup", "solve", "sens", "filter", "oc")
        stage_ms = {n: ev[i].elapsed_time(ev[i + 1]) for i, n in enumerate(names)}
        return g, {"compliance": float(comp), "objective": float(obj), "change": change, "cg_iters": iters,
                   "relres": relres, "oc_steps": steps, "stage_ms": stage_ms}

    # ---- state transfer ----------------------------------------------------------------------------------------------------
    def set_rho(self, rho_global):
        rho = np.asarray(rho_global, dtype=np.float64).reshape(self.ny, self.nx)
        for t, g in zip(self.slabs, self.ranks):
            t.set(self._capi.F_RHO, rho[g*self.nyl:(g + 1)*self.nyl].reshape(-1))

    def local_rho(self):
        return np.concatenate([t.get(self._capi.F_RHO).reshape(self.nyl, self.nx) for t in self.slabs])

    def local_u(self):
        """Displacements of this process' slabs, node rows stacked without the duplicated interface rows."""
        out = []
        for i, t in enumerate(self.slabs):
            u = t.get(self._capi.F_U).reshape(self.nyl + 1, self.nx + 1, 2)
            out.append(u if i == len(self.slabs) - 1 else u[:-1])
        return np.concatenate(out)

    def close(self):
        self._graph = None
        self._views = {}
        for t in self.slabs:
            t.close()
        self.slabs = []
