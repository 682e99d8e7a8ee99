// Stage 2: preconditioned conjugate gradients on K u = f in full DOF space (replaces
// scipy.sparse.linalg.spsolve, reference optimize.py:61,84,161,181,260,298,437).
//
// One PCG iteration with the multigrid preconditioner is
//   K2  u' = u + alpha p ; r' = r - alpha Ap ; ||r'||^2 ; x^ = omega D^-1 r' ; res = r' - A x^   (fine.cu)
//       restrict res, coarse levels of the V-cycle, prolongate into x^                                (gmg.cu)
//   K4  z = x^ + omega D^-1 (r' - A x^) ; r'.z ; beta                                                 (fine.cu)
//   K1  p' = z + beta p ; Ap = A p' ; p'.Ap                                                           (fine.cu)
// i.e. three passes over the fine grid, every vector update and dot product fused into them.  All PCG
// scalars (alpha, beta, norms, iteration count, convergence flag) live in device memory
// (SolverScalars); the host only enqueues batches of iterations and polls the flag between batches.
// Dot products are warp-shuffle + block + ordered grid reductions (deterministic).
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "topo_internal.cuh"

int topo_fine_smooth(topo_handle *h, const double2 *x, const double2 *b, double2 *x_out, double omega);
int topo_fine_pcg_matvec(topo_handle *h, const double2 *z, const double2 *p_in, double2 *p_out, double2 *Ap);
int topo_fine_pcg_matvec_jacobi(topo_handle *h, const double2 *r, const double2 *p_in, double2 *p_out, double2 *Ap);
int topo_fine_update_resid0(topo_handle *h, const double2 *p, const double2 *Ap, const double2 *u_in,
                            const double2 *r_in, double2 *u_out, double2 *r_out, double2 *xhat, double2 *res,
                            double omega, int init);
int topo_fine_prolong_smooth(topo_handle *h, const double2 *xhat, const double2 *r, const double2 *xc, int cnx,
                             double2 *z, double omega, int init);
int topo_gmg_coarse(topo_handle *h, const double2 *res_fine, double2 *xhat_fine, const double2 **x1_out);

namespace {

constexpr int VEC_THREADS = 256;
inline int vec_blocks(int64_t n) { return (int)std::min<int64_t>(div_up(n, VEC_THREADS), 148 * 8); }

// Jacobi path: rz = r.(dinv r), rr, bb ; beta = 0
__global__ void k_pcg_init_jacobi(const double2 *__restrict__ r, const double2 *__restrict__ f,
                                  const double2 *__restrict__ dinv, int64_t nn, double *partials,
                                  unsigned int *ticket, SolverScalars *sc) {
  double v[3] = {0.0, 0.0, 0.0};
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < nn; n += (int64_t)gridDim.x * blockDim.x) {
    const double2 rv = r[n], fv = f[n], d = dinv[n];
    v[0] = fma(rv.x, d.x * rv.x, fma(rv.y, d.y * rv.y, v[0]));
    v[1] = fma(rv.x, rv.x, fma(rv.y, rv.y, v[1]));
    v[2] = fma(fv.x, fv.x, fma(fv.y, fv.y, v[2]));
  }
  if (grid_reduce<3>(v, partials, ticket, sc->aux) && threadIdx.x == 0) {
    sc->rz[0] = v[0];
    sc->rz[1] = 0.0;
    sc->rr = v[1];
    sc->bb = v[2];
    sc->iters = 0;
    sc->done = (v[1] <= sc->tol2 * v[2]) ? 1 : 0;
  }
}

// Jacobi path: x += alpha p ; r -= alpha Ap ; rz_new = r.(dinv r) ; rr ; beta ; convergence flag
__global__ void k_pcg_update_jacobi(double2 *__restrict__ x, double2 *__restrict__ r,
                                    const double2 *__restrict__ p, const double2 *__restrict__ Ap,
                                    const double2 *__restrict__ dinv, int64_t nn, double *partials,
                                    unsigned int *ticket, SolverScalars *sc) {
  if (sc->done) return;
  const double alpha = sc->rz[0] / sc->pAp;
  double v[2] = {0.0, 0.0};
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < nn; n += (int64_t)gridDim.x * blockDim.x) {
    const double2 pv = p[n], av = Ap[n], d = dinv[n];
    double2 xv = x[n], rv = r[n];
    xv.x = fma(alpha, pv.x, xv.x);
    xv.y = fma(alpha, pv.y, xv.y);
    rv.x = fma(-alpha, av.x, rv.x);
    rv.y = fma(-alpha, av.y, rv.y);
    x[n] = xv;
    r[n] = rv;
    v[0] = fma(rv.x, d.x * rv.x, fma(rv.y, d.y * rv.y, v[0]));
    v[1] = fma(rv.x, rv.x, fma(rv.y, rv.y, v[1]));
  }
  if (grid_reduce<2>(v, partials, ticket, sc->aux) && threadIdx.x == 0) {
    sc->rz[1] = v[0] / sc->rz[0];
    sc->rz[0] = v[0];
    sc->rr = v[1];
    sc->iters += 1;
    if (v[1] <= sc->tol2 * sc->bb || !(v[1] == v[1])) sc->done = 1;
    else if (sc->iters >= sc->maxit) sc->done = 2;
  }
}

__global__ void k_dot(const double2 *__restrict__ a, const double2 *__restrict__ b, int64_t nn,
                      double *partials, unsigned int *ticket, double *out) {
  double v[1] = {0.0};
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < nn; n += (int64_t)gridDim.x * blockDim.x) {
    const double2 x = a[n], y = b[n];
    v[0] = fma(x.x, y.x, fma(x.y, y.y, v[0]));
  }
  grid_reduce<1>(v, partials, ticket, out);
}

// Equilibrium guard of the reference: all(|Ku - f|/kmax <= atol + rtol*|f|/kmax), atol=1e-8, rtol=1e-5
// (np.allclose defaults; optimize.py:455).  `r` holds f - K u on free DOFs.  Reduces the count of
// violating DOFs (NaN counts as a violation, like np.allclose).
__global__ void k_equilibrium(const double2 *__restrict__ r, const double2 *__restrict__ f, int64_t nn,
                              const double *kmax, double *partials, unsigned int *ticket, double *out) {
  const double km = *kmax;
  double v[1] = {0.0};
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < nn; n += (int64_t)gridDim.x * blockDim.x) {
    const double2 rv = r[n], fv = f[n];
    const bool okx = fabs(rv.x / km) <= 1e-8 + 1e-5 * fabs(fv.x / km);
    const bool oky = fabs(rv.y / km) <= 1e-8 + 1e-5 * fabs(fv.y / km);
    v[0] += (okx ? 0.0 : 1.0) + (oky ? 0.0 : 1.0);
  }
  grid_reduce<1>(v, partials, ticket, out);
}

__global__ void k_zero2(double2 *p, int64_t nn) {
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < nn; n += (int64_t)gridDim.x * blockDim.x)
    p[n] = make_double2(0.0, 0.0);
}

// loadasem semantics (assignment in row order; constrained DOFs skipped): one thread walks the rows.
__global__ void k_apply_loads(double2 *f, const uint8_t *fixed, const int64_t *node, const double *fx,
                              const double *fy, int n) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int i = 0; i < n; ++i) {
      const int64_t nd = node[i];
      const unsigned m = fixed[nd];
      double2 v = f[nd];
      if (!(m & 1u)) v.x = fx[i];
      if (!(m & 2u)) v.y = fy[i];
      f[nd] = v;
    }
  }
}

__global__ void k_mask2(double2 *v, const uint8_t *fixed, int64_t nn) {
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < nn; n += (int64_t)gridDim.x * blockDim.x) {
    const unsigned m = fixed[n];
    double2 x = v[n];
    if (m & 1u) x.x = 0.0;
    if (m & 2u) x.y = 0.0;
    v[n] = x;
  }
}

__global__ void k_set_bb(SolverScalars *sc, const double *bb) {
  sc->bb = *bb;
}

}  // namespace


int topo_apply_loads(topo_handle *h) {
  const int64_t nn = h->nn;
  KL(h, "k_zero2", k_zero2<<<vec_blocks(nn), VEC_THREADS, 0, h->stream>>>(h->f, nn));
  if (h->n_loads > 0) {
    KL(h, "k_apply_loads", k_apply_loads<<<1, 32, 0, h->stream>>>(h->f, h->fixed, h->d_load_node, h->d_load_fx, h->d_load_fy,
                                           h->n_loads));
  }
  CUDA_TRY(cudaGetLastError());
  return TOPO_OK;
}

int topo_mask_vector(topo_handle *h, double2 *v) {
  KL(h, "k_mask2", k_mask2<<<vec_blocks(h->nn), VEC_THREADS, 0, h->stream>>>(v, h->fixed, h->nn));
  CUDA_TRY(cudaGetLastError());
  return TOPO_OK;
}

int topo_dot(topo_handle *h, const double2 *a, const double2 *b, double *host_out) {
  KL(h, "k_dot", k_dot<<<vec_blocks(h->nn), VEC_THREADS, 0, h->stream>>>(a, b, h->nn, h->partials, h->ticket, h->red_out));
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(host_out, h->red_out, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return TOPO_OK;
}

extern "C" int topo_solve(topo_t *h, int precond, double rtol, int maxit, int warm, int *iters, double *relres) {
  if (!h) return TOPO_EINVAL;
  if (!h->loads_set || !h->cons_set) {
    topo_set_error("topo_solve: set BC flags and loads first");
    return TOPO_ESTATE;
  }
  if (precond != TOPO_PRECOND_JACOBI && precond != TOPO_PRECOND_GMG) {
    topo_set_error("topo_solve: unknown preconditioner %d", precond);
    return TOPO_EINVAL;
  }
  if (h->dist.on) {
    topo_set_error("topo_solve: row-slab handle -- the solve is driven through topo_dist_*");
    return TOPO_ESTATE;
  }
  CUDA_TRY(cudaSetDevice(h->device));
  const int64_t nn = h->nn;
  const int nb = vec_blocks(nn);
  const size_t vbytes = sizeof(double2) * (size_t)nn;
  cudaStream_t st = h->stream;
  CUDA_TRY(cudaEventRecord(h->ev0, st));

  TOPO_TRY(topo_build_diag(h));
  const bool jacobi = (precond == TOPO_PRECOND_JACOBI) || h->n_levels < 2;
  if (!jacobi) TOPO_TRY(topo_gmg_setup(h));

  if (!warm) {
    KL(h, "k_zero2", k_zero2<<<nb, VEC_THREADS, 0, st>>>(h->u, nn));
  } else {
    TOPO_TRY(topo_mask_vector(h, h->u));   // BC flags may have changed since the last solve
  }
  SolverScalars init = {};
  init.tol2 = rtol * rtol;
  init.maxit = maxit;
  CUDA_TRY(cudaMemcpyAsync(h->sc, &init, sizeof(init), cudaMemcpyHostToDevice, st));

  // ping-pong pairs (inputs that are also outputs are written to the other buffer)
  double2 *ub[2] = {h->u, h->u2}, *rb[2] = {h->r, h->r2}, *pb[2] = {h->p, h->p2};
  double2 *xhat = h->tmp2, *res = h->tmp;
  TOPO_TRY(topo_fine_residual(h, ub[0], h->f, rb[0]));
  SolverScalars *hs = h->sc_host;
  hs->done = 0;
  hs->iters = 0;
  int launched = 0;

  if (jacobi) {
    CUDA_TRY(cudaMemsetAsync(pb[0], 0, vbytes, st));
    KL(h, "k_pcg_init_jacobi", k_pcg_init_jacobi<<<nb, VEC_THREADS, 0, st>>>(rb[0], h->f, h->dinv, nn, h->partials, h->ticket, h->sc));
    int pcur = 0;
    while (true) {
      CUDA_TRY(cudaMemcpyAsync(hs, h->sc, sizeof(SolverScalars), cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaStreamSynchronize(st));
      if (hs->done || launched >= maxit) break;
      const int nrun = std::min(64, maxit - launched);
      for (int i = 0; i < nrun; ++i) {
        TOPO_TRY(topo_fine_pcg_matvec_jacobi(h, rb[0], pb[pcur], pb[pcur ^ 1], h->Ap));
        pcur ^= 1;
        KL(h, "k_pcg_update_jacobi", k_pcg_update_jacobi<<<nb, VEC_THREADS, 0, st>>>(ub[0], rb[0], pb[pcur], h->Ap, h->dinv, nn,
                                                                                    h->partials, h->ticket, h->sc));
      }
      launched += nrun;
      CUDA_TRY(cudaGetLastError());
    }
  } else {
    const bool fuse_prolong = h->knobs.fuse_prolong != 0;
    // one PCG iteration reading the buffers of parity `par` and writing those of parity par^1
    auto iteration = [&](int par) -> int {
      TOPO_TRY(topo_fine_pcg_matvec(h, h->z, pb[par], pb[par ^ 1], h->Ap));
      TOPO_TRY(topo_fine_update_resid0(h, pb[par ^ 1], h->Ap, ub[par], rb[par], ub[par ^ 1], rb[par ^ 1], xhat, res,
                                       h->omega, 0));
      const double2 *x1 = nullptr;
      TOPO_TRY(topo_gmg_coarse(h, res, xhat, fuse_prolong ? &x1 : nullptr));
      TOPO_TRY(topo_fine_prolong_smooth(h, xhat, rb[par ^ 1], x1, h->lvl_nx[1], h->z, h->omega, 0));
      return TOPO_OK;
    };
    CUDA_TRY(cudaMemsetAsync(pb[1], 0, vbytes, st));
    CUDA_TRY(cudaMemsetAsync(h->Ap, 0, vbytes, st));
    KL(h, "k_dot", k_dot<<<nb, VEC_THREADS, 0, st>>>(h->f, h->f, nn, h->partials, h->ticket, h->red_out));
    KL(h, "k_set_bb", k_set_bb<<<1, 1, 0, st>>>(h->sc, h->red_out));
    // z0 = V(r0): the update kernel with alpha = 0 copies u, r into the parity-1 buffers
    TOPO_TRY(topo_fine_update_resid0(h, pb[1], h->Ap, ub[0], rb[0], ub[1], rb[1], xhat, res, h->omega, 1));
    {
      const double2 *x1 = nullptr;
      TOPO_TRY(topo_gmg_coarse(h, res, xhat, fuse_prolong ? &x1 : nullptr));
      TOPO_TRY(topo_fine_prolong_smooth(h, xhat, rb[1], x1, h->lvl_nx[1], h->z, h->omega, 1));
    }

    // Two iterations (parity 1 then 0) form a closed cycle of buffer roles: capture them once in a CUDA
    // graph (per handle) and replay it; converged iterations turn into no-op kernels.
    const bool use_graph = !h->prof_on && h->use_graph;
    if (use_graph && !h->pcg_graph_valid) {
      if (h->pcg_graph_exec) { cudaGraphExecDestroy(h->pcg_graph_exec); h->pcg_graph_exec = nullptr; }
      cudaGraph_t graph = nullptr;
      CUDA_TRY(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      const long long l0 = h->timing.kernel_launches;
      int s1 = iteration(1);
      int s2 = (s1 == TOPO_OK) ? iteration(0) : s1;
      h->pcg_graph_launches = (int)(h->timing.kernel_launches - l0);
      h->timing.kernel_launches = l0;            // captured, not launched
      cudaError_t ce = cudaStreamEndCapture(st, &graph);
      if (s2 != TOPO_OK) { if (graph) cudaGraphDestroy(graph); return s2; }
      CUDA_TRY(ce);
      CUDA_TRY(cudaGraphInstantiate(&h->pcg_graph_exec, graph, 0));
      CUDA_TRY(cudaGraphDestroy(graph));
      h->pcg_graph_valid = true;
    }
    while (true) {
      CUDA_TRY(cudaMemcpyAsync(hs, h->sc, sizeof(SolverScalars), cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaStreamSynchronize(st));
      if (hs->done || launched >= maxit) break;
      // consecutive design iterations need almost the same number of PCG iterations: enqueue the expected count
      // (minus one) before the first poll, then poll after every pair -- few host round trips, at most one
      // pair of no-op iterations after convergence
      int pairs = (launched == 0 && h->last_pcg_iters > 4) ? (h->last_pcg_iters - 1) / 2 : (h->last_pcg_iters > 4 ? 1 : 4);
      pairs = std::max(1, std::min(pairs, (maxit - launched + 1) / 2));
      for (int i = 0; i < pairs; ++i) {
        if (use_graph) {
          CUDA_TRY(cudaGraphLaunch(h->pcg_graph_exec, st));
          h->timing.kernel_launches += h->pcg_graph_launches;
        } else {
          TOPO_TRY(iteration(1));
          TOPO_TRY(iteration(0));
        }
      }
      launched += 2 * pairs;
      CUDA_TRY(cudaGetLastError());
    }
    // iteration k wrote the buffers of parity (k+1)&1 (the alpha = 0 pass wrote parity 1)
    if (((hs->iters + 1) & 1) == 1)
      CUDA_TRY(cudaMemcpyAsync(h->u, h->u2, vbytes, cudaMemcpyDeviceToDevice, st));
  }
  CUDA_TRY(cudaEventRecord(h->ev1, st));
  CUDA_TRY(cudaEventSynchronize(h->ev1));
  CUDA_TRY(cudaEventElapsedTime(&h->timing.solve_ms, h->ev0, h->ev1));
  h->last_pcg_iters = jacobi ? 0 : hs->iters;
  if (iters) *iters = hs->iters;
  const double rel = (hs->bb > 0.0) ? std::sqrt(hs->rr / hs->bb) : (hs->rr == 0.0 ? 0.0 : INFINITY);
  if (relres) *relres = rel;
  if (hs->tail_error) {
    topo_set_error("multigrid tail: a hand-over between thread blocks timed out (results invalid)");
    return TOPO_ECUDA;
  }
  if (!(hs->rr == hs->rr)) {
    topo_set_error("PCG produced NaN (singular system?)");
    return TOPO_ENOCONV;
  }
  if (hs->done != 1) {
    topo_set_error("PCG did not converge: %d iterations, relative residual %.3e (rtol %.1e)", hs->iters, rel,
                   rtol);
    return TOPO_ENOCONV;
  }
  return TOPO_OK;
}

extern "C" int topo_compliance(topo_t *h, double *c) {
  if (!h || !c) return TOPO_EINVAL;
  CUDA_TRY(cudaSetDevice(h->device));
  return topo_dot(h, h->f, h->u, c);
}

extern "C" int topo_equilibrium_check(topo_t *h, int *ok) {
  if (!h || !ok) return TOPO_EINVAL;
  CUDA_TRY(cudaSetDevice(h->device));
  if (!h->diag_valid) TOPO_TRY(topo_build_diag(h));
  TOPO_TRY(topo_fine_residual(h, h->u, h->f, h->tmp));
  KL(h, "k_equilibrium", k_equilibrium<<<vec_blocks(h->nn), VEC_THREADS, 0, h->stream>>>(h->tmp, h->f, h->nn, h->kmax, h->partials,
                                                                  h->ticket, h->red_out));
  CUDA_TRY(cudaGetLastError());
  double bad = 0.0;
  CUDA_TRY(cudaMemcpyAsync(&bad, h->red_out, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  *ok = (bad == 0.0) ? 1 : 0;
  return TOPO_OK;
}
