// Row-slab decomposition of the SIMP design iteration over several GPUs (no counterpart in the
// reference, which is single-process).  One handle = one slab of element rows on one GPU.  This file holds the
// C-ABI entry points the host loop drives: the PCG iteration and the V-cycle cut into SEGMENTS at the points where
// interface node rows must be completed by the neighbours (halo sum), where level-K data is gathered, and where
// scalars are all-reduced.  The library never communicates itself: it exposes device pointers of what has to be
// moved (topo_dist_xchg_ptrs / topo_dist_gather_ptrs / topo_dist_scalar_ptr) and finishes an exchange with a small
// row kernel (topo_dist_xchg_finish); the host moves the bytes with NCCL (solidspy_opt/distributed.py).
//
// Vector forms (see DistState): u, p, x, z, r are ACCUMULATED; K_loc p, restricted residuals and local smoothing
// results are DISTRIBUTED/partial on interface rows.
//   Ap  = K_loc p                      -> halo sum          (p.Ap needs no weights: sum over ranks of p.Ap_loc)
//   res = W r - K_loc x^               distributed, restricted as is (restriction is linear, no exchange)
//   b_l = R res                        -> halo sum, then x = w D^-1 b needs the accumulated b
//   x2  = x + w D^-1 (W b - A_loc x)   partial on interface rows: x2 <- x2 + x2_neighbour - x
//   z   likewise; r.z partial = r.(z - (1-W) x~)
#include <algorithm>
#include <cmath>
#include <cstring>

#include "topo_internal.cuh"
#include "p2p_ll.cuh"

int topo_fine_pcg_matvec(topo_handle *h, const double2 *z, const double2 *p_in, double2 *p_out, double2 *Ap);
int topo_fine_update_resid0(topo_handle *h, const double2 *p, const double2 *Ap, const double2 *u_in,
                            const double2 *r_in, double2 *u_out, double2 *r_out, double2 *xhat, double2 *res,
                            double omega, int init);
int topo_fine_prolong_smooth(topo_handle *h, const double2 *xhat, const double2 *r, const double2 *xc, int cnx,
                             double2 *z, double omega, int init);
int topo_mask_vector(topo_handle *h, double2 *v);
int topo_dist_restrict_fine(topo_handle *h, const double2 *res_fine);
int topo_dist_level_down(topo_handle *h, int l);
int topo_dist_tail(topo_handle *h, int p2p);
int topo_dist_level_up(topo_handle *h, int l);
int topo_dist_prolong_fine(topo_handle *h, double2 *xhat_fine);
int topo_dist_setup_local(topo_handle *h);
int topo_dist_setup_global(topo_handle *h);

namespace {

inline int nblk(int64_t n, int t = 256) { return (int)std::min<int64_t>((n + t - 1) / t, 148 * 8); }

// dst += src
__global__ void k_row_add(double2 *__restrict__ dst, const double2 *__restrict__ src, int n, const int *done) {
  if (done && *done) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double2 a = dst[i];
  const double2 b = src[i];
  a.x += b.x; a.y += b.y;
  dst[i] = a;
}
// dst = dst + src - base   (completes a locally smoothed interface row: see the file header)
// (base_f32: the base row is a float2 array -- the fine-grid x~)
__global__ void k_row_add_sub(double2 *__restrict__ dst, const double2 *__restrict__ src, const double2 *__restrict__ base,
                              int base_f32, int n, const int *done) {
  if (done && *done) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double2 a = dst[i];
  const double2 b = src[i];
  double2 c;
  if (base_f32) {
    const float2 f = reinterpret_cast<const float2 *>(base)[i];
    c = make_double2((double)f.x, (double)f.y);
  } else {
    c = base[i];
  }
  a.x = (a.x + b.x) - c.x; a.y = (a.y + b.y) - c.y;
  dst[i] = a;
}

// sum_n W(row) a.b  (accumulated x accumulated)
__global__ void k_dot_weighted(const double2 *__restrict__ a, const double2 *__restrict__ b, int npr, int ny, double wf,
                               double wl, double *partials, unsigned int *ticket, double *out) {
  const int64_t nn = (int64_t)npr * (ny + 1);
  double v[1] = {0.0};
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < nn; n += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(n / npr);
    const double w = r == 0 ? wf : (r == ny ? wl : 1.0);
    const double2 x = a[n], y = b[n];
    v[0] = fma(w * x.x, y.x, fma(w * x.y, y.y, v[0]));
  }
  grid_reduce<1>(v, partials, ticket, out);
}

__global__ void k_zero2d(double2 *p, int64_t nn) {
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < nn; n += (int64_t)gridDim.x * blockDim.x)
    p[n] = make_double2(0.0, 0.0);
}

// scalar bookkeeping after the all-reduce of (rr, rz) in sc->aux[1..2] (what the last block of the fused kernels
// does on a single GPU)
__global__ void k_dist_pcg_finish(SolverScalars *sc, int init) {
  if (!init && sc->done) return;
  const double rr = sc->aux[1], rz = sc->aux[2];
  sc->rr = rr;
  if (!init) sc->iters += 1;
  sc->rz[1] = init ? 0.0 : rz / sc->rz[0];
  sc->rz[0] = rz;
  if (rr <= sc->tol2 * sc->bb || !(rr == rr)) sc->done = 1;
  else if (sc->iters >= sc->maxit) sc->done = 2;
}
__global__ void k_dist_set_bb(SolverScalars *sc) { sc->bb = sc->aux[3]; }

int check(topo_handle *h) {
  if (!h) return TOPO_EINVAL;
  if (!h->dist.on) {
    topo_set_error("handle is not in row-slab mode (call topo_dist_init first)");
    return TOPO_ESTATE;
  }
  return TOPO_OK;
}

void free_dist(DistState &D) {
  cudaFree(D.recv_lo); cudaFree(D.recv_hi);
  cudaFree(D.diag_send_lo); cudaFree(D.diag_send_hi); cudaFree(D.diag_recv_lo); cudaFree(D.diag_recv_hi);
  cudaFree(D.cellK_send); cudaFree(D.cellK_recv);
  cudaFree(D.bK_send); cudaFree(D.bK_recv);
  cudaFree(D.filt_send_lo); cudaFree(D.filt_send_hi); cudaFree(D.filt_recv_lo); cudaFree(D.filt_recv_hi);
  for (int r = 0; r < TOPO_P2P_MAXW; ++r)
    if (D.p2p_opened[r]) cudaIpcCloseMemHandle(D.p2p_peer[r]);
  cudaFree(D.p2p_region);
  cudaFree(D.p2p_cnt);
  cudaFree(D.mid_bar);
  cudaGetLastError();
  D = DistState{};
}

}  // namespace

void topo_dist_free(topo_handle *h) { free_dist(h->dist); }

extern "C" int topo_dist_init(topo_t *h, int rank, int world, int K_request) {
  if (!h || world < 1 || rank < 0 || rank >= world) return TOPO_EINVAL;
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  const int ny_global = h->ny * world;
  // K: the first level of the GLOBAL hierarchy small enough for the tail kernel (or the coarsest one)
  const int coarsest_max_nodes = h->knobs.coarsest_nodes;
  int K = 0;
  {
    int nx = h->nx, ny = ny_global, l = 0;
    while ((nx + 1) * (int64_t)(ny + 1) > coarsest_max_nodes && std::max(nx, ny) > 1) {
      nx = (nx + 1) / 2; ny = (ny + 1) / 2; ++l;
      if (K == 0 && (int64_t)(nx + 1) * (ny + 1) <= h->tail_max_nodes) K = l;
    }
    if (K == 0) K = l;
    if (K_request > 0) K = std::min(K_request, std::max(l, 1));
  }
  if (K < 1) {
    topo_set_error("topo_dist_init: the mesh is too small for a multigrid hierarchy");
    return TOPO_EINVAL;
  }
  if (h->ny % (1 << K) != 0) {
    topo_set_error("topo_dist_init: slab height %d must be a multiple of 2^K = %d (K = first global level)", h->ny, 1 << K);
    return TOPO_EINVAL;
  }
  if (h->nu_pre != 1 || h->nu_post != 1 || h->nu_coarse > 0) {
    topo_set_error("topo_dist_init: row-slab mode supports one smoothing sweep on the slab-local levels");
    return TOPO_EINVAL;
  }
  topo_gmg_free(h);
  free_dist(h->dist);
  DistState &D = h->dist;
  D.on = true;
  D.rank = rank; D.world = world; D.K = K;
  D.has_lo = rank > 0; D.has_hi = rank < world - 1;
  D.ny_global = ny_global;
  D.w_first = D.has_lo ? 0.5 : 1.0;
  D.w_last = D.has_hi ? 0.5 : 1.0;
  TOPO_TRY(topo_gmg_plan(h));
  const size_t row = sizeof(double2) * (size_t)(h->nx + 1);
  CUDA_TRY(cudaMalloc(&D.recv_lo, row));
  CUDA_TRY(cudaMalloc(&D.recv_hi, row));
  D.diag_count = 0;
  for (int l = 0; l < K; ++l) D.diag_count += h->lvl_nx[l] + 1;
  const size_t dbytes = sizeof(double2) * (size_t)D.diag_count;
  CUDA_TRY(cudaMalloc(&D.diag_send_lo, dbytes)); CUDA_TRY(cudaMalloc(&D.diag_send_hi, dbytes));
  CUDA_TRY(cudaMalloc(&D.diag_recv_lo, dbytes)); CUDA_TRY(cudaMalloc(&D.diag_recv_hi, dbytes));
  CUDA_TRY(cudaMemset(D.diag_recv_lo, 0, dbytes)); CUDA_TRY(cudaMemset(D.diag_recv_hi, 0, dbytes));
  const int nxK = h->lvl_nx[K], rowsK = h->ny >> K;
  D.ncellK_loc = (long long)nxK * rowsK;
  CUDA_TRY(cudaMalloc(&D.cellK_send, sizeof(double) * 64 * (size_t)D.ncellK_loc));
  CUDA_TRY(cudaMalloc(&D.cellK_recv, sizeof(double) * 64 * (size_t)D.ncellK_loc * world));
  D.bK_count = (long long)(nxK + 1) * (rowsK + 1);
  CUDA_TRY(cudaMalloc(&D.bK_send, sizeof(double2) * (size_t)D.bK_count));
  CUDA_TRY(cudaMalloc(&D.bK_recv, sizeof(double2) * (size_t)D.bK_count * world));
  const size_t fbytes = sizeof(double) * 8 * (size_t)h->nx;      // FH_MAX = 8 element rows
  CUDA_TRY(cudaMalloc(&D.filt_send_lo, fbytes)); CUDA_TRY(cudaMalloc(&D.filt_send_hi, fbytes));
  CUDA_TRY(cudaMalloc(&D.filt_recv_lo, fbytes)); CUDA_TRY(cudaMalloc(&D.filt_recv_hi, fbytes));
  D.filt_hy = 0;
  {
    // peer-memory region (mapped by the peers through topo_dist_p2p_attach) + local counters
    const size_t rowb = (sizeof(double2) * (size_t)(h->nx + 1) + 255) & ~(size_t)255;
    const size_t gb = (sizeof(double2) * (size_t)D.bK_count * world + 255) & ~(size_t)255;
    D.p2p_row_stride = rowb;
    D.p2p_gather_stride = gb;
    D.p2p_off_lo = (sizeof(P2PHeader) + 4095) & ~(size_t)4095;
    D.p2p_off_hi = D.p2p_off_lo + 2 * rowb;
    D.p2p_off_gather = D.p2p_off_hi + 2 * rowb;
    D.p2p_off_ll_lo = D.p2p_off_gather + 2 * gb;
    D.p2p_off_ll_hi = D.p2p_off_ll_lo + 4 * rowb;          // [2 parities][2 * rowb]
    D.p2p_off_ll_gather = D.p2p_off_ll_hi + 4 * rowb;
    D.p2p_bytes = D.p2p_off_ll_gather + 4 * gb;
    CUDA_TRY(cudaMalloc(&D.p2p_region, D.p2p_bytes));
    CUDA_TRY(cudaMemset(D.p2p_region, 0, D.p2p_bytes));
    CUDA_TRY(cudaMalloc(&D.p2p_cnt, sizeof(P2PCounters)));
    CUDA_TRY(cudaMemset(D.p2p_cnt, 0, sizeof(P2PCounters)));
    CUDA_TRY(cudaMalloc(&D.mid_bar, 2 * sizeof(unsigned int)));
    CUDA_TRY(cudaMemset(D.mid_bar, 0, 2 * sizeof(unsigned int)));
    D.mid_from = 0;
    D.p2p_on = false;
  }
  h->diag_valid = h->gmg_valid = false;
  h->pcg_graph_valid = false;
  return TOPO_OK;
}

extern "C" int topo_dist_info(topo_t *h, int *rank, int *world, int *K, int *n_levels) {
  TOPO_TRY(check(h));
  if (rank) *rank = h->dist.rank;
  if (world) *world = h->dist.world;
  if (K) *K = h->dist.K;
  if (n_levels) *n_levels = h->n_levels;
  return TOPO_OK;
}

// What an exchange moves: send_lo goes to rank-1 (and arrives in ITS recv_hi), send_hi goes to rank+1 (arrives in
// its recv_lo); n_doubles per direction.
extern "C" int topo_dist_xchg_ptrs(topo_t *h, int what, int level, void **send_lo, void **send_hi, void **recv_lo,
                                   void **recv_hi, long long *n_doubles) {
  TOPO_TRY(check(h));
  if (!send_lo || !send_hi || !recv_lo || !recv_hi || !n_doubles) return TOPO_EINVAL;
  DistState &D = h->dist;
  double2 *v = nullptr;
  int npr = h->nx + 1, ny = h->ny;
  *recv_lo = D.recv_lo; *recv_hi = D.recv_hi;
  switch (what) {
    case TOPO_X_AP: v = h->Ap; break;
    case TOPO_X_R0: v = h->r; break;
    case TOPO_X_Z: v = h->z; break;
    case TOPO_X_LEVEL_B:
    case TOPO_X_LEVEL_X2:
      if (level < 1 || level >= D.K) { topo_set_error("exchange: level %d is not a slab-local coarse level", level); return TOPO_EINVAL; }
      v = (what == TOPO_X_LEVEL_B) ? h->levels[level].b : h->levels[level].x2;
      npr = h->levels[level].nx + 1; ny = h->levels[level].ny;
      break;
    case TOPO_X_DIAG:
      *send_lo = D.diag_send_lo; *send_hi = D.diag_send_hi; *recv_lo = D.diag_recv_lo; *recv_hi = D.diag_recv_hi;
      *n_doubles = 2 * D.diag_count;
      return TOPO_OK;
    case TOPO_X_FILTER:
      *send_lo = D.filt_send_lo; *send_hi = D.filt_send_hi; *recv_lo = D.filt_recv_lo; *recv_hi = D.filt_recv_hi;
      *n_doubles = (long long)D.filt_hy * h->nx;
      return TOPO_OK;
    default:
      topo_set_error("exchange: unknown kind %d", what);
      return TOPO_EINVAL;
  }
  *send_lo = v;
  *send_hi = v + (size_t)ny * npr;
  *n_doubles = 2LL * npr;
  return TOPO_OK;
}

// completes the interface rows after the host moved the neighbours' rows into recv_lo / recv_hi
extern "C" int topo_dist_xchg_finish(topo_t *h, int what, int level) {
  TOPO_TRY(check(h));
  CUDA_TRY(cudaSetDevice(h->device));
  DistState &D = h->dist;
  cudaStream_t st = h->stream;
  const int *done = &h->sc->done;
  double2 *v = nullptr;
  const double2 *base = nullptr;
  int npr = h->nx + 1, ny = h->ny;
  switch (what) {
    case TOPO_X_AP: v = h->Ap; break;
    case TOPO_X_R0: v = h->r; done = nullptr; break;
    case TOPO_X_Z: v = h->z; base = h->tmp2; break;                 // x~ (x^ after the prolongation) lives in tmp2
    case TOPO_X_LEVEL_B:
    case TOPO_X_LEVEL_X2:
      if (level < 1 || level >= D.K) return TOPO_EINVAL;
      v = (what == TOPO_X_LEVEL_B) ? h->levels[level].b : h->levels[level].x2;
      if (what == TOPO_X_LEVEL_X2) base = h->levels[level].x;
      npr = h->levels[level].nx + 1; ny = h->levels[level].ny;
      break;
    case TOPO_X_DIAG:
    case TOPO_X_FILTER:
      return TOPO_OK;                                                // consumed by topo_dist_setup_global / topo_dist_filter
    default:
      return TOPO_EINVAL;
  }
  const size_t last = (size_t)ny * npr;
  const int nb = (npr + 127) / 128;
  if (base == nullptr) {
    if (D.has_lo) KL(h, "k_row_add", k_row_add<<<nb, 128, 0, st>>>(v, D.recv_lo, npr, done));
    if (D.has_hi) KL(h, "k_row_add", k_row_add<<<nb, 128, 0, st>>>(v + last, D.recv_hi, npr, done));
  } else {
    const int f32 = (what == TOPO_X_Z) ? 1 : 0;                     // x~ is a float2 array: its last row starts at float2 index `last`
    const double2 *base_hi = f32 ? reinterpret_cast<const double2 *>(reinterpret_cast<const float2 *>(base) + last) : base + last;
    if (D.has_lo) KL(h, "k_row_add_sub", k_row_add_sub<<<nb, 128, 0, st>>>(v, D.recv_lo, base, f32, npr, done));
    if (D.has_hi) KL(h, "k_row_add_sub", k_row_add_sub<<<nb, 128, 0, st>>>(v + last, D.recv_hi, base_hi, f32, npr, done));
  }
  CUDA_TRY(cudaGetLastError());
  return TOPO_OK;
}

extern "C" int topo_dist_gather_ptrs(topo_t *h, int what, void **send, void **recv, long long *n_doubles_per_rank) {
  TOPO_TRY(check(h));
  if (!send || !recv || !n_doubles_per_rank) return TOPO_EINVAL;
  DistState &D = h->dist;
  if (what == TOPO_G_CELLS) {
    *send = D.cellK_send; *recv = D.cellK_recv; *n_doubles_per_rank = 64 * D.ncellK_loc;
  } else if (what == TOPO_G_BK) {
    *send = D.bK_send; *recv = D.bK_recv; *n_doubles_per_rank = 2 * D.bK_count;
  } else {
    return TOPO_EINVAL;
  }
  return TOPO_OK;
}

// device scalars the host all-reduces in place
extern "C" int topo_dist_scalar_ptr(topo_t *h, int what, void **ptr, int *n) {
  TOPO_TRY(check(h));
  if (!ptr || !n) return TOPO_EINVAL;
  switch (what) {
    case TOPO_S_PAP: *ptr = &h->sc->aux[0]; *n = 1; break;
    case TOPO_S_RR_RZ: *ptr = &h->sc->aux[1]; *n = 2; break;
    case TOPO_S_BB: *ptr = &h->sc->aux[3]; *n = 1; break;
    case TOPO_S_RED: *ptr = h->red_out; *n = 16; break;
    default: return TOPO_EINVAL;
  }
  return TOPO_OK;
}

extern "C" int topo_dist_read_scalars(topo_t *h, double *out, int n) {
  TOPO_TRY(check(h));
  if (!out || n < 1 || n > 16) return TOPO_EINVAL;
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaMemcpyAsync(out, h->red_out, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return TOPO_OK;
}

// ---- multigrid set-up ---------------------------------------------------------------------------------------
extern "C" int topo_dist_setup(topo_t *h, int phase) {
  TOPO_TRY(check(h));
  CUDA_TRY(cudaSetDevice(h->device));
  return phase == 0 ? topo_dist_setup_local(h) : topo_dist_setup_global(h);
}

// ---- PCG in segments ----------------------------------------------------------------------------------------
// buffer roles as in topo_solve (pcg.cu): parity `par` reads u/r/p of parity par and writes parity par^1
#define DIST_BUFS                                                                                 \
  double2 *ub[2] = {h->u, h->u2}, *rb[2] = {h->r, h->r2}, *pb[2] = {h->p, h->p2};                \
  double2 *xhat = h->tmp2, *res = h->tmp;                                                        \
  (void)ub; (void)rb; (void)pb; (void)xhat; (void)res

// r0 = W f - K_loc u (distributed; host: exchange TOPO_X_R0), ||f||^2 partial (host: all-reduce TOPO_S_BB)
extern "C" int topo_dist_solve_begin(topo_t *h, double rtol, int maxit, int warm) {
  TOPO_TRY(check(h));
  if (!h->loads_set || !h->cons_set) {
    topo_set_error("topo_dist_solve_begin: set BC flags and loads first");
    return TOPO_ESTATE;
  }
  if (!h->gmg_valid) {
    topo_set_error("topo_dist_solve_begin: multigrid set-up missing (topo_dist_setup phases 0 and 1)");
    return TOPO_ESTATE;
  }
  CUDA_TRY(cudaSetDevice(h->device));
  DIST_BUFS;
  cudaStream_t st = h->stream;
  const int64_t nn = h->nn;
  const size_t vbytes = sizeof(double2) * (size_t)nn;
  if (!warm) KL(h, "k_zero2", k_zero2d<<<nblk(nn), 256, 0, st>>>(h->u, nn));
  else TOPO_TRY(topo_mask_vector(h, h->u));
  SolverScalars init = {};
  init.tol2 = rtol * rtol;
  init.maxit = maxit;
  *h->sc_host = init;
  CUDA_TRY(cudaMemcpyAsync(h->sc, h->sc_host, sizeof(init), cudaMemcpyHostToDevice, st));
  TOPO_TRY(topo_fine_residual(h, ub[0], h->f, rb[0]));
  CUDA_TRY(cudaMemsetAsync(pb[1], 0, vbytes, st));
  CUDA_TRY(cudaMemsetAsync(h->Ap, 0, vbytes, st));
  KL(h, "k_dot_weighted", k_dot_weighted<<<nblk(nn), 256, 0, st>>>(h->f, h->f, h->nx + 1, h->ny, h->dist.w_first, h->dist.w_last,
                                                                  h->partials, h->ticket, &h->sc->aux[3]));
  CUDA_TRY(cudaGetLastError());
  return TOPO_OK;
}

// p' = z + beta p ; Ap = K_loc p' ; p'.Ap partial      (host: exchange TOPO_X_AP, all-reduce TOPO_S_PAP)
extern "C" int topo_dist_pcg_matvec(topo_t *h, int par) {
  TOPO_TRY(check(h));
  DIST_BUFS;
  return topo_fine_pcg_matvec(h, h->z, pb[par & 1], pb[(par & 1) ^ 1], h->Ap);
}

// u' = u + alpha p ; r' = r - alpha Ap ; rr partial ; x^ = w D^-1 r' ; res = W r' - K_loc x^ ; restrict res
// (host: K > 1: exchange TOPO_X_LEVEL_B level 1 ; K == 1: gather TOPO_G_BK).  init: the alpha = 0 pass that
// starts the solve (after TOPO_X_R0 / TOPO_S_BB).
extern "C" int topo_dist_pcg_update(topo_t *h, int par, int init) {
  TOPO_TRY(check(h));
  DIST_BUFS;
  par &= 1;
  if (init) {
    par = 0;
    KL(h, "k_dist_set_bb", k_dist_set_bb<<<1, 1, 0, h->stream>>>(h->sc));
  }
  TOPO_TRY(topo_fine_update_resid0(h, pb[par ^ 1], h->Ap, ub[par], rb[par], ub[par ^ 1], rb[par ^ 1], xhat, res, h->omega, init));
  return topo_dist_restrict_fine(h, res);
}

// level l (1 <= l < K) with b_l accumulated: pre-smoothing, residual, restriction
// (host: l+1 < K: exchange TOPO_X_LEVEL_B level l+1 ; else gather TOPO_G_BK)
extern "C" int topo_dist_vcycle_down(topo_t *h, int level) {
  TOPO_TRY(check(h));
  return topo_dist_level_down(h, level);
}

// global levels >= K from the gathered right-hand side (every rank, redundantly)
extern "C" int topo_dist_vcycle_tail(topo_t *h, int p2p) {
  TOPO_TRY(check(h));
  if (p2p && !h->dist.p2p_on) return TOPO_ESTATE;
  return topo_dist_tail(h, p2p);
}

// level l (K > l >= 1): coarse correction + post-smoothing (host: exchange TOPO_X_LEVEL_X2 level l)
extern "C" int topo_dist_vcycle_up(topo_t *h, int level) {
  TOPO_TRY(check(h));
  return topo_dist_level_up(h, level);
}

// x~ = x^ + P x_1 ; z = x~ + w D^-1 (W r - K_loc x~) ; r.z partial  (host: exchange TOPO_X_Z, all-reduce TOPO_S_RR_RZ)
extern "C" int topo_dist_pcg_smooth(topo_t *h, int par, int init) {
  TOPO_TRY(check(h));
  DIST_BUFS;
  par = init ? 0 : (par & 1);
  TOPO_TRY(topo_dist_prolong_fine(h, xhat));
  return topo_fine_prolong_smooth(h, xhat, rb[par ^ 1], nullptr, h->lvl_nx[1], h->z, h->omega, init);
}

// beta, ||r||^2, iteration count, convergence flag from the all-reduced scalars
extern "C" int topo_dist_pcg_finish(topo_t *h, int init) {
  TOPO_TRY(check(h));
  KL(h, "k_dist_pcg_finish", k_dist_pcg_finish<<<1, 1, 0, h->stream>>>(h->sc, init));
  CUDA_TRY(cudaGetLastError());
  return TOPO_OK;
}

extern "C" int topo_dist_solve_poll(topo_t *h, int *done, int *iters, double *relres) {
  TOPO_TRY(check(h));
  CUDA_TRY(cudaSetDevice(h->device));
  SolverScalars *hs = h->sc_host;
  CUDA_TRY(cudaMemcpyAsync(hs, h->sc, sizeof(SolverScalars), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (done) *done = hs->done;
  if (iters) *iters = hs->iters;
  if (relres) *relres = (hs->bb > 0.0) ? std::sqrt(hs->rr / hs->bb) : (hs->rr == 0.0 ? 0.0 : INFINITY);
  return TOPO_OK;
}

// the solution of an odd number of updates lives in u2: bring it home
extern "C" int topo_dist_solve_end(topo_t *h, int iters) {
  TOPO_TRY(check(h));
  CUDA_TRY(cudaSetDevice(h->device));
  if (((iters + 1) & 1) == 1)
    CUDA_TRY(cudaMemcpyAsync(h->u, h->u2, sizeof(double2) * (size_t)h->nn, cudaMemcpyDeviceToDevice, h->stream));
  return TOPO_OK;
}

// compliance pieces: red_out[1] = sum W f.u (partial)  -- red_out[0] is written by topo_dist_sensitivity
extern "C" int topo_dist_compliance_partial(topo_t *h) {
  TOPO_TRY(check(h));
  CUDA_TRY(cudaSetDevice(h->device));
  KL(h, "k_dot_weighted", k_dot_weighted<<<nblk(h->nn), 256, 0, h->stream>>>(h->f, h->u, h->nx + 1, h->ny, h->dist.w_first,
                                                                            h->dist.w_last, h->partials, h->ticket, h->red_out + 1));
  CUDA_TRY(cudaGetLastError());
  return TOPO_OK;
}

// =====================================================================================================================
// Peer-memory transport (NVLink / NVSwitch): the same exchanges, gathers and all-reduces without NCCL launches
// =====================================================================================================================
namespace {

PeerTable peer_table(const DistState &D);
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// bounded spin (~ seconds): a missing peer must not hang the GPU
__device__ __forceinline__ bool wait_flag(const unsigned long long *flag, unsigned long long seq, unsigned int *error) {
  if (*reinterpret_cast<volatile unsigned int *>(error)) return false;   // a previous wait already gave up
  for (long long it = 0; it < (1ll << 24); ++it) {
    if (ld_acquire_sys(flag) >= seq) return true;
    __nanosleep(64);
  }
  *error = 1u;
  return false;
}

__device__ __forceinline__ P2PHeader *hdr(char *region) { return reinterpret_cast<P2PHeader *>(region); }
__device__ __forceinline__ void rc_of_host(int64_t n, int npr, int &r, int &c) {
  const unsigned un = (unsigned)n, q = un / (unsigned)npr;
  r = (int)q;
  c = (int)(un - q * (unsigned)npr);
}

// my first node row -> lower neighbour's recv_hi, my last node row -> upper neighbour's recv_lo, then the flags
__global__ void __launch_bounds__(1024) k_p2p_push_rows(const double2 *__restrict__ row_lo, const double2 *__restrict__ row_hi,
                                                        int n, char *peer_lo, char *peer_hi, size_t off_lo, size_t off_hi,
                                                        size_t stride, P2PCounters *c) {
  const unsigned long long seq = c->halo_push + 1;
  const size_t buf = (size_t)(seq & 1ull) * stride;
  if (peer_lo != nullptr) {
    double2 *dst = reinterpret_cast<double2 *>(peer_lo + off_hi + buf);
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = row_lo[i];
  }
  if (peer_hi != nullptr) {
    double2 *dst = reinterpret_cast<double2 *>(peer_hi + off_lo + buf);
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = row_hi[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (peer_lo != nullptr) st_release_sys(&hdr(peer_lo)->flag_from_hi, seq);
    if (peer_hi != nullptr) st_release_sys(&hdr(peer_hi)->flag_from_lo, seq);
    c->halo_push = seq;
  }
}

// block 0: first row, block 1: last row.  dst += recv  or  dst = dst + recv - base
__global__ void __launch_bounds__(1024) k_p2p_finish_rows(double2 *__restrict__ v_lo, double2 *__restrict__ v_hi,
                                                          const double2 *__restrict__ base_lo, const double2 *__restrict__ base_hi,
                                                          int base_f32, int n, char *region, int has_lo, int has_hi, size_t off_lo, size_t off_hi,
                                                          size_t stride, P2PCounters *c) {
  const unsigned long long seq = c->halo_wait + 1;
  const size_t buf = (size_t)(seq & 1ull) * stride;
  const bool lo = blockIdx.x == 0;
  if (lo ? has_lo : has_hi) {
    __shared__ bool ok;
    if (threadIdx.x == 0) ok = wait_flag(lo ? &hdr(region)->flag_from_lo : &hdr(region)->flag_from_hi, seq, &c->error);
    __syncthreads();
    if (ok) {
      const double2 *src = reinterpret_cast<const double2 *>(region + (lo ? off_lo : off_hi) + buf);
      double2 *dst = lo ? v_lo : v_hi;
      const double2 *base = lo ? base_lo : base_hi;
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double2 a = dst[i];
        const double2 b = src[i];
        if (base != nullptr) {
          double2 q;
          if (base_f32) {
            const float2 f = reinterpret_cast<const float2 *>(base)[i];
            q = make_double2((double)f.x, (double)f.y);
          } else {
            q = base[i];
          }
          a.x = (a.x + b.x) - q.x; a.y = (a.y + b.y) - q.y;
        } else {
          a.x += b.x; a.y += b.y;
        }
        dst[i] = a;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&c->ticket[0], 1u) == gridDim.x - 1) {
      c->ticket[0] = 0u;
      c->halo_wait = seq;
    }
  }
}

// block r: my level-K piece -> rank r's gather buffer, slot [my rank], then the flag
__global__ void __launch_bounds__(1024) k_p2p_gather_push(const double2 *__restrict__ src, long long n, PeerTable peers, int my_rank,
                                                          size_t off_gather, size_t stride, P2PCounters *c) {
  const unsigned long long seq = c->gather_push + 1;
  char *peer = peers.p[blockIdx.x];
  double2 *dst = reinterpret_cast<double2 *>(peer + off_gather + (size_t)(seq & 1ull) * stride) + (size_t)my_rank * n;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    st_release_sys(&hdr(peer)->flag_gather[my_rank], seq);
    __threadfence();
    if (atomicAdd(&c->ticket[1], 1u) == gridDim.x - 1) {
      c->ticket[1] = 0u;
      c->gather_push = seq;
    }
  }
}
// thread r waits for rank r's piece
__global__ void k_p2p_gather_wait(char *region, int world, P2PCounters *c) {
  const unsigned long long seq = c->gather_wait + 1;
  if ((int)threadIdx.x < world) wait_flag(&hdr(region)->flag_gather[threadIdx.x], seq, &c->error);
  __syncthreads();
  if (threadIdx.x == 0) c->gather_wait = seq;
}

// thread r: my n partial sums -> rank r's slot [parity][my rank], then the flag
__global__ void k_p2p_scalar_push(const double *__restrict__ vals, int n, PeerTable peers, int my_rank, int world, P2PCounters *c) {
  const unsigned long long seq = c->scal_push + 1;
  if ((int)threadIdx.x < world) {
    P2PHeader *H = hdr(peers.p[threadIdx.x]);
    double *dst = H->scalar_slot[seq & 1ull][my_rank];
    for (int j = 0; j < n; ++j) dst[j] = vals[j];
    __threadfence_system();
    st_release_sys(&H->flag_scalar[my_rank], seq);
  }
  __syncthreads();
  if (threadIdx.x == 0) c->scal_push = seq;
}
// waits for every rank's partial sums, folds them in rank order (identical on all ranks), writes them back
__global__ void k_p2p_scalar_reduce(double *__restrict__ vals, int n, char *region, int world, int is_max, P2PCounters *c) {
  const unsigned long long seq = c->scal_wait + 1;
  if ((int)threadIdx.x < world) wait_flag(&hdr(region)->flag_scalar[threadIdx.x], seq, &c->error);
  __syncthreads();
  if ((int)threadIdx.x < n) {
    const P2PHeader *H = hdr(region);
    double acc = H->scalar_slot[seq & 1ull][0][threadIdx.x];
    for (int r = 1; r < world; ++r) {
      const double t = H->scalar_slot[seq & 1ull][r][threadIdx.x];
      acc = is_max ? ((t > acc || t != t) ? t : acc) : acc + t;
    }
    vals[threadIdx.x] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) c->scal_wait = seq;
}

// ---- one-launch exchanges (one slab per process: a kernel may wait for its peers because every peer pushes before it
// waits; with several slabs of one process on one stream the two-phase kernels above are the only safe form) -------------
// Measured on 8 GPUs (scripts/slab_variants.py, 50 exchanges per CUDA graph): an exchange built from "payload, fence.sys,
// release-store of a sequence number / acquire-poll, read payload" costs 13-19 us even for a 4 KB row -- thirteen of them
// were a third of a PCG iteration.  Here every double travels as ONE 16-byte store (lo, seq, hi, seq) -- each 8-byte half
// carries its own copy of the sequence number, the protocol of NCCL's LL primitives -- so there is no fence, no separate
// flag and no ticket on the critical path: a thread pushes the values of its node, polls its own slot until both halves
// show the expected number, and completes its node.  Slots are double-buffered by the parity of the sequence number and
// a stale slot holds an OLDER number, so equality is the only test needed.
// Blocks [0, nb): my first row <-> rank-1; [nb, 2nb): my last row <-> rank+1, one node per thread; block 2nb (if
// nvals > 0): all-reduce of nvals scalars, folded in rank order on every rank (+ the PCG bookkeeping that follows the
// (rr, rz) all-reduce).
struct FusedXchg {
  double2 *v_lo, *v_hi;
  const double2 *base_lo, *base_hi;
  int base_f32, n, nb;
  char *region, *peer_lo, *peer_hi;
  size_t off_lo, off_hi, stride;           // flag-in-data row slots: [parity][n] x 2 doubles x 16 B
  P2PCounters *c;
  double *vals;
  int nvals, my_rank, world, finish;       // finish: 0 none, 1 PCG bookkeeping, 2 the same for the initial pass, 3: max instead of sum
  SolverScalars *sc;
};
__device__ __forceinline__ void pcg_bookkeeping(SolverScalars *sc, int init) {
  if (!init && sc->done) return;
  const double rr = sc->aux[1], rz = sc->aux[2];
  sc->rr = rr;
  if (!init) sc->iters += 1;
  sc->rz[1] = init ? 0.0 : rz / sc->rz[0];
  sc->rz[0] = rz;
  if (rr <= sc->tol2 * sc->bb || !(rr == rr)) sc->done = 1;
  else if (sc->iters >= sc->maxit) sc->done = 2;
}
__global__ void __launch_bounds__(256) k_p2p_fused(const __grid_constant__ FusedXchg a, const __grid_constant__ PeerTable peers) {
  P2PCounters *c = a.c;
  const int t = threadIdx.x;
  if ((int)blockIdx.x == 2 * a.nb) {                       // ---- scalars ----
    const unsigned long long seq = c->scal_push + 1;
    const unsigned s32 = (unsigned)seq;
    double mine[16];
    if (t < a.world)
      for (int j = 0; j < a.nvals; ++j) mine[j] = a.vals[j];
    __syncthreads();                                       // every thread holds my values before anybody overwrites them
    if (t < a.world) {
      uint4 *dst = hdr(peers.p[t])->ll_scalar[seq & 1ull][a.my_rank];
      for (int j = 0; j < a.nvals; ++j) ll_store(dst + j, mine[j], s32);
    }
    if (t < a.nvals) {
      const P2PHeader *H = hdr(a.region);
      double acc = 0.0;
      bool ok = true;
      for (int r = 0; r < a.world && ok; ++r) {
        double q;
        ok = ll_load(&H->ll_scalar[seq & 1ull][r][t], s32, q, &c->error);
        acc = r == 0 ? q : (a.finish == 3 ? ((q > acc || q != q) ? q : acc) : acc + q);
      }
      if (ok) a.vals[t] = acc;
    }
    __syncthreads();
    if (t == 0) {
      if (a.finish == 1 || a.finish == 2) pcg_bookkeeping(a.sc, a.finish == 2);
      c->scal_push = seq;
      c->scal_wait = seq;
    }
    return;
  }
  // ---- rows ----
  const RowXchg x = {a.region, a.peer_lo, a.peer_hi, a.off_lo, a.off_hi, a.stride, a.c};
  const unsigned long long seq = ll_rows_seq(x);
  const bool lo = (int)blockIdx.x < a.nb;
  const int i = (lo ? blockIdx.x : blockIdx.x - a.nb) * blockDim.x + t;
  if ((lo ? a.peer_lo : a.peer_hi) != nullptr && i < a.n) {
    double2 *v = lo ? a.v_lo : a.v_hi;
    double2 mine = v[i], theirs;
    const double2 *base = lo ? a.base_lo : a.base_hi;
    double2 q = make_double2(0.0, 0.0);
    if (base != nullptr) {
      if (a.base_f32) {
        const float2 f = reinterpret_cast<const float2 *>(base)[i];
        q = make_double2((double)f.x, (double)f.y);
      } else {
        q = base[i];
      }
    }
    if (ll_row_swap(x, lo, i, seq, mine, theirs)) {
      if (base != nullptr) { mine.x = (mine.x + theirs.x) - q.x; mine.y = (mine.y + theirs.y) - q.y; }
      else { mine.x += theirs.x; mine.y += theirs.y; }
      v[i] = mine;
    }
  }
  __syncthreads();
  if (t == 0) ll_rows_done(x, seq, 2u * (unsigned)a.nb);
}

// level-K pieces to every rank, wait for everybody's, assemble the global right-hand side -- one launch.
// Blocks [0, world) push first (block r -> rank r); every block then assembles its nodes, polling the values it needs.
__global__ void __launch_bounds__(256) k_p2p_gather_fused(const double2 *__restrict__ src, long long n, const __grid_constant__ PeerTable peers,
                                                          int my_rank, int world, char *region, size_t off_gather, size_t stride,
                                                          P2PCounters *c, double2 *__restrict__ b, int npr, int rows) {
  const unsigned long long seq = c->gather_push + 1;
  const unsigned s32 = (unsigned)seq;
  const int t = threadIdx.x, T = blockDim.x;
  if ((int)blockIdx.x < world) {
    uint4 *dst = reinterpret_cast<uint4 *>(peers.p[blockIdx.x] + off_gather + (size_t)(seq & 1ull) * stride) + 2 * (size_t)my_rank * n;
    for (long long i = t; i < n; i += T) {
      const double2 x = src[i];
      ll_store(dst + 2 * i, x.x, s32);
      ll_store(dst + 2 * i + 1, x.y, s32);
    }
  }
  const uint4 *__restrict__ recv = reinterpret_cast<const uint4 *>(region + off_gather + (size_t)(seq & 1ull) * stride);
  const int64_t nn = (int64_t)npr * ((int64_t)rows * world + 1);
  const int64_t per = (int64_t)(rows + 1) * npr;
  for (int64_t m = (int64_t)blockIdx.x * T + t; m < nn; m += (int64_t)gridDim.x * T) {
    int J, I;
    rc_of_host(m, npr, J, I);
    const int g = J / rows, j = J - g * rows;
    double2 acc = make_double2(0.0, 0.0);
    bool ok = true;
    if (g < world) {
      const uint4 *w = recv + 2 * ((int64_t)g * per + (int64_t)j * npr + I);
      ok = ll_load(w, s32, acc.x, &c->error) && ll_load(w + 1, s32, acc.y, &c->error);
    }
    if (ok && j == 0 && g > 0) {
      const uint4 *w = recv + 2 * ((int64_t)(g - 1) * per + (int64_t)rows * npr + I);
      double qx, qy;
      ok = ll_load(w, s32, qx, &c->error) && ll_load(w + 1, s32, qy, &c->error);
      acc.x += qx; acc.y += qy;
    }
    if (ok) b[m] = acc;
  }
  __syncthreads();
  if (t == 0) {
    __threadfence();
    if (atomicAdd(&c->fticket[3], 1u) == gridDim.x - 1u) {
      c->fticket[3] = 0u;
      c->gather_push = seq;
      c->gather_wait = seq;
    }
  }
}

}  // namespace
PeerTable topo_dist_peer_table(const topo_handle *h) { return peer_table(h->dist); }
RowXchg topo_dist_row_xchg(const topo_handle *h) {
  RowXchg x = {};
  const DistState &D = h->dist;
  if (!D.on || !D.p2p_on || !D.p2p_inkernel) return x;
  x.region = D.p2p_region;
  x.peer_lo = D.has_lo ? D.p2p_peer[D.rank - 1] : nullptr;
  x.peer_hi = D.has_hi ? D.p2p_peer[D.rank + 1] : nullptr;
  x.off_lo = D.p2p_off_ll_lo; x.off_hi = D.p2p_off_ll_hi; x.stride = 2 * D.p2p_row_stride;
  x.c = D.p2p_cnt;
  return x;
}
namespace {

int p2p_check(topo_handle *h) {
  TOPO_TRY(check(h));
  if (!h->dist.p2p_on) {
    topo_set_error("peer-memory transport not attached (topo_dist_p2p_attach)");
    return TOPO_ESTATE;
  }
  return TOPO_OK;
}

PeerTable peer_table(const DistState &D) {
  PeerTable t;
  for (int r = 0; r < TOPO_P2P_MAXW; ++r) t.p[r] = D.p2p_peer[r];
  return t;
}

}  // namespace

// 64-byte CUDA IPC handle of this slab's region (for peers in other processes) and its address (for peers in this one)
extern "C" int topo_dist_p2p_export(topo_t *h, void *ipc_handle_64, void **local_ptr) {
  TOPO_TRY(check(h));
  CUDA_TRY(cudaSetDevice(h->device));
  if (ipc_handle_64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t mh;
    CUDA_TRY(cudaIpcGetMemHandle(&mh, h->dist.p2p_region));
    std::memcpy(ipc_handle_64, &mh, 64);
  }
  if (local_ptr) *local_ptr = h->dist.p2p_region;
  return TOPO_OK;
}

// regions of ALL ranks: local_ptrs[r] != NULL -> a handle of this process (used as is), else ipc_handles + 64*r is
// opened (peer access over NVLink).  Entry [own rank] is ignored.
extern "C" int topo_dist_p2p_attach(topo_t *h, const void *ipc_handles, void *const *local_ptrs) {
  TOPO_TRY(check(h));
  CUDA_TRY(cudaSetDevice(h->device));
  DistState &D = h->dist;
  if (D.world > TOPO_P2P_MAXW) {
    topo_set_error("peer-memory transport supports at most %d slabs", TOPO_P2P_MAXW);
    return TOPO_EINVAL;
  }
  for (int r = 0; r < D.world; ++r) {
    if (r == D.rank) { D.p2p_peer[r] = D.p2p_region; continue; }
    if (local_ptrs && local_ptrs[r]) { D.p2p_peer[r] = static_cast<char *>(local_ptrs[r]); continue; }
    if (!ipc_handles) return TOPO_EINVAL;
    cudaIpcMemHandle_t mh;
    std::memcpy(&mh, static_cast<const char *>(ipc_handles) + 64 * (size_t)r, 64);
    void *p = nullptr;
    CUDA_TRY(cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess));
    D.p2p_peer[r] = static_cast<char *>(p);
    D.p2p_opened[r] = true;
  }
  D.p2p_on = true;
  return TOPO_OK;
}

// exchange through peer memory.  phase 0: push my interface rows (never waits); phase 1: wait for the neighbours'
// rows and complete mine.  Row kinds only (TOPO_X_AP, _R0, _Z, _LEVEL_B, _LEVEL_X2).
extern "C" int topo_dist_p2p_xchg(topo_t *h, int phase, int what, int level) {
  TOPO_TRY(p2p_check(h));
  CUDA_TRY(cudaSetDevice(h->device));
  DistState &D = h->dist;
  double2 *v = nullptr;
  const double2 *base = nullptr;
  int npr = h->nx + 1, ny = h->ny;
  switch (what) {
    case TOPO_X_AP: v = h->Ap; break;
    case TOPO_X_R0: v = h->r; break;
    case TOPO_X_Z: v = h->z; base = h->tmp2; break;
    case TOPO_X_LEVEL_B:
    case TOPO_X_LEVEL_X2:
      if (level < 1 || level >= D.K) return TOPO_EINVAL;
      v = (what == TOPO_X_LEVEL_B) ? h->levels[level].b : h->levels[level].x2;
      if (what == TOPO_X_LEVEL_X2) base = h->levels[level].x;
      npr = h->levels[level].nx + 1; ny = h->levels[level].ny;
      break;
    default:
      topo_set_error("peer-memory exchange: kind %d is not a node-row exchange", what);
      return TOPO_EINVAL;
  }
  const size_t last = (size_t)ny * npr;
  char *peer_lo = D.has_lo ? D.p2p_peer[D.rank - 1] : nullptr;
  char *peer_hi = D.has_hi ? D.p2p_peer[D.rank + 1] : nullptr;
  if (phase == 0) {
    KL(h, "k_p2p_push_rows", k_p2p_push_rows<<<1, 1024, 0, h->stream>>>(v, v + last, npr, peer_lo, peer_hi, D.p2p_off_lo, D.p2p_off_hi,
                                                                        D.p2p_row_stride, D.p2p_cnt));
  } else {
    const int f32 = (what == TOPO_X_Z) ? 1 : 0;
    const double2 *base_hi = !base ? nullptr
                             : (f32 ? reinterpret_cast<const double2 *>(reinterpret_cast<const float2 *>(base) + last) : base + last);
    KL(h, "k_p2p_finish_rows", k_p2p_finish_rows<<<2, 1024, 0, h->stream>>>(v, v + last, base, base_hi, f32, npr,
                                                                            D.p2p_region, D.has_lo ? 1 : 0, D.has_hi ? 1 : 0,
                                                                            D.p2p_off_lo, D.p2p_off_hi, D.p2p_row_stride, D.p2p_cnt));
  }
  CUDA_TRY(cudaGetLastError());
  return TOPO_OK;
}

// level-K right-hand-side pieces to every rank (phase 0) / wait for all pieces (phase 1; then topo_dist_vcycle_tail(h, 1))
extern "C" int topo_dist_p2p_gather(topo_t *h, int phase) {
  TOPO_TRY(p2p_check(h));
  CUDA_TRY(cudaSetDevice(h->device));
  DistState &D = h->dist;
  if (phase == 0) {
    KL(h, "k_p2p_gather_push", k_p2p_gather_push<<<D.world, 1024, 0, h->stream>>>(D.bK_send, D.bK_count, peer_table(D), D.rank,
                                                                                D.p2p_off_gather, D.p2p_gather_stride, D.p2p_cnt));
  } else {
    KL(h, "k_p2p_gather_wait", k_p2p_gather_wait<<<1, 32, 0, h->stream>>>(D.p2p_region, D.world, D.p2p_cnt));
  }
  CUDA_TRY(cudaGetLastError());
  return TOPO_OK;
}

// all-reduce of the first n doubles of a scalar block (TOPO_S_*), sum or max, folded in rank order on every rank
extern "C" int topo_dist_p2p_allreduce(topo_t *h, int phase, int what, int n, int is_max) {
  TOPO_TRY(p2p_check(h));
  CUDA_TRY(cudaSetDevice(h->device));
  DistState &D = h->dist;
  void *ptr = nullptr;
  int nmax = 0;
  TOPO_TRY(topo_dist_scalar_ptr(h, what, &ptr, &nmax));
  if (n < 1 || n > nmax || n > 16) return TOPO_EINVAL;
  if (phase == 0) {
    KL(h, "k_p2p_scalar_push", k_p2p_scalar_push<<<1, 32, 0, h->stream>>>(static_cast<const double *>(ptr), n, peer_table(D), D.rank, D.world,
                                                                        D.p2p_cnt));
  } else {
    KL(h, "k_p2p_scalar_reduce", k_p2p_scalar_reduce<<<1, 32, 0, h->stream>>>(static_cast<double *>(ptr), n, D.p2p_region, D.world, is_max,
                                                                            D.p2p_cnt));
  }
  CUDA_TRY(cudaGetLastError());
  return TOPO_OK;
}

int topo_dist_tail_after_assemble(topo_handle *h);

// one-launch exchange (see k_p2p_fused): interface rows of `what` (+ level), optionally the all-reduce (sum) of the first
// scal_n doubles of scalar block scal_what (scal_n = 0: none) and the PCG bookkeeping (finish = 1; 2 for the initial pass)
extern "C" int topo_dist_p2p_fused(topo_t *h, int what, int level, int scal_what, int scal_n, int finish) {
  TOPO_TRY(p2p_check(h));
  CUDA_TRY(cudaSetDevice(h->device));
  DistState &D = h->dist;
  double2 *v = nullptr;
  const double2 *base = nullptr;
  int npr = h->nx + 1, ny = h->ny;
  switch (what) {
    case -1: npr = 0; break;                                 // scalars only
    case TOPO_X_AP: v = h->Ap; break;
    case TOPO_X_R0: v = h->r; break;
    case TOPO_X_Z: v = h->z; base = h->tmp2; break;
    case TOPO_X_LEVEL_B:
    case TOPO_X_LEVEL_X2:
      if (level < 1 || level >= D.K) return TOPO_EINVAL;
      v = (what == TOPO_X_LEVEL_B) ? h->levels[level].b : h->levels[level].x2;
      if (what == TOPO_X_LEVEL_X2) base = h->levels[level].x;
      npr = h->levels[level].nx + 1; ny = h->levels[level].ny;
      break;
    default:
      topo_set_error("peer-memory exchange: kind %d is not a node-row exchange", what);
      return TOPO_EINVAL;
  }
  if (npr == 0 && scal_n < 1) return TOPO_EINVAL;
  const size_t last = (size_t)ny * npr;
  FusedXchg a = {};
  a.v_lo = v; a.v_hi = v + last;
  a.base_f32 = (what == TOPO_X_Z) ? 1 : 0;
  a.base_lo = base;
  a.base_hi = !base ? nullptr : (a.base_f32 ? reinterpret_cast<const double2 *>(reinterpret_cast<const float2 *>(base) + last) : base + last);
  a.n = npr;
  a.nb = (npr + 255) / 256;                              // one node per thread
  a.region = D.p2p_region;
  a.peer_lo = D.has_lo ? D.p2p_peer[D.rank - 1] : nullptr;
  a.peer_hi = D.has_hi ? D.p2p_peer[D.rank + 1] : nullptr;
  a.off_lo = D.p2p_off_ll_lo; a.off_hi = D.p2p_off_ll_hi; a.stride = 2 * D.p2p_row_stride;
  a.c = D.p2p_cnt;
  a.my_rank = D.rank; a.world = D.world; a.finish = finish; a.sc = h->sc;
  if (scal_n > 0) {
    void *ptr = nullptr;
    int nmax = 0;
    TOPO_TRY(topo_dist_scalar_ptr(h, scal_what, &ptr, &nmax));
    if (scal_n > nmax || scal_n > 16) return TOPO_EINVAL;
    a.vals = static_cast<double *>(ptr);
    a.nvals = scal_n;
  }
  const int blocks = 2 * a.nb + (a.nvals > 0 ? 1 : 0);
  KL(h, "k_p2p_fused", k_p2p_fused<<<blocks, 256, 0, h->stream>>>(a, peer_table(D)));
  CUDA_TRY(cudaGetLastError());
  return TOPO_OK;
}

// in-kernel exchanges on (1) / off (0): the restriction kernels of topo_dist_pcg_update / topo_dist_vcycle_down and the up
// kernel of topo_dist_vcycle_up swap their interface rows with the neighbours themselves (flag-in-data slots), so the host
// issues NO TOPO_X_LEVEL_B / TOPO_X_LEVEL_X2 exchange.  One slab per process only (see topo_dist_p2p_fused).
int topo_dist_mid_plan(topo_handle *h, int enable);
extern "C" int topo_dist_p2p_mode(topo_t *h, int mode) {
  TOPO_TRY(p2p_check(h));
  h->dist.p2p_inkernel = (mode & 1) != 0;
  // bit 1: the small slab-local levels (+ the level-K gather) in one persistent kernel per direction
  return topo_dist_mid_plan(h, (mode & 3) == 3);
}

// one-launch gather of the level-K right-hand side + assembly, then the tail kernel (levels >= K on every rank)
extern "C" int topo_dist_p2p_gather_tail(topo_t *h) {
  TOPO_TRY(p2p_check(h));
  CUDA_TRY(cudaSetDevice(h->device));
  DistState &D = h->dist;
  if (D.mid_from > 0) return topo_dist_tail_after_assemble(h);      // (k_slab_mid_down pushed, gathered and assembled)
  GmgLevel &G = h->levels[D.K];
  const int rows = (h->ny >> D.K);
  const int blocks = std::max(D.world, (int)std::min<int64_t>((G.nn + 255) / 256, 64));
  KL(h, "k_p2p_gather_fused", k_p2p_gather_fused<<<blocks, 256, 0, h->stream>>>(D.bK_send, D.bK_count, peer_table(D), D.rank, D.world,
                                                                              D.p2p_region, D.p2p_off_ll_gather, 2 * D.p2p_gather_stride,
                                                                              D.p2p_cnt, G.b, G.nx + 1, rows));
  CUDA_TRY(cudaGetLastError());
  return topo_dist_tail_after_assemble(h);
}

// 1 if a peer wait timed out since the handle was created
extern "C" int topo_dist_p2p_error(topo_t *h, int *error) {
  TOPO_TRY(check(h));
  if (!error) return TOPO_EINVAL;
  CUDA_TRY(cudaSetDevice(h->device));
  P2PCounters c;
  CUDA_TRY(cudaMemcpyAsync(&c, h->dist.p2p_cnt, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  *error = (int)c.error;
  return TOPO_OK;
}
