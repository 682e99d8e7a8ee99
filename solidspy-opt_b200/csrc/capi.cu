// extern "C" glue of libtopo_b200.so: life cycle, state transfer, the fused SIMP step and timing.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <new>

#include "topo_internal.cuh"

int topo_mask_vector(topo_handle *h, double2 *v);
int topo_dot(topo_handle *h, const double2 *a, const double2 *b, double *host_out);

static thread_local char g_err[512] = "";

void topo_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char *topo_last_error(void) { return g_err; }
extern "C" int topo_abi_version(void) { return TOPO_ABI_VERSION; }

extern "C" int topo_device_count(int *count) {
  if (!count) return TOPO_EINVAL;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    *count = 0;
    topo_set_error("no CUDA device: %s", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    cudaGetLastError();
    return TOPO_ENODEV;
  }
  *count = n;
  return TOPO_OK;
}

namespace {

__global__ void k_fill(double *p, int64_t n, double v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void k_fill8(uint8_t *p, int64_t n, uint8_t v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}
// reference BC flags (-1 constrained / 0 free, two per node) <-> device bitmask
__global__ void k_cons_to_fixed(const int8_t *cons, uint8_t *fixed, int64_t nn) {
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < nn; n += (int64_t)gridDim.x * blockDim.x)
    fixed[n] = (cons[2 * n] != 0 ? 1u : 0u) | (cons[2 * n + 1] != 0 ? 2u : 0u);
}
__global__ void k_fixed_to_cons(const uint8_t *fixed, int8_t *cons, int64_t nn) {
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < nn; n += (int64_t)gridDim.x * blockDim.x) {
    const unsigned m = fixed[n];
    cons[2 * n] = (m & 1u) ? -1 : 0;
    cons[2 * n + 1] = (m & 2u) ? -1 : 0;
  }
}

inline int nblk(int64_t n) { return (int)std::min<int64_t>((n + 255) / 256, 148 * 8); }

// warm start by extrapolation along the design trajectory: u <- u + theta (u - u_prev), u_prev <- u (have = 0: copy only)
__global__ void k_extrapolate(double2 *__restrict__ u, double2 *__restrict__ up, int64_t nn, double theta, int have) {
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < nn; n += (int64_t)gridDim.x * blockDim.x) {
    const double2 a = u[n], b = up[n];
    up[n] = a;
    if (have) u[n] = make_double2(fma(theta, a.x - b.x, a.x), fma(theta, a.y - b.y, a.y));
  }
}

struct FieldRef {
  void *ptr;
  size_t bytes;
};

int field_ref(topo_handle *h, int field, FieldRef *out) {
  const size_t nv = sizeof(double2) * (size_t)h->nn, ev = sizeof(double) * (size_t)h->ne;
  switch (field) {
    case TOPO_F_U: *out = {h->u, nv}; return TOPO_OK;
    case TOPO_F_LOAD: *out = {h->f, nv}; return TOPO_OK;
    case TOPO_F_RESID: *out = {h->tmp, nv}; return TOPO_OK;
    case TOPO_F_RHO: *out = {h->rho, ev}; return TOPO_OK;
    case TOPO_F_DC: *out = {h->dc, ev}; return TOPO_OK;
    case TOPO_F_SCALE: *out = {h->scale, ev}; return TOPO_OK;
    case TOPO_F_SENS: *out = {h->sens, ev}; return TOPO_OK;
    case TOPO_F_SENS_PREV: *out = {h->sens_prev, ev}; return TOPO_OK;
    case TOPO_F_KEEP: *out = {h->keep, (size_t)h->ne}; return TOPO_OK;
    case TOPO_F_NODAL: *out = {h->nodal, sizeof(double) * (size_t)h->nn}; return TOPO_OK;
    case TOPO_F_CONS: *out = {nullptr, 2 * (size_t)h->nn}; return TOPO_OK;
    default:
      topo_set_error("unknown field id %d", field);
      return TOPO_EINVAL;
  }
}

}  // namespace

extern "C" int topo_create(topo_t **out, int nx, int ny, const double ke[64], int device) {
  if (!out || !ke || nx < 1 || ny < 1) {
    topo_set_error("topo_create: bad arguments (nx=%d, ny=%d)", nx, ny);
    return TOPO_EINVAL;
  }
  if ((int64_t)(nx + 1) * (ny + 1) * 2 >= (int64_t)1 << 31) {
    topo_set_error("topo_create: mesh too large for 32-bit DOF indices on one GPU");
    return TOPO_EINVAL;
  }
  {
    // The stage-1 kernel derives the row blocks of local nodes 1..3 from that of node 0 through the
    // mirror symmetries of a rectangular, isotropic element (x-mirror: nodes [1,0,3,2], ux -> -ux;
    // y-mirror: nodes [3,2,1,0], uy -> -uy).  Every matrix the reference builds on rect_grid meshes has
    // them (closed form optimize.py:408-416; Gauss elast_quad4 on rectangles); anything else is refused.
    const int px[4] = {1, 0, 3, 2}, py[4] = {3, 2, 1, 0};
    double kmax = 0.0, dev = 0.0;
    for (int i = 0; i < 64; ++i) kmax = fmax(kmax, fabs(ke[i]));
    for (int a = 0; a < 4; ++a)
      for (int d = 0; d < 2; ++d)
        for (int b = 0; b < 4; ++b)
          for (int e = 0; e < 2; ++e) {
            const double v = ke[(2 * a + d) * 8 + 2 * b + e];
            const double sx = ((d == 0) ? -1.0 : 1.0) * ((e == 0) ? -1.0 : 1.0);
            const double sy = ((d == 1) ? -1.0 : 1.0) * ((e == 1) ? -1.0 : 1.0);
            dev = fmax(dev, fabs(ke[(2 * px[a] + d) * 8 + 2 * px[b] + e] - sx * v));
            dev = fmax(dev, fabs(ke[(2 * py[a] + d) * 8 + 2 * py[b] + e] - sy * v));
            dev = fmax(dev, fabs(ke[(2 * b + e) * 8 + 2 * a + d] - v));
          }
    if (!(kmax > 0.0) || dev > 1e-9 * kmax) {
      topo_set_error("topo_create: element matrix is not that of a rectangular isotropic quad4 "
                     "(symmetry defect %.3e of max entry %.3e)", dev, kmax);
      return TOPO_EINVAL;
    }
  }
  int ndev = 0;
  TOPO_TRY(topo_device_count(&ndev));
  if (device < 0 || device >= ndev) {
    topo_set_error("topo_create: device %d not in [0,%d)", device, ndev);
    return TOPO_ENODEV;
  }
  CUDA_TRY(cudaSetDevice(device));
  topo_handle *h = new (std::nothrow) topo_handle();
  if (!h) return TOPO_EINVAL;
  h->device = device;
  h->nx = nx;
  h->ny = ny;
  h->nn = (int64_t)(nx + 1) * (ny + 1);
  h->ne = (int64_t)nx * ny;
  std::memcpy(h->ke.k, ke, sizeof(double) * 64);
  {
    // ke = sum_k l_k l_k^T by diagonally pivoted outer-product Cholesky (ke is PSD of rank 5: three rigid-body
    // modes).  The sensitivity kernel evaluates u^T ke u as sum_k (l_k . u)^2: never negative, and the rigid
    // part of u cancels inside each 8-term dot product instead of between 64 large terms.
    double A[64];
    std::memcpy(A, ke, sizeof(A));
    std::memset(h->ke_factor.k, 0, sizeof(h->ke_factor.k));
    double dmax = 0.0;
    for (int i = 0; i < 8; ++i) dmax = fmax(dmax, A[i * 8 + i]);
    bool used[8] = {false, false, false, false, false, false, false, false};
    for (int k = 0; k < 8; ++k) {
      int p = -1;
      for (int i = 0; i < 8; ++i)
        if (!used[i] && (p < 0 || A[i * 8 + i] > A[p * 8 + p])) p = i;
      if (p < 0 || !(A[p * 8 + p] > 1e-10 * dmax)) break;
      used[p] = true;
      const double inv = 1.0 / std::sqrt(A[p * 8 + p]);
      double *l = h->ke_factor.k + 8 * k;
      for (int i = 0; i < 8; ++i) l[i] = A[i * 8 + p] * inv;
      for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 8; ++j) A[i * 8 + j] -= l[i] * l[j];
    }
  }
  {
    auto env_i = [](const char *n, long long d) { const char *v = getenv(n); return v ? atoll(v) : d; };
    TopoKnobs &k = h->knobs;
    k.fine_waves = getenv("TOPO_FINE_WAVES") ? atof(getenv("TOPO_FINE_WAVES")) : 1.0;
    k.fine_nc_mask = (int)env_i("TOPO_FINE_NC_MASK", 0);
    k.sym_min = env_i("TOPO_SYM_MIN", 100000);
    k.nu_coarse_maxlvl = (int)env_i("TOPO_NU_COARSE_MAXLVL", 99);
    k.tail_cluster_max = (int)env_i("TOPO_TAIL_CLUSTER", 16);
    k.unfused_min = (int)env_i("TOPO_UNFUSED_MIN", 200000);
    k.fuse_prolong = getenv("TOPO_FUSE_PROLONG") ? 1 : 0;
    k.oc_ctas_per_sm = (int)std::max<long long>(1, env_i("TOPO_OC_CTAS_PER_SM", 2));
    k.oc_no_probe = getenv("TOPO_OC_NO_PROBE") ? 1 : 0;
    k.coarsest_nodes = (int)env_i("TOPO_COARSEST_NODES", 48);
    k.no_l2persist = getenv("TOPO_NO_L2PERSIST") ? 1 : 0;
    k.tail_kind = (int)env_i("TOPO_TAIL_KIND", 2);
    k.extrap = getenv("TOPO_EXTRAP") ? atof(getenv("TOPO_EXTRAP")) : 0.0;
    k.lvl_tiled = (int)env_i("TOPO_LVL_TILED", 1);
    k.setup_fast = (int)env_i("TOPO_SETUP_FAST", 1);
  }
  h->omega = getenv("TOPO_OMEGA") ? atof(getenv("TOPO_OMEGA")) : 0.7;
  h->nu_pre = 1;
  h->nu_post = 1;
  h->nu_coarse = getenv("TOPO_NU_COARSE") ? atoi(getenv("TOPO_NU_COARSE")) : 0;
  h->nu_tail_big = getenv("TOPO_NU_TAIL_BIG") ? atoi(getenv("TOPO_NU_TAIL_BIG")) : 2;
  h->nu_tail_small = getenv("TOPO_NU_TAIL_SMALL") ? atoi(getenv("TOPO_NU_TAIL_SMALL")) : 12;
  h->tail_max_nodes = getenv("TOPO_TAIL_MAX") ? atoi(getenv("TOPO_TAIL_MAX")) : 10000;
  h->tail_dbg = nullptr;
  h->u_prev = nullptr;
  h->u_prev_valid = false;
  h->tail2_cache = nullptr;
  h->tail2_cache_tail = -1;
  h->use_graph = getenv("TOPO_NO_GRAPH") ? false : true;
  h->pcg_graph_valid = false;
  h->pcg_graph_exec = nullptr;
  h->pcg_graph_launches = 0;
  h->n_loads = 0;
  h->oc_lam_prev = 0.0;
  h->last_pcg_iters = 0;
  h->loads_set = h->cons_set = h->diag_valid = h->gmg_valid = false;
  h->d_load_node = nullptr;
  h->d_load_fx = h->d_load_fy = nullptr;
  std::memset(&h->timing, 0, sizeof(h->timing));
  CUDA_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  h->owns_stream = true;
  CUDA_TRY(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&h->ev_inv, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&h->ev_copy, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&h->ev_edge[0], cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&h->ev_edge[1], cudaEventDisableTiming));
  h->edge_pending = false;
  h->rho_out_host = nullptr;
  h->inv_pending = false;
  h->inv_overlap = getenv("TOPO_NO_INV_OVERLAP") ? false : true;
  CUDA_TRY(cudaEventCreate(&h->ev0));
  CUDA_TRY(cudaEventCreate(&h->ev1));
  for (int i = 0; i < 8; ++i) CUDA_TRY(cudaEventCreate(&h->evs[i]));
  CUDA_TRY(cudaEventCreate(&h->ev_t0));
  CUDA_TRY(cudaEventCreate(&h->ev_t1));
  h->prof_on = false;
  h->prof_n = 0;

  // nodal vectors in one slab (z and p adjacent: ESO_stress borrows them as 3 planes of nn doubles)
  const size_t nvec = 12;
  double2 *slab = nullptr;
  CUDA_TRY(cudaMalloc(&slab, sizeof(double2) * (size_t)h->nn * nvec));
  CUDA_TRY(cudaMemsetAsync(slab, 0, sizeof(double2) * (size_t)h->nn * nvec, h->stream));
  h->slab_nodal = slab;
  double2 **vecs[nvec] = {&h->u, &h->f, &h->r, &h->z, &h->p, &h->Ap, &h->dinv, &h->tmp, &h->tmp2, &h->u2, &h->r2, &h->p2};
  for (size_t i = 0; i < nvec; ++i) *vecs[i] = slab + i * (size_t)h->nn;
  double *eslab = nullptr;
  const size_t nev = 7;
  const size_t ne_pad = ((size_t)h->ne + 1) & ~(size_t)1;        // every element field 16-byte aligned (double2 loads)
  CUDA_TRY(cudaMalloc(&eslab, sizeof(double) * ne_pad * nev));
  CUDA_TRY(cudaMemsetAsync(eslab, 0, sizeof(double) * ne_pad * nev, h->stream));
  h->slab_elem = eslab;
  double **evs[nev] = {&h->rho, &h->rho_new, &h->dc, &h->dc2, &h->scale, &h->sens, &h->sens_prev};
  for (size_t i = 0; i < nev; ++i) *evs[i] = eslab + i * ne_pad;
  CUDA_TRY(cudaMalloc(&h->nodal, sizeof(double) * (size_t)h->nn));
  CUDA_TRY(cudaMalloc(&h->dinv_f, sizeof(float2) * (size_t)h->nn));
  CUDA_TRY(cudaMalloc(&h->fixed, (size_t)h->nn + 128));         // + slack: the stage-1 kernel copies up to 68 bytes of aligned words from the last node on
  CUDA_TRY(cudaMemsetAsync(h->fixed, 0, (size_t)h->nn + 128, h->stream));
  CUDA_TRY(cudaMalloc(&h->keep, (size_t)h->ne));
  CUDA_TRY(cudaMemsetAsync(h->keep, 1, (size_t)h->ne, h->stream));
  CUDA_TRY(cudaMalloc(&h->partials, sizeof(double) * TOPO_RED_SLOTS * TOPO_MAX_PARTIAL_BLOCKS));
  CUDA_TRY(cudaMalloc(&h->ticket, 4 * sizeof(unsigned int)));          // [1]: second reduction of a kernel that folds twice
  CUDA_TRY(cudaMemsetAsync(h->ticket, 0, 4 * sizeof(unsigned int), h->stream));
  CUDA_TRY(cudaMalloc(&h->red_out, sizeof(double) * 16));
  CUDA_TRY(cudaMemsetAsync(h->red_out, 0, sizeof(double) * 16, h->stream));
  CUDA_TRY(cudaMalloc(&h->sc, sizeof(SolverScalars)));
  CUDA_TRY(cudaMemsetAsync(h->sc, 0, sizeof(SolverScalars), h->stream));
  CUDA_TRY(cudaMallocHost(&h->sc_host, sizeof(SolverScalars)));
  CUDA_TRY(cudaMalloc(&h->oc, sizeof(OcState)));
  CUDA_TRY(cudaMallocHost(&h->oc_host, sizeof(OcState)));
  CUDA_TRY(cudaMalloc(&h->kmax, sizeof(double)));
  CUDA_TRY(cudaMalloc(&h->sel, sizeof(unsigned long long) * 260));
  k_fill<<<nblk(h->ne), 256, 0, h->stream>>>(h->scale, h->ne, 1.0);
  CUDA_TRY(cudaGetLastError());

  h->fine_nc = getenv("TOPO_FINE_NC") ? atoi(getenv("TOPO_FINE_NC")) : (nx + 1 >= 1024 ? 2 : 1);
  if (h->fine_nc != 2) h->fine_nc = 1;
  int s = topo_gmg_plan(h);
  if (s != TOPO_OK) return s;
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  *out = h;
  return TOPO_OK;
}

void topo_dist_free(topo_handle *h);

extern "C" int topo_destroy(topo_t *h) {
  if (!h) return TOPO_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  cudaStreamSynchronize(h->side);
  cudaFree(h->slab_nodal);
  cudaFree(h->slab_elem);
  for (int i = 0; i < 8; ++i) cudaEventDestroy(h->evs[i]);
  cudaEventDestroy(h->ev_t0);
  cudaEventDestroy(h->ev_t1);
  for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
  cudaFree(h->nodal);
  cudaFree(h->dinv_f);
  cudaFree(h->u_prev);
  cudaFree(h->fixed);
  cudaFree(h->keep);
  cudaFree(h->partials);
  cudaFree(h->ticket);
  cudaFree(h->red_out);
  cudaFree(h->sc);
  cudaFreeHost(h->sc_host);
  cudaFree(h->oc);
  cudaFreeHost(h->oc_host);
  cudaFree(h->kmax);
  cudaFree(h->sel);
  cudaFree(h->d_load_node);
  cudaFree(h->d_load_fx);
  cudaFree(h->d_load_fy);
  topo_gmg_free(h);
  topo_dist_free(h);
  if (h->pcg_graph_exec) cudaGraphExecDestroy(h->pcg_graph_exec);
  cudaEventDestroy(h->ev0);
  cudaEventDestroy(h->ev1);
  if (h->owns_stream) cudaStreamDestroy(h->stream);
  cudaStreamDestroy(h->side);
  cudaEventDestroy(h->ev_fork);
  cudaEventDestroy(h->ev_inv);
  cudaEventDestroy(h->ev_copy);
  cudaEventDestroy(h->ev_edge[0]);
  cudaEventDestroy(h->ev_edge[1]);
  cudaGetLastError();
  delete h;
  return TOPO_OK;
}

extern "C" int topo_set_cons(topo_t *h, const int8_t *cons) {
  if (!h || !cons) return TOPO_EINVAL;
  CUDA_TRY(cudaSetDevice(h->device));
  int8_t *d = reinterpret_cast<int8_t *>(h->tmp);
  CUDA_TRY(cudaMemcpyAsync(d, cons, 2 * (size_t)h->nn, cudaMemcpyHostToDevice, h->stream));
  k_cons_to_fixed<<<nblk(h->nn), 256, 0, h->stream>>>(d, h->fixed, h->nn);
  CUDA_TRY(cudaGetLastError());
  h->cons_set = true;
  h->diag_valid = h->gmg_valid = false;
  if (h->loads_set) TOPO_TRY(topo_apply_loads(h));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return TOPO_OK;
}

extern "C" int topo_set_loads(topo_t *h, const int64_t *node, const double *fx, const double *fy, int n_loads) {
  if (!h || n_loads < 0 || (n_loads > 0 && (!node || !fx || !fy))) return TOPO_EINVAL;
  for (int i = 0; i < n_loads; ++i)
    if (node[i] < 0 || node[i] >= h->nn) {
      topo_set_error("load %d: node id %lld out of range [0,%lld)", i, (long long)node[i], (long long)h->nn);
      return TOPO_EINVAL;
    }
  CUDA_TRY(cudaSetDevice(h->device));
  cudaFree(h->d_load_node); cudaFree(h->d_load_fx); cudaFree(h->d_load_fy);
  h->d_load_node = nullptr; h->d_load_fx = h->d_load_fy = nullptr;
  h->n_loads = n_loads;
  if (n_loads > 0) {
    CUDA_TRY(cudaMalloc(&h->d_load_node, sizeof(int64_t) * n_loads));
    CUDA_TRY(cudaMalloc(&h->d_load_fx, sizeof(double) * n_loads));
    CUDA_TRY(cudaMalloc(&h->d_load_fy, sizeof(double) * n_loads));
    CUDA_TRY(cudaMemcpyAsync(h->d_load_node, node, sizeof(int64_t) * n_loads, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->d_load_fx, fx, sizeof(double) * n_loads, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->d_load_fy, fy, sizeof(double) * n_loads, cudaMemcpyHostToDevice, h->stream));
  }
  h->loads_set = true;
  TOPO_TRY(topo_apply_loads(h));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return TOPO_OK;
}

extern "C" int topo_set(topo_t *h, int field, const void *host, size_t bytes) {
  if (!h || !host) return TOPO_EINVAL;
  if (field == TOPO_F_CONS) {
    if (bytes != 2 * (size_t)h->nn) { topo_set_error("topo_set(CONS): expected %zu bytes", 2 * (size_t)h->nn); return TOPO_EINVAL; }
    return topo_set_cons(h, static_cast<const int8_t *>(host));
  }
  FieldRef f;
  TOPO_TRY(field_ref(h, field, &f));
  if (bytes != f.bytes) {
    topo_set_error("topo_set(field %d): got %zu bytes, expected %zu", field, bytes, f.bytes);
    return TOPO_EINVAL;
  }
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaMemcpyAsync(f.ptr, host, bytes, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (field == TOPO_F_SCALE || field == TOPO_F_RHO) h->diag_valid = h->gmg_valid = false;
  if (field == TOPO_F_LOAD) h->loads_set = true;
  return TOPO_OK;
}

extern "C" int topo_get(topo_t *h, int field, void *host, size_t bytes) {
  if (!h || !host) return TOPO_EINVAL;
  FieldRef f;
  TOPO_TRY(field_ref(h, field, &f));
  if (bytes != f.bytes) {
    topo_set_error("topo_get(field %d): got %zu bytes, expected %zu", field, bytes, f.bytes);
    return TOPO_EINVAL;
  }
  CUDA_TRY(cudaSetDevice(h->device));
  if (field == TOPO_F_CONS) {
    int8_t *d = reinterpret_cast<int8_t *>(h->tmp);
    k_fixed_to_cons<<<nblk(h->nn), 256, 0, h->stream>>>(h->fixed, d, h->nn);
    CUDA_TRY(cudaGetLastError());
    f.ptr = d;
  }
  CUDA_TRY(cudaMemcpyAsync(host, f.ptr, bytes, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return TOPO_OK;
}

extern "C" int topo_fill(topo_t *h, int field, double value) {
  if (!h) return TOPO_EINVAL;
  CUDA_TRY(cudaSetDevice(h->device));
  if (field == TOPO_F_KEEP) {
    k_fill8<<<nblk(h->ne), 256, 0, h->stream>>>(h->keep, h->ne, value != 0.0 ? 1 : 0);
  } else if (field == TOPO_F_RHO || field == TOPO_F_SCALE || field == TOPO_F_SENS || field == TOPO_F_SENS_PREV ||
             field == TOPO_F_DC) {
    FieldRef f;
    TOPO_TRY(field_ref(h, field, &f));
    k_fill<<<nblk(h->ne), 256, 0, h->stream>>>(static_cast<double *>(f.ptr), h->ne, value);
    h->diag_valid = h->gmg_valid = false;
  } else {
    topo_set_error("topo_fill: field %d is not an element field", field);
    return TOPO_EINVAL;
  }
  CUDA_TRY(cudaGetLastError());
  return TOPO_OK;
}

extern "C" int topo_matvec(topo_t *h, const double *u_host, double *y_host) {
  if (!h || !u_host || !y_host) return TOPO_EINVAL;
  CUDA_TRY(cudaSetDevice(h->device));
  const size_t nv = sizeof(double2) * (size_t)h->nn;
  CUDA_TRY(cudaMemcpyAsync(h->p, u_host, nv, cudaMemcpyHostToDevice, h->stream));
  TOPO_TRY(topo_mask_vector(h, h->p));
  TOPO_TRY(topo_fine_matvec(h, h->p, h->Ap));
  CUDA_TRY(cudaMemcpyAsync(y_host, h->Ap, nv, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return TOPO_OK;
}

extern "C" int topo_matvec_device(topo_t *h, int repeats, float *ms_per_launch) {
  if (!h || repeats < 1) return TOPO_EINVAL;
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
  for (int i = 0; i < repeats; ++i) TOPO_TRY(topo_fine_matvec(h, h->p, h->Ap));
  CUDA_TRY(cudaEventRecord(h->ev1, h->stream));
  CUDA_TRY(cudaEventSynchronize(h->ev1));
  float ms = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->timing.matvec_ms = ms / repeats;
  if (ms_per_launch) *ms_per_launch = ms / repeats;
  return TOPO_OK;
}

extern "C" int topo_simp_step(topo_t *h, const topo_simp_params *p, double *g_inout, topo_simp_result *res) {
  if (!h || !p || !g_inout || !res) return TOPO_EINVAL;
  std::memset(res, 0, sizeof(*res));
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  CUDA_TRY(cudaEventRecord(h->evs[0], st));
  TOPO_TRY(topo_simp_scale(h, p->penal, p->emin, p->emax));
  CUDA_TRY(cudaEventRecord(h->evs[1], st));
  // the multigrid attempt is capped; a stalled solve restarts from its last iterate with the Jacobi preconditioner
  // (converges on any SPD system) -- a safety net: SIMP designs keep Emin > 0 and have not needed it in any run
  if (h->knobs.extrap != 0.0 && p->warm) {
    if (!h->u_prev) CUDA_TRY(cudaMalloc(&h->u_prev, sizeof(double2) * (size_t)h->nn));
    KL(h, "k_extrapolate", k_extrapolate<<<nblk(h->nn), 256, 0, st>>>(h->u, h->u_prev, h->nn, h->knobs.extrap, h->u_prev_valid ? 1 : 0));
    h->u_prev_valid = true;
  }
  const int gmg_cap = 4000;
  const bool gmg = p->precond == TOPO_PRECOND_GMG;
  int s = topo_solve(h, p->precond, p->rtol, gmg ? std::min(p->maxit, gmg_cap) : p->maxit, p->warm, &res->cg_iters, &res->relres);
  if (gmg && s == TOPO_ENOCONV && p->maxit > gmg_cap) {
    int it2 = 0;
    const int warm2 = (res->relres == res->relres && res->relres < 1.0) ? 1 : 0;
    s = topo_solve(h, TOPO_PRECOND_JACOBI, p->rtol, p->maxit, warm2, &it2, &res->relres);
    res->cg_iters += it2;
  }
  res->status = s;
  if (s != TOPO_OK && s != TOPO_ENOCONV) return s;
  CUDA_TRY(cudaEventRecord(h->evs[2], st));
  TOPO_TRY(topo_compliance(h, &res->compliance));
  TOPO_TRY(topo_simp_sensitivity(h, p->penal, p->emin, p->emax, &res->objective));
  CUDA_TRY(cudaEventRecord(h->evs[3], st));
  TOPO_TRY(topo_filter_sensitivity(h, p->rmin, p->dx, p->dy));
  CUDA_TRY(cudaEventRecord(h->evs[4], st));
  double g = *g_inout;
  TOPO_TRY(topo_oc_update(h, p->move, &g, &res->change, &res->oc_steps));
  *g_inout = g;
  res->g = g;
  CUDA_TRY(cudaEventRecord(h->evs[5], st));
  if (h->rho_out_host != nullptr) {              // (topo_simp_step_host) download of the new densities || equilibrium guard
    CUDA_TRY(cudaEventRecord(h->ev_fork, st));
    CUDA_TRY(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
    CUDA_TRY(cudaMemcpyAsync(h->rho_out_host, h->rho, sizeof(double) * (size_t)h->ne, cudaMemcpyDeviceToHost, h->side));
    CUDA_TRY(cudaEventRecord(h->ev_copy, h->side));
    h->rho_out_host = nullptr;                  // (topo_simp_step_host waits for the side stream)
  }
  TOPO_TRY(topo_equilibrium_check(h, &res->equilibrium_ok));
  CUDA_TRY(cudaEventRecord(h->evs[6], st));
  CUDA_TRY(cudaEventSynchronize(h->evs[6]));
  CUDA_TRY(cudaEventElapsedTime(&h->timing.setup_ms, h->evs[0], h->evs[1]));
  CUDA_TRY(cudaEventElapsedTime(&h->timing.solve_ms, h->evs[1], h->evs[2]));
  CUDA_TRY(cudaEventElapsedTime(&h->timing.sens_ms, h->evs[2], h->evs[3]));
  CUDA_TRY(cudaEventElapsedTime(&h->timing.filter_ms, h->evs[3], h->evs[4]));
  CUDA_TRY(cudaEventElapsedTime(&h->timing.oc_ms, h->evs[4], h->evs[5]));
  CUDA_TRY(cudaEventElapsedTime(&h->timing.total_ms, h->evs[0], h->evs[6]));
  return s;
}

extern "C" int topo_simp_step_host(topo_t *h, const topo_simp_params *p, const double *rho_in, double *rho_out,
                                   double *g_inout, topo_simp_result *res) {
  if (!h || !rho_in || !rho_out) return TOPO_EINVAL;
  const size_t ev = sizeof(double) * (size_t)h->ne;
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaMemcpyAsync(h->rho, rho_in, ev, cudaMemcpyHostToDevice, h->stream));
  h->rho_out_host = rho_out;                     // downloaded by the step itself, overlapping its equilibrium guard
  int s = topo_simp_step(h, p, g_inout, res);
  const bool copied = h->rho_out_host == nullptr;
  h->rho_out_host = nullptr;
  if (s != TOPO_OK && s != TOPO_ENOCONV) return s;
  if (!copied) CUDA_TRY(cudaMemcpyAsync(rho_out, h->rho, ev, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->side));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return s;
}

extern "C" int topo_get_timing(topo_t *h, topo_timing *t) {
  if (!h || !t) return TOPO_EINVAL;
  *t = h->timing;
  return TOPO_OK;
}

extern "C" int topo_timer_start(topo_t *h) {
  if (!h) return TOPO_EINVAL;
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaEventRecord(h->ev_t0, h->stream));
  return TOPO_OK;
}

extern "C" int topo_timer_stop(topo_t *h, float *ms) {
  if (!h || !ms) return TOPO_EINVAL;
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaEventRecord(h->ev_t1, h->stream));
  CUDA_TRY(cudaEventSynchronize(h->ev_t1));
  CUDA_TRY(cudaEventElapsedTime(ms, h->ev_t0, h->ev_t1));
  return TOPO_OK;
}

extern "C" int topo_profile_matvec(topo_t *h, int enable) {
  if (!h) return TOPO_EINVAL;
  CUDA_TRY(cudaSetDevice(h->device));
  if (enable && h->prof_ev.empty()) {
    h->prof_ev.resize(2 * 32768);
    h->prof_name.assign(32768, nullptr);
    for (auto &e : h->prof_ev) CUDA_TRY(cudaEventCreate(&e));
  }
  h->prof_on = enable != 0;
  if (enable) h->prof_n = 0;
  if (!enable && !h->prof_ev.empty()) {
    // read the report BEFORE switching off: tens of thousands of recorded events slow down every later
    // synchronisation of the process (measured: 3.5 -> 87 ms per BESO iteration on another handle)
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
    h->prof_ev.clear();
    h->prof_name.clear();
    h->prof_n = 0;
  }
  return TOPO_OK;
}

extern "C" int topo_profile_read(topo_t *h, int *launches, float *total_ms) {
  // stage-1 kernel only: every variant of the fine-grid K application
  if (!h || !launches || !total_ms) return TOPO_EINVAL;
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  float tot = 0.f;
  int cnt = 0;
  for (int i = 0; i < h->prof_n; ++i) {
    if (!h->prof_name[i] || strncmp(h->prof_name[i], "k_fine", 6) != 0) continue;
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, h->prof_ev[2 * i], h->prof_ev[2 * i + 1]));
    tot += ms;
    ++cnt;
  }
  *launches = cnt;
  *total_ms = tot;
  return TOPO_OK;
}

extern "C" int topo_profile_report(topo_t *h, char *buf, size_t buflen) {
  // text table "name count total_ms" per kernel, sorted by total time
  if (!h || !buf || buflen == 0) return TOPO_EINVAL;
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  struct Row { const char *name; int n; double ms; };
  std::vector<Row> rows;
  for (int i = 0; i < h->prof_n; ++i) {
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, h->prof_ev[2 * i], h->prof_ev[2 * i + 1]));
    const char *nm = h->prof_name[i] ? h->prof_name[i] : "?";
    bool found = false;
    for (auto &r : rows)
      if (strcmp(r.name, nm) == 0) { r.n++; r.ms += ms; found = true; break; }
    if (!found) rows.push_back({nm, 1, (double)ms});
  }
  std::sort(rows.begin(), rows.end(), [](const Row &a, const Row &b) { return a.ms > b.ms; });
  size_t off = 0;
  buf[0] = 0;
  for (auto &r : rows) {
    int w = snprintf(buf + off, buflen - off, "%s %d %.6f\n", r.name, r.n, r.ms);
    if (w < 0 || (size_t)w >= buflen - off) break;
    off += (size_t)w;
  }
  return TOPO_OK;
}

// development aid: stage-boundary time stamps (ns) of the most recent k_tail launch
extern "C" int topo_debug_tail_stamps(topo_t *h, unsigned long long *out, int max_out, int *n) {
  if (!h || !out || !n) return TOPO_EINVAL;
  CUDA_TRY(cudaSetDevice(h->device));
  if (!h->tail_dbg) {
    CUDA_TRY(cudaMalloc(&h->tail_dbg, sizeof(unsigned long long) * 512));
    CUDA_TRY(cudaMemset(h->tail_dbg, 0, sizeof(unsigned long long) * 512));
    h->pcg_graph_valid = false;
    *n = 0;
    return TOPO_OK;
  }
  unsigned long long buf[512];
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  CUDA_TRY(cudaMemcpy(buf, h->tail_dbg, sizeof(buf), cudaMemcpyDeviceToHost));
  int cnt = (int)std::min<unsigned long long>(buf[0], (unsigned long long)std::min(max_out, 511));
  for (int i = 0; i < cnt; ++i) out[i] = buf[1 + i];
  *n = cnt;
  return TOPO_OK;
}

// ---- multi-GPU building blocks: operate on caller-owned device pointers ------------------------------------
namespace {
__global__ void k_axpby(const double2 *__restrict__ a, const double2 *__restrict__ b, double2 *__restrict__ c, int64_t n,
                        double alpha, double beta) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double2 x = a[i], y = b[i];
    c[i] = make_double2(alpha * x.x + beta * y.x, alpha * x.y + beta * y.y);
  }
}
__global__ void k_vmul(const double2 *__restrict__ a, const double2 *__restrict__ b, double2 *__restrict__ c, int64_t n,
                       int divide) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double2 x = a[i], y = b[i];
    c[i] = divide ? make_double2(y.x != 0.0 ? x.x / y.x : 0.0, y.y != 0.0 ? x.y / y.y : 0.0)
                  : make_double2(x.x * y.x, x.y * y.y);
  }
}
__global__ void k_dot_rows(const double2 *__restrict__ a, const double2 *__restrict__ b, int64_t n0, int64_t n1,
                           double *partials, unsigned int *ticket, double *out) {
  double v[1] = {0.0};
  for (int64_t i = n0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n1; i += (int64_t)gridDim.x * blockDim.x) {
    const double2 x = a[i], y = b[i];
    v[0] = fma(x.x, y.x, fma(x.y, y.y, v[0]));
  }
  grid_reduce<1>(v, partials, ticket, out);
}
__global__ void k_diag_from_dinv(const double2 *__restrict__ dinv, double2 *__restrict__ d, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double2 v = dinv[i];
    d[i] = make_double2(v.x != 0.0 ? 1.0 / v.x : 0.0, v.y != 0.0 ? 1.0 / v.y : 0.0);
  }
}
}  // namespace

extern "C" int topo_set_stream(topo_t *h, void *cuda_stream) {
  if (!h) return TOPO_EINVAL;
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (h->owns_stream) cudaStreamDestroy(h->stream);
  h->stream = static_cast<cudaStream_t>(cuda_stream);
  h->owns_stream = false;
  h->pcg_graph_valid = false;
  return TOPO_OK;
}

extern "C" int topo_matvec_dev(topo_t *h, const double *u_dev, double *y_dev) {
  if (!h || !u_dev || !y_dev) return TOPO_EINVAL;
  CUDA_TRY(cudaSetDevice(h->device));
  return topo_fine_matvec(h, reinterpret_cast<const double2 *>(u_dev), reinterpret_cast<double2 *>(y_dev));
}

extern "C" int topo_diag_dev(topo_t *h, double *diag_dev) {
  if (!h || !diag_dev) return TOPO_EINVAL;
  CUDA_TRY(cudaSetDevice(h->device));
  TOPO_TRY(topo_build_diag(h));
  KL(h, "k_diag_from_dinv", k_diag_from_dinv<<<nblk(h->nn), 256, 0, h->stream>>>(h->dinv, reinterpret_cast<double2 *>(diag_dev), h->nn));
  CUDA_TRY(cudaGetLastError());
  return TOPO_OK;
}

extern "C" int topo_vec_axpby_dev(topo_t *h, long long n_nodes, double alpha, const double *a_dev, double beta,
                                  const double *b_dev, double *c_dev) {
  if (!h || !a_dev || !b_dev || !c_dev || n_nodes < 0) return TOPO_EINVAL;
  CUDA_TRY(cudaSetDevice(h->device));
  KL(h, "k_axpby", k_axpby<<<nblk(n_nodes), 256, 0, h->stream>>>(reinterpret_cast<const double2 *>(a_dev),
                                                                 reinterpret_cast<const double2 *>(b_dev),
                                                                 reinterpret_cast<double2 *>(c_dev), n_nodes, alpha, beta));
  CUDA_TRY(cudaGetLastError());
  return TOPO_OK;
}

extern "C" int topo_vec_mul_dev(topo_t *h, long long n_nodes, const double *a_dev, const double *b_dev, double *c_dev,
                                int divide) {
  if (!h || !a_dev || !b_dev || !c_dev || n_nodes < 0) return TOPO_EINVAL;
  CUDA_TRY(cudaSetDevice(h->device));
  KL(h, "k_vmul", k_vmul<<<nblk(n_nodes), 256, 0, h->stream>>>(reinterpret_cast<const double2 *>(a_dev),
                                                               reinterpret_cast<const double2 *>(b_dev),
                                                               reinterpret_cast<double2 *>(c_dev), n_nodes, divide));
  CUDA_TRY(cudaGetLastError());
  return TOPO_OK;
}

extern "C" int topo_vec_dot_dev(topo_t *h, const double *a_dev, const double *b_dev, int row_lo, int row_hi,
                                double *out_dev) {
  if (!h || !a_dev || !b_dev || !out_dev || row_lo < 0 || row_hi > h->ny + 1 || row_lo > row_hi) return TOPO_EINVAL;
  CUDA_TRY(cudaSetDevice(h->device));
  const int64_t npr = h->nx + 1, n0 = (int64_t)row_lo * npr, n1 = (int64_t)row_hi * npr;
  KL(h, "k_dot_rows", k_dot_rows<<<nblk(std::max<int64_t>(n1 - n0, 1)), 256, 0, h->stream>>>(
         reinterpret_cast<const double2 *>(a_dev), reinterpret_cast<const double2 *>(b_dev), n0, n1, h->partials, h->ticket,
         out_dev));
  CUDA_TRY(cudaGetLastError());
  return TOPO_OK;
}
