// Internal declarations shared by the translation units of libtopo_b200.so (sm_100a only).
// Public ABI: include/topo_b200.h.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/topo_b200.h"

#define TOPO_MAX_LEVELS 32
#define TOPO_RED_SLOTS 16         // scalars reduced per kernel (max)
#define TOPO_MAX_PARTIAL_BLOCKS 8192

void topo_set_error(const char *fmt, ...);

#define CUDA_TRY(expr)                                                                     \
  do {                                                                                     \
    cudaError_t e__ = (expr);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      topo_set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,                 \
                     cudaGetErrorString(e__));                                             \
      return TOPO_ECUDA;                                                                   \
    }                                                                                      \
  } while (0)

#define TOPO_TRY(expr)                                                                     \
  do {                                                                                     \
    int s__ = (expr);                                                                      \
    if (s__ != TOPO_OK) return s__;                                                        \
  } while (0)

// 8x8 element matrix passed by value as a kernel parameter: lives in the constant bank, so DFMA can
// take its entries as c[][] operands without occupying registers.
struct KeParam {
  double k[64];
};

// Device-resident scalars shared between the solver kernels (no host round trip inside the PCG loop).
struct SolverScalars {
  double rz[2];      // r.z of the current / next iteration (indexed by parity)
  double pAp;
  double rr;         // ||r||^2
  double bb;         // ||f||^2
  double tol2;       // rtol^2
  double aux[4];
  int iters;         // PCG iterations completed
  int done;          // 1 once rr <= tol2*bb (kernels become no-ops)
  int maxit;
  int tail_error;    // a bounded hand-over wait inside the tail kernel gave up (results invalid)
};

struct OcState {
  double l1, l2, lmid, gt, g, change;
  int steps, done;
  int need_exact;   // the next bisection decision must come from the exact (reference-formula) pass
  int pad;
  // gt(lambda) is non-increasing in lambda: gt(ka) > margin and gt(kb) < -margin were MEASURED (probe pass around
  // the previous design iteration's multiplier), so every bisection midpoint <= ka or >= kb has a known decision
  double ka, kb;
};

// Coarse-level stencil coefficients are stored in fp32: they only shape the PRECONDITIONER (the outer PCG
// residual is formed with the exact fp64 fine operator), A(n,m) and A(m,n)^T round identically so symmetry
// is kept, and the stencil stream is the dominant traffic of levels >= 1.  Vectors stay fp64.
typedef float stencil_t;

// one coarse level of the multigrid hierarchy (level 0 = the matrix-free fine grid, not stored here)
struct GmgLevel {
  int nx, ny;          // cells
  int nn;              // nodes
  int ncell;
  double *cell;        // 64 planes x ncell : per-cell 8x8 Galerkin matrix (SoA, plane = 8*i+j)
  stencil_t *st;        // 9 planes x nn x float4: node stencil, plane nb = 3*(dr+1)+(dc+1), entry (a00,a01,a10,a11)
  double2 *dinv;       // 1/diag (0 at constrained DOFs)
  double2 *x, *x2, *b, *r;
  uint8_t *fixed;
  // level 1: one byte per NODE, level 2: one byte per CELL -- 1 where the Galerkin set-up may take its branch-free path
  // (all cells / children around exist in full and no constrained DOF is near); rewritten by every set-up because
  // ESO / BESO pin nodes as they go
  uint8_t *interior;
};

// Row-slab decomposition: this handle is slab `rank` of `world` equal slabs of a mesh with ny_global
// element rows.  Vectors are stored in ACCUMULATED form (interface node rows hold the full value on both
// neighbours); a local operator application yields a DISTRIBUTED result (partial sums on interface rows) that
// the host completes with a halo sum.  Where a kernel needs a distributed copy of an accumulated vector it
// scales the interface rows by w_first / w_last (1/2 each side = a partition of unity).
// Peer-memory transport of the row-slab mode (NVLink/NVSwitch): every rank owns one cudaMalloc'd REGION that its
// peers map (CUDA IPC across processes, plain pointers inside one process) and WRITE into -- interface rows,
// level-K right-hand-side pieces, scalar partial sums -- followed by a monotonically increasing sequence number
// (st.release.sys).  The consumer kernel spins on its own memory (ld.acquire.sys) until the number arrives.
// Payload buffers are double-buffered by the parity of the sequence number: a rank can be at most one exchange
// ahead of its neighbour.  Counters live in local, unshared memory.
#define TOPO_P2P_MAXW 16
struct P2PHeader {                       // start of the region; written by peers
  unsigned long long flag_from_lo, flag_from_hi;           // halo rows: sequence number of the neighbour's last push
  unsigned long long flag_gather[TOPO_P2P_MAXW];           // by rank
  unsigned long long flag_scalar[TOPO_P2P_MAXW];
  double scalar_slot[2][TOPO_P2P_MAXW][16];
  uint4 ll_scalar[2][TOPO_P2P_MAXW][16];                    // one-launch all-reduce: every double travels as (lo, seq, hi, seq)
};
struct P2PCounters {                     // local
  unsigned long long halo_push, halo_wait, gather_push, gather_wait, scal_push, scal_wait;
  unsigned int ticket[4];
  unsigned int error;                    // a wait gave up (peer never wrote): results are invalid
  unsigned int pad;
  unsigned int fticket[4];               // one-launch exchanges: push tickets towards rank-1 / rank+1, end-of-kernel ticket, gather
};

struct DistState {
  bool on;
  int rank, world, K;                  // K = first level that is global (replicated on every rank)
  bool has_lo, has_hi;
  int ny_global;
  double w_first, w_last;
  double2 *recv_lo, *recv_hi;          // one node row of the fine grid each (largest exchange)
  double2 *diag_send_lo, *diag_send_hi, *diag_recv_lo, *diag_recv_hi;   // interface diagonals of levels 0..K-1
  long long diag_count;                // double2 per side
  double *cellK_send, *cellK_recv;     // level-K cell matrices: [64][ncellK_loc] -> [world][64][ncellK_loc]
  long long ncellK_loc;
  double2 *bK_send, *bK_recv;          // level-K right-hand side: (rowsK_loc+1) x nprK per rank
  long long bK_count;                  // double2 per rank
  double *filt_send_lo, *filt_send_hi, *filt_recv_lo, *filt_recv_hi;    // rho*dc of hy element rows
  int filt_hy;
  // peer-memory transport
  bool p2p_on;
  bool p2p_inkernel;                   // the restriction / up kernels of slab-local levels swap their interface rows themselves
  int mid_from;                        // > 0: slab-local levels mid_from..K-1 run inside k_slab_mid_down / k_slab_mid_up
  int mid_grid;                        // blocks of those persistent kernels (all resident; >= tiles of level mid_from)
  unsigned int *mid_bar;               // their grid barrier: [0] arrivals, [1] generation
  char *p2p_region;                    // P2PHeader | recv_lo[2][npr] | recv_hi[2][npr] | gather_recv[2][world][bK_count]
  size_t p2p_bytes, p2p_off_lo, p2p_off_hi, p2p_off_gather, p2p_row_stride, p2p_gather_stride;
  size_t p2p_off_ll_lo, p2p_off_ll_hi, p2p_off_ll_gather;   // flag-in-data copies of the three payload areas (32 B per double2)
  char *p2p_peer[TOPO_P2P_MAXW];       // mapped regions by rank (own region at [rank])
  bool p2p_opened[TOPO_P2P_MAXW];      // mapped through cudaIpcOpenMemHandle (must be closed)
  P2PCounters *p2p_cnt;
};

// Experiment knobs (environment variables): read ONCE per handle in topo_create, never in a
// launch path.  Defaults are what every number in DESIGN.md / profiles/ was measured with.
struct TopoKnobs {
  double fine_waves;        // TOPO_FINE_WAVES      waves of CTAs the fine-grid row chunks are cut for (1)
  int fine_nc_mask;         // TOPO_FINE_NC_MASK    bit k: two node columns per lane for fine-grid kind k (0)
  long long sym_min;        // TOPO_SYM_MIN         levels with at least this many nodes read the symmetric stencil half (100000)
  int nu_coarse_maxlvl;     // TOPO_NU_COARSE_MAXLVL
  int tail_cluster_max;     // TOPO_TAIL_CLUSTER    largest cluster tried for the tail kernel (16)
  int unfused_min;          // TOPO_UNFUSED_MIN     levels above this many nodes: separate prolongation + smoothing (200000)
  int fuse_prolong;         // TOPO_FUSE_PROLONG
  int oc_ctas_per_sm;       // TOPO_OC_CTAS_PER_SM  (2)
  int oc_no_probe;          // TOPO_OC_NO_PROBE
  int coarsest_nodes;       // TOPO_COARSEST_NODES  first level solved densely (48)
  int no_l2persist;         // TOPO_NO_L2PERSIST
  double extrap;            // TOPO_EXTRAP          initial guess of a warm solve: u + extrap*(u - u_previous_design) (0: off)
  int tail_kind;            // TOPO_TAIL_KIND       2: shared-memory tail k_tail2 where its plan fits (default), 0: generic k_tail
  int setup_fast;           // TOPO_SETUP_FAST      1 (default): branch-free level-1 / level-2 Galerkin kernels where k_interior_flags
                            //                      allows; 0: every node / cell takes the general instantiation
  int lvl_tiled;            // TOPO_LVL_TILED       1 (default): tile kernels for levels 1..tail-1 (restriction inside the down
                            //                      kernel, prolongation inside the up kernel); 0: separate transfer kernels
};

struct topo_handle {
  int device;
  TopoKnobs knobs;
  int nx, ny;
  int64_t nn, ne;
  KeParam ke;
  KeParam ke_factor;    // rows l_k of a pivoted Cholesky factor: ke = sum_k l_k l_k^T (zero rows beyond the rank)
  cudaStream_t stream;
  bool owns_stream;
  // side stream for the dense coarsest-level inverse (one latency-bound CTA, ~165 us): it overlaps the rest of the
  // multigrid set-up and the first fine-grid passes of the solve; the first tail launch waits for ev_inv
  cudaStream_t side;
  cudaEvent_t ev_fork, ev_inv;
  bool inv_pending, inv_overlap;
  // topo_simp_step_host: the download of the new densities starts on the side stream as soon as the OC update has
  // finished and overlaps the equilibrium guard (K u = f check: three passes over the fine grid)
  double *rho_out_host;
  cudaEvent_t ev_copy;
  // multigrid set-up: the general ("edge") instantiations of the level-1 / level-2 Galerkin kernels do a strip of nodes
  // along the mesh edge and the supports -- a few warps with a long dependent instruction stream (70 us whatever the
  // mesh) -- and run on the side stream next to the branch-free kernels; joined through these events
  cudaEvent_t ev_edge[2];
  bool edge_pending;

  // fine-grid fields
  double2 *u, *f, *r, *z, *p, *Ap, *dinv, *tmp, *tmp2, *u2, *r2, *p2;
  double2 *u_prev;     // displacements of the previous design iteration (allocated on first use; TOPO_EXTRAP)
  bool u_prev_valid;
  float2 *dinv_f;      // fp32 copy of dinv for the fused multigrid-PCG passes (tmp = res and tmp2 = x^ hold float2 there)
  double *rho, *rho_new, *dc, *dc2, *scale, *sens, *sens_prev, *nodal;
  uint8_t *fixed, *keep;
  int8_t *cons_host_shadow;

  // loads (kept to re-apply when BC flags change)
  std::vector<int64_t> load_node;
  std::vector<double> load_fx, load_fy;
  int64_t *d_load_node;
  double *d_load_fx, *d_load_fy;
  int n_loads;
  bool loads_set, cons_set, diag_valid, gmg_valid;

  // reductions
  double *partials;        // TOPO_RED_SLOTS x TOPO_MAX_PARTIAL_BLOCKS
  unsigned int *ticket;
  double *red_out;         // TOPO_RED_SLOTS results of the most recent reduction
  SolverScalars *sc;       // device
  SolverScalars *sc_host;  // pinned
  OcState *oc;             // device
  OcState *oc_host;        // pinned
  double oc_lam_prev;      // last multiplier of the previous OC update (0: none) -- centre of the probe pass
  double *kmax;            // device scalar: max diagonal of K over free DOFs
  unsigned long long *sel; // radix-select scratch

  // multigrid
  int n_levels;            // including the fine level
  int lvl_nx[TOPO_MAX_LEVELS], lvl_ny[TOPO_MAX_LEVELS];
  std::vector<GmgLevel> levels;   // levels[0] unused placeholder; levels[l] for l>=1
  double *gtab;            // 9 x 64 level-1 Galerkin tables (device)
  double *ptab;            // 9 x 64 child interpolation matrices (device)
  void *lvl1_slab, *coarse_slab;   // allocations behind levels[1] and levels[>=2] (+ coarse_inv)
  size_t coarse_slab_bytes;
  double *coarse_inv;      // dense inverse of the coarsest operator (inside coarse_slab)
  int coarse_n;            // its dimension (2*nn of last level)
  double omega;
  int nu_pre, nu_post;
  int nu_coarse;           // > 0: sweeps on levels >= 1 (overrides nu_pre/nu_post there)
  int nu_tail_big, nu_tail_small;   // sweeps on tail levels with > / <= 1024 nodes
  unsigned long long *tail_dbg;   // device buffer for k_tail stage stamps (nullptr = off)
  int tail_max_nodes;
  void *tail2_cache;       // gmg_tail.cu: the plan of the shared-memory tail kernel (nullptr: not computed yet)
  int tail2_cache_tail;
  bool use_graph, pcg_graph_valid;
  cudaGraphExec_t pcg_graph_exec;
  int last_pcg_iters;          // iterations of the previous multigrid-PCG solve (sizes the first batch of the next)
  int pcg_graph_launches;      // levels with at most this many nodes run inside the single-CTA tail kernel

  void *slab_nodal, *slab_elem;   // allocation bases (field pointers get swapped)

  // timing
  cudaEvent_t ev0, ev1;
  cudaEvent_t evs[8];             // stage boundaries of topo_simp_step
  cudaEvent_t ev_t0, ev_t1;       // topo_timer_start/stop
  bool prof_on;
  int prof_n;
  std::vector<cudaEvent_t> prof_ev;   // pairs
  std::vector<const char *> prof_name;
  topo_timing timing;
  int fine_nc;                    // node columns per lane of the stage-1 kernel (1 or 2)
  DistState dist;
};

// Kernel launch wrapper: counts the launch and, in profile mode, brackets it with CUDA events on the
// handle's stream (per-kernel in-situ durations; read back through topo_profile_report).
static inline void topo_prof_pre(topo_handle *h, const char *name) {
  if (h->prof_on && (size_t)(2 * h->prof_n + 1) < h->prof_ev.size()) {
    h->prof_name[h->prof_n] = name;
    cudaEventRecord(h->prof_ev[2 * h->prof_n], h->stream);
  }
}
static inline void topo_prof_post(topo_handle *h) {
  if (h->prof_on && (size_t)(2 * h->prof_n + 1) < h->prof_ev.size()) {
    cudaEventRecord(h->prof_ev[2 * h->prof_n + 1], h->stream);
    h->prof_n++;
  }
  h->timing.kernel_launches++;
}
#define KL(h, name, ...)        \
  do {                          \
    topo_prof_pre((h), (name)); \
    __VA_ARGS__;                \
    topo_prof_post((h));        \
  } while (0)

static inline int div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
// Function attributes (dynamic shared memory, cluster size) and __constant__ symbols are per DEVICE: one-time set-up
// flags in the launch helpers are indexed by the handle's device so that handles on several GPUs can share a process.
constexpr int TOPO_MAX_DEVICES = 64;
static inline int topo_dev_slot(const topo_handle *h) { return h->device & (TOPO_MAX_DEVICES - 1); }

// ---- kernels / drivers implemented across translation units ----
int topo_fine_matvec(topo_handle *h, const double2 *u, double2 *y);                  // y = A u
int topo_fine_residual(topo_handle *h, const double2 *u, const double2 *b, double2 *r); // r = b - A u
int topo_build_diag(topo_handle *h);
int topo_gmg_plan(topo_handle *h);
int topo_gmg_setup(topo_handle *h);
int topo_apply_loads(topo_handle *h);
void topo_gmg_free(topo_handle *h);

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------
// Kernel launch through cudaLaunchKernelEx (one place for launch attributes).  Programmatic dependent launch was
// tried here for the ~16 short kernels of a PCG iteration (griddepcontrol.wait at every kernel entry, optional early
// release of the dependents): inside the CUDA graph it made the design iteration slower in every combination
// (8.9-10.0 ms against 8.8 ms on 2048x1024) and the entry instructions alone cost the level kernels time, so it was
// removed again.
// ------------------------------------------------------------------------------------------
template <class... KA, class... A>
static inline cudaError_t topo_launch(topo_handle *h, void (*kern)(KA...), dim3 grid, dim3 block, size_t smem,
                                      A &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = h->stream;
  return cudaLaunchKernelEx(&cfg, kern, KA(args)...);
}

// ------------------------------------------------------------------------------------------
// Deterministic two-stage grid reduction.  Every block reduces its values in a fixed order and stores
// one partial per slot; the block that draws the last ticket folds the partials in index order, so
// the result does not depend on block scheduling (needed for reproducible OC bisection decisions,
// reference solver.py:443-447).
// ------------------------------------------------------------------------------------------
template <int NV, bool IS_MAX = false>
__device__ __forceinline__ void block_reduce(double (&v)[NV], double *smem /* NV*32 */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarp = (blockDim.x * blockDim.y + 31) >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double t = __shfl_down_sync(0xffffffffu, v[k], o);
      v[k] = IS_MAX ? fmax(v[k], t) : v[k] + t;
    }
    if (lane == 0) smem[k * 32 + warp] = v[k];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double t = (lane < nwarp) ? smem[k * 32 + lane] : (IS_MAX ? -1.0 / 0.0 : 0.0);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        double s = __shfl_down_sync(0xffffffffu, t, o);
        t = IS_MAX ? fmax(t, s) : t + s;
      }
      v[k] = t;
    }
  }
  __syncthreads();
}

// Returns true in ALL threads of the last block after `out[k]` holds the final values (thread 0 of
// that block wrote them); other blocks return false.  blockDim must be 1-D here (x only) with <=1024.
template <int NV, bool IS_MAX = false>
__device__ __forceinline__ bool grid_reduce(double (&v)[NV], double *partials, unsigned int *ticket,
                                            double *out) {
  __shared__ double s_red[NV * 32];
  __shared__ bool s_last;
  const int nblk = gridDim.x * gridDim.y;
  const int bid = blockIdx.y * gridDim.x + blockIdx.x;
  block_reduce<NV, IS_MAX>(v, s_red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) partials[k * TOPO_MAX_PARTIAL_BLOCKS + bid] = v[k];
    __threadfence();
    unsigned int t = atomicAdd(ticket, 1u);
    s_last = (t == (unsigned int)(nblk - 1));
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  double w[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) w[k] = IS_MAX ? -1.0 / 0.0 : 0.0;
  // fixed assignment of partials to threads, fixed order inside a thread; the slot loop is the INNER one so that a
  // thread has NV independent L2 loads in flight (slot-outer order cost the 16-slot OC pass ~10 us of serialised loads)
  for (int i = threadIdx.x; i < nblk; i += blockDim.x) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const double t = __ldcg(&partials[k * TOPO_MAX_PARTIAL_BLOCKS + i]);
      w[k] = IS_MAX ? fmax(w[k], t) : w[k] + t;
    }
  }
  block_reduce<NV, IS_MAX>(w, s_red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) out[k] = w[k];
    *ticket = 0u;
    __threadfence();
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = w[k];
  __syncthreads();
  return true;
}

__device__ __forceinline__ double2 shfl_up2(double2 v, int d) {
  return make_double2(__shfl_up_sync(0xffffffffu, v.x, d), __shfl_up_sync(0xffffffffu, v.y, d));
}
__device__ __forceinline__ double2 shfl_down2(double2 v, int d) {
  return make_double2(__shfl_down_sync(0xffffffffu, v.x, d), __shfl_down_sync(0xffffffffu, v.y, d));
}
#endif  // __CUDACC__
