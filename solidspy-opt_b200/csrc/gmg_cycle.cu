// Coarse part of the multigrid V-cycle (levels >= 1): node-stencil kernels, transfers, and the "tail" kernel that
// runs all small levels in one thread-block-cluster launch.  Set-up (Galerkin products) is in gmg.cu.  No counterpart
// in the reference (it calls SuperLU, optimize.py:437).
//
// Per level with a zero initial guess and one damped-Jacobi sweep before / after the coarse correction:
//   down:  x = w D^-1 b ; r = b - A x          (one kernel: x of the 9 neighbours is recomputed on the fly)
//          b_coarse = P^T r                     (restrict)
//   up:    x~ = x + P x_coarse ; x' = x~ + w D^-1 (b - A x~)   (one kernel, x~ of the neighbours on the fly)
// Levels with few nodes are latency/launch bound, so every level from the first with <= tail_max_nodes nodes on is
// handled by ONE launch of a 16-CTA cluster (640..768 threads per CTA): levels with more than TAIL_SMALL nodes are
// worked on by the whole cluster with barrier.cluster between stages, the smaller ones by CTA 0 alone with the
// node operators in registers, the coarsest by a dense inverse.  The row-slab drivers (topo_dist_*) at the end of
// the file run the same kernels level by level around the host's exchanges.
#include <algorithm>

#include <cooperative_groups.h>

#include "topo_internal.cuh"
#include "gmg_common.cuh"
#include "p2p_ll.cuh"

namespace cg = cooperative_groups;

bool topo_tail2_applicable(const topo_handle *h, int tail);
int topo_tail2_launch(topo_handle *h, int tail, const double2 *r_parent, int pnx, int pny, int parent_f32);

int topo_dist_mid_down(topo_handle *h);
int topo_dist_mid_up(topo_handle *h);

namespace {

inline int blocks_for(int64_t n, int threads = 128) { return (int)((n + threads - 1) / threads); }

// ---- per-node bodies shared by the multi-block kernels and the tail kernel -------------------------------
// down: x = w dinv b ; r = b - A x   (x at neighbours recomputed from b, dinv)
template <bool SYM = false>
__device__ __forceinline__ void node_down(const LevelView &L, int64_t n, double omega) {
  const int npr = L.nx + 1;
  const int64_t nn = (int64_t)npr * (L.ny + 1);
  int r, c;
  rc_of(n, npr, r, c);
  double ax, ay;
  stencil_row<SYM>(L.st, nn, n, r, c, L.nx, L.ny,
              [&](int rr, int cc) {
                const size_t m = (size_t)rr * npr + cc;
                const double2 bv = L.b[m], dv = L.dinv[m];
                return make_double2(omega * dv.x * bv.x, omega * dv.y * bv.y);
              },
              ax, ay);
  const double2 bv = L.b[n], dv = L.dinv[n];
  L.x[n] = make_double2(omega * dv.x * bv.x, omega * dv.y * bv.y);
  const double wr = row_weight(L, r);
  L.r[n] = make_double2(dv.x != 0.0 ? fma(wr, bv.x, -ax) : 0.0, dv.y != 0.0 ? fma(wr, bv.y, -ay) : 0.0);
}

// up: x~ = x + mask(P xc) ; x2 = x~ + w dinv (b - A x~)
template <bool SYM = false>
__device__ __forceinline__ void node_up(const LevelView &L, const double2 *__restrict__ xc, int cnx, int64_t n,
                                        double omega) {
  const int npr = L.nx + 1, nprc = cnx + 1;
  const int64_t nn = (int64_t)npr * (L.ny + 1);
  int r, c;
  rc_of(n, npr, r, c);
  auto xt = [&](int rr, int cc) {
    const size_t m = (size_t)rr * npr + cc;
    const double2 xv = L.x[m], dv = L.dinv[m];
    const double2 p = prolong_node(xc, cc, rr, L.nx, L.ny, nprc);
    return make_double2(dv.x != 0.0 ? xv.x + p.x : 0.0, dv.y != 0.0 ? xv.y + p.y : 0.0);
  };
  double ax, ay;
  stencil_row<SYM>(L.st, nn, n, r, c, L.nx, L.ny, xt, ax, ay);
  const double2 bv = L.b[n], dv = L.dinv[n], xv = xt(r, c);
  L.x2[n] = make_double2(fma(omega * dv.x, bv.x - ax, xv.x), fma(omega * dv.y, bv.y - ay, xv.y));
}

// plain extra smoothing sweep: x2 = x + w dinv (b - A x)
template <bool SYM = false>
__device__ __forceinline__ void node_smooth(const LevelView &L, int64_t n, double omega) {
  const int npr = L.nx + 1;
  const int64_t nn = (int64_t)npr * (L.ny + 1);
  int r, c;
  rc_of(n, npr, r, c);
  double ax, ay;
  stencil_row<SYM>(L.st, nn, n, r, c, L.nx, L.ny, [&](int rr, int cc) { return L.x[(size_t)rr * npr + cc]; }, ax, ay);
  const double2 bv = L.b[n], dv = L.dinv[n], xv = L.x[n];
  const double wr = row_weight(L, r);
  L.x2[n] = make_double2(fma(omega * dv.x, fma(wr, bv.x, -ax), xv.x), fma(omega * dv.y, fma(wr, bv.y, -ay), xv.y));
}
// residual of the current iterate (needed when nu_pre > 1): r = b - A x
template <bool SYM = false>
__device__ __forceinline__ void node_resid(const LevelView &L, int64_t n) {
  const int npr = L.nx + 1;
  const int64_t nn = (int64_t)npr * (L.ny + 1);
  int r, c;
  rc_of(n, npr, r, c);
  double ax, ay;
  stencil_row<SYM>(L.st, nn, n, r, c, L.nx, L.ny, [&](int rr, int cc) { return L.x[(size_t)rr * npr + cc]; }, ax, ay);
  const double2 bv = L.b[n], dv = L.dinv[n];
  L.r[n] = make_double2(dv.x != 0.0 ? bv.x - ax : 0.0, dv.y != 0.0 ? bv.y - ay : 0.0);
}

// unfused stages (used by the tail, where a barrier is cheap and independent loads matter more)
__device__ __forceinline__ void node_smooth0(const LevelView &L, int64_t n, double omega) {
  const double2 bv = L.b[n], dv = L.dinv[n];
  L.x[n] = make_double2(omega * dv.x * bv.x, omega * dv.y * bv.y);
}
__device__ __forceinline__ void node_prolong_add(const LevelView &L, const double2 *__restrict__ xc, int cnx, int64_t n) {
  const int npr = L.nx + 1;
  int r, c;
  rc_of(n, npr, r, c);
  const double2 p = prolong_node(xc, c, r, L.nx, L.ny, cnx + 1);
  const double2 dv = L.dinv[n];
  double2 xv = L.x[n];
  if (dv.x != 0.0) xv.x += p.x;
  if (dv.y != 0.0) xv.y += p.y;
  L.x[n] = xv;
}

// ---- the same bodies for the multi-block kernels: two-dimensional indexing, every load issued up front ------------------
// ncu on the level kernels of round 1 (profiles/r2_ncu_cycle.txt): 30-40 registers, 20-30 cycles of long-scoreboard stall
// per issued instruction, 10 % occupancy on levels 2-3 -- a thread walked its nine neighbours one after the other, each
// behind a bounds branch, and paid an L2 round trip per neighbour.  Here the neighbour index is clamped, the loads are
// predicated instead of branched around and all of them precede the arithmetic (missing neighbours: zero coefficient);
// the products are accumulated in the same order as stencil_row, so results are bit-identical.
struct Nb9 {
  int m[9];        // neighbour node (clamped to the node itself where it does not exist)
  bool ok[9];
};
__device__ __forceinline__ Nb9 nb9_of(int n, int r, int c, int nx, int ny) {
  Nb9 q;
  const int npr = nx + 1;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const int dr = k / 3 - 1, dc = k % 3 - 1;
    q.ok[k] = (r + dr >= 0) && (r + dr <= ny) && (c + dc >= 0) && (c + dc <= nx);
    q.m[k] = q.ok[k] ? n + dr * npr + dc : n;
  }
  return q;
}
template <bool SYM>
__device__ __forceinline__ void stencil_load9(const stencil_t *__restrict__ st, int nn, int n, const Nb9 &q, float4 (&cf)[9]) {
  const float4 *__restrict__ pl = reinterpret_cast<const float4 *>(st);
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    if (SYM && k < 4) cf[k] = q.ok[k] ? pl[(size_t)(8 - k) * nn + q.m[k]] : make_float4(0.f, 0.f, 0.f, 0.f);
    else cf[k] = pl[(size_t)k * nn + n];             // planes of missing neighbours hold zeros
  }
}
template <bool SYM>
__device__ __forceinline__ void stencil_dot9(const float4 (&cf)[9], const double2 (&xv)[9], double &ax, double &ay) {
  ax = 0.0;
  ay = 0.0;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    if (SYM && k < 4) {                                // transpose of the neighbour's forward block
      ax = fma((double)cf[k].x, xv[k].x, ax);
      ax = fma((double)cf[k].z, xv[k].y, ax);
      ay = fma((double)cf[k].y, xv[k].x, ay);
      ay = fma((double)cf[k].w, xv[k].y, ay);
    } else {
      ax = fma((double)cf[k].x, xv[k].x, ax);
      ax = fma((double)cf[k].y, xv[k].y, ax);
      ay = fma((double)cf[k].z, xv[k].x, ay);
      ay = fma((double)cf[k].w, xv[k].y, ay);
    }
  }
}
constexpr int LB_X = 64, LB_Y = 2;                     // block of the two-dimensional level kernels
inline dim3 lvl_grid(int nx, int ny) { return dim3((unsigned)div_up(nx + 1, LB_X), (unsigned)div_up(ny + 1, LB_Y)); }
#define LVL_RC(L)                                                  \
  const int c = blockIdx.x * LB_X + threadIdx.x;                   \
  const int r = blockIdx.y * LB_Y + threadIdx.y;                   \
  if (c > (L).nx || r > (L).ny) return;                            \
  const int npr = (L).nx + 1, nn = npr * ((L).ny + 1), n = r * npr + c

// down: x = w dinv b ; r = W b - A x   (x at the neighbours recomputed from b, dinv)
template <bool SYM>
__global__ void __launch_bounds__(LB_X * LB_Y) k_lvl_down2(const LevelView L, double omega, const int *done) {
  if (done && *done) return;
  LVL_RC(L);
  const Nb9 q = nb9_of(n, r, c, L.nx, L.ny);
  float4 cf[9];
  double2 bv[9], dv[9], xv[9];
  stencil_load9<SYM>(L.st, nn, n, q, cf);
#pragma unroll
  for (int k = 0; k < 9; ++k) { bv[k] = L.b[q.m[k]]; dv[k] = L.dinv[q.m[k]]; }
#pragma unroll
  for (int k = 0; k < 9; ++k) xv[k] = make_double2(omega * dv[k].x * bv[k].x, omega * dv[k].y * bv[k].y);
  double ax, ay;
  stencil_dot9<SYM>(cf, xv, ax, ay);
  L.x[n] = xv[4];
  const double wr = row_weight(L, r);
  L.r[n] = make_double2(dv[4].x != 0.0 ? fma(wr, bv[4].x, -ax) : 0.0, dv[4].y != 0.0 ? fma(wr, bv[4].y, -ay) : 0.0);
}
// plain sweep: x2 = x + w dinv (W b - A x)
template <bool SYM>
__global__ void __launch_bounds__(LB_X * LB_Y) k_lvl_smooth2(const LevelView L, double omega, const int *done) {
  if (done && *done) return;
  LVL_RC(L);
  const Nb9 q = nb9_of(n, r, c, L.nx, L.ny);
  float4 cf[9];
  double2 xv[9];
  stencil_load9<SYM>(L.st, nn, n, q, cf);
#pragma unroll
  for (int k = 0; k < 9; ++k) xv[k] = L.x[q.m[k]];
  const double2 bv = L.b[n], dv = L.dinv[n];
  double ax, ay;
  stencil_dot9<SYM>(cf, xv, ax, ay);
  const double wr = row_weight(L, r);
  L.x2[n] = make_double2(fma(omega * dv.x, fma(wr, bv.x, -ax), xv[4].x), fma(omega * dv.y, fma(wr, bv.y, -ay), xv[4].y));
}
// up: x~ = x + mask(P xc) ; x2 = x~ + w dinv (b - A x~)   (x~ of the neighbours on the fly)
template <bool SYM>
__global__ void __launch_bounds__(LB_X * LB_Y) k_lvl_up2(const LevelView L, const double2 *__restrict__ xc, int cnx, double omega,
                                                         const int *done) {
  if (done && *done) return;
  LVL_RC(L);
  const Nb9 q = nb9_of(n, r, c, L.nx, L.ny);
  float4 cf[9];
  double2 xv[9], dn[9];
  stencil_load9<SYM>(L.st, nn, n, q, cf);
#pragma unroll
  for (int k = 0; k < 9; ++k) { xv[k] = L.x[q.m[k]]; dn[k] = L.dinv[q.m[k]]; }
  const double2 bv = L.b[n];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const int rr = q.ok[k] ? r + k / 3 - 1 : r, cc = q.ok[k] ? c + k % 3 - 1 : c;
    const double2 p = prolong_node(xc, cc, rr, L.nx, L.ny, cnx + 1);
    xv[k] = make_double2(dn[k].x != 0.0 ? xv[k].x + p.x : 0.0, dn[k].y != 0.0 ? xv[k].y + p.y : 0.0);
  }
  double ax, ay;
  stencil_dot9<SYM>(cf, xv, ax, ay);
  L.x2[n] = make_double2(fma(omega * dn[4].x, bv.x - ax, xv[4].x), fma(omega * dn[4].y, bv.y - ay, xv[4].y));
}
__global__ void __launch_bounds__(LB_X * LB_Y) k_lvl_prolong2(const LevelView L, const double2 *__restrict__ xc, int cnx,
                                                              const int *done) {
  if (done && *done) return;
  LVL_RC(L);
  (void)nn;
  const double2 dv = L.dinv[n];
  double2 xv = L.x[n];
  const double2 p = prolong_node(xc, c, r, L.nx, L.ny, cnx + 1);
  if (dv.x != 0.0) xv.x += p.x;
  if (dv.y != 0.0) xv.y += p.y;
  L.x[n] = xv;
}
template <class V2>
__global__ void __launch_bounds__(LB_X * LB_Y) k_restrict2(const V2 *__restrict__ rf, double2 *__restrict__ bc, int nxf, int nyf,
                                                           int nxc, int nyc, const int *done) {
  if (done && *done) return;
  const int I = blockIdx.x * LB_X + threadIdx.x, J = blockIdx.y * LB_Y + threadIdx.y;
  if (I > nxc || J > nyc) return;
  bc[(size_t)J * (nxc + 1) + I] = restrict_node(rf, I, J, nxf, nyf);
}
// row-slab mode, in-kernel exchange: the threads of the first / last coarse row swap their partial sums with the
// neighbouring slabs (flag-in-data slots) and store the ACCUMULATED right-hand side -- no separate exchange kernel
template <class V2>
__global__ void __launch_bounds__(LB_X * LB_Y) k_restrict_x(const V2 *__restrict__ rf, double2 *__restrict__ bc, int nxf, int nyf,
                                                            int nxc, int nyc, const int *done, const __grid_constant__ RowXchg x) {
  if (done && *done) return;
  const unsigned long long seq = ll_rows_seq(x);
  // the blocks of the LAST row are scheduled right after those of the first: both interface rows are pushed at the start of
  // the kernel, so a neighbour that started later is waited for behind the interior work, not on top of it
  const int nby = gridDim.y, by = nby < 3 ? (int)blockIdx.y : (blockIdx.y == 0 ? 0 : (blockIdx.y == 1 ? nby - 1 : (int)blockIdx.y - 1));
  const int I = blockIdx.x * LB_X + threadIdx.x, J = by * LB_Y + threadIdx.y;
  if (I <= nxc && J <= nyc) {
    double2 b = restrict_node(rf, I, J, nxf, nyf), q;
    if (J == 0 && x.peer_lo != nullptr && ll_row_swap(x, true, I, seq, b, q)) { b.x += q.x; b.y += q.y; }
    if (J == nyc && x.peer_hi != nullptr && ll_row_swap(x, false, I, seq, b, q)) { b.x += q.x; b.y += q.y; }
    bc[(size_t)J * (nxc + 1) + I] = b;
  }
  // only the blocks that hold an interface row take part in the end-of-kernel ticket (one atomic per block of a 30 000-block
  // grid cost more than the exchange saved)
  if (by != 0 && by != nby - 1) return;
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) ll_rows_done(x, seq, gridDim.x * (gridDim.y > 1 ? 2u : 1u));
}
// fine grid: x^ += mask(P x1)   (the fine-grid x^ is a float2 array: see FK<>::F32MASK in fine.cu)
__global__ void __launch_bounds__(LB_X * LB_Y) k_prolong_add_fine2(const double2 *__restrict__ xc, float2 *__restrict__ xf,
                                                                   const uint8_t *__restrict__ fixed, int nxf, int nyf, int nxc,
                                                                   const int *done) {
  if (done && *done) return;
  const int i = blockIdx.x * LB_X + threadIdx.x, j = blockIdx.y * LB_Y + threadIdx.y;
  if (i > nxf || j > nyf) return;
  const size_t n = (size_t)j * (nxf + 1) + i;
  const unsigned m = fixed[n];
  float2 acc = xf[n];
  const double2 p = prolong_node(xc, i, j, nxf, nyf, nxc + 1);
  if (!(m & 1u)) acc.x = (float)((double)acc.x + p.x);
  if (!(m & 2u)) acc.y = (float)((double)acc.y + p.y);
  xf[n] = acc;
}

// ---- tile kernels for levels 1..tail-1 (single GPU, one sweep before / after the coarse correction) -------------------
// The separate transfer kernels cost a launch each (6-10 us at these sizes) and a round trip of b / x through memory.
// Here a block owns a TX x TY tile of nodes and first forms the smoothing INPUT of the tile and of its one-node halo ring
// in shared memory -- going down: b = P^T r_parent (restriction on the fly) and x = w D^-1 b; going up: x~ = w D^-1 b +
// mask(P x_coarse) -- so every transfer value is computed once per tile (+ 2(TX+TY+2) halo entries) instead of once per
// stencil neighbour, and the nine neighbour values of the stencil product are shared-memory reads.  x = w D^-1 b is never
// stored: the up kernel recomputes it from b and D^-1, which it reads anyway.  Same operations in the same order as
// k_restrict2 + k_lvl_down2 / k_lvl_prolong2 + k_lvl_smooth2: results are bit-identical.
constexpr int TLX = 32, TLY = 8;
__device__ __forceinline__ void tile_halo_slot(int h, int &hy, int &hx) {
  if (h < TLX + 2) { hy = 0; hx = h; }
  else if (h < 2 * (TLX + 2)) { hy = TLY + 1; hx = h - (TLX + 2); }
  else { const int k = h - 2 * (TLX + 2); hy = 1 + (k >> 1); hx = (k & 1) ? TLX + 1 : 0; }
}
constexpr int TL_HALO = 2 * (TLX + 2) + 2 * TLY;
inline dim3 tile_grid(int nx, int ny) { return dim3((unsigned)div_up(nx + 1, TLX), (unsigned)div_up(ny + 1, TLY)); }

template <bool SYM>
__device__ __forceinline__ void tile_apply(const LevelView &L, double2 (*xs)[TLX + 2], int n, int r, int c, int nn, double &ax,
                                           double &ay, double2 &xc) {
  const Nb9 q = nb9_of(n, r, c, L.nx, L.ny);
  float4 cf[9];
  double2 xv[9];
  stencil_load9<SYM>(L.st, nn, n, q, cf);
#pragma unroll
  for (int k = 0; k < 9; ++k) xv[k] = xs[threadIdx.y + k / 3][threadIdx.x + k % 3];
  stencil_dot9<SYM>(cf, xv, ax, ay);
  xc = xv[4];
}

// down: b = P^T r_parent ; x = w D^-1 b (not stored) ; r = W b - A x
template <bool SYM, class V2>
__global__ void __launch_bounds__(TLX * TLY) k_lvl_down_t(const LevelView L, const V2 *__restrict__ rp, int pnx, int pny, double omega,
                                                          const int *done) {
  if (done && *done) return;
  __shared__ double2 xs[TLY + 2][TLX + 2];
  const int c0 = blockIdx.x * TLX, r0 = blockIdx.y * TLY;
  const int c = c0 + threadIdx.x, r = r0 + threadIdx.y;
  const int npr = L.nx + 1, nn = npr * (L.ny + 1), n = r * npr + c;
  const bool own = c <= L.nx && r <= L.ny;
  double2 bv = make_double2(0.0, 0.0), dv = bv;
  if (own) {
    bv = restrict_node(rp, c, r, pnx, pny);
    dv = L.dinv[n];
    L.b[n] = bv;
  }
  xs[threadIdx.y + 1][threadIdx.x + 1] = make_double2(omega * dv.x * bv.x, omega * dv.y * bv.y);
  for (int h = threadIdx.y * TLX + threadIdx.x; h < TL_HALO; h += TLX * TLY) {
    int hy, hx;
    tile_halo_slot(h, hy, hx);
    const int rr = r0 + hy - 1, cc = c0 + hx - 1;
    double2 xv = make_double2(0.0, 0.0);
    if (rr >= 0 && rr <= L.ny && cc >= 0 && cc <= L.nx) {
      const double2 b2 = restrict_node(rp, cc, rr, pnx, pny), d2 = L.dinv[rr * npr + cc];
      xv = make_double2(omega * d2.x * b2.x, omega * d2.y * b2.y);
    }
    xs[hy][hx] = xv;
  }
  __syncthreads();
  if (!own) return;
  double ax, ay;
  double2 xc;
  tile_apply<SYM>(L, xs, n, r, c, nn, ax, ay, xc);
  const double wr = row_weight(L, r);
  L.r[n] = make_double2(dv.x != 0.0 ? fma(wr, bv.x, -ax) : 0.0, dv.y != 0.0 ? fma(wr, bv.y, -ay) : 0.0);
}

// up: x~ = w D^-1 b + mask(P xc) ; x2 = x~ + w D^-1 (b - A x~)
// XCH: row-slab mode (row weights, interface rows of x~ stored or swapped); false: the single-GPU instantiation carries none of it
// (bx, by) = tile; xs = the block's (TLY + 2) x (TLX + 2) shared tile.  Ends WITHOUT a block barrier.
template <bool SYM, bool XCH>
__device__ __forceinline__ void lvl_up_tile(const LevelView &L, const double2 *__restrict__ xcoarse, int cnx, double omega,
                                            const RowXchg &x, unsigned long long seq, int bx, int by, double2 (*xs)[TLX + 2]) {
  const int c0 = bx * TLX, r0 = by * TLY;
  const int c = c0 + threadIdx.x, r = r0 + threadIdx.y;
  const int npr = L.nx + 1, nn = npr * (L.ny + 1), n = r * npr + c;
  const bool own = c <= L.nx && r <= L.ny;
  auto xtilde = [&](int rr, int cc, double2 &b2, double2 &d2) {
    const int m = rr * npr + cc;
    b2 = L.b[m];
    d2 = L.dinv[m];
    const double2 p = prolong_node(xcoarse, cc, rr, L.nx, L.ny, cnx + 1);
    // (w d) b is rounded before the coarse correction is added, exactly like the stored x of the separate kernels
    return make_double2(d2.x != 0.0 ? __dadd_rn(__dmul_rn(omega * d2.x, b2.x), p.x) : 0.0,
                        d2.y != 0.0 ? __dadd_rn(__dmul_rn(omega * d2.y, b2.y), p.y) : 0.0);
  };
  double2 bv = make_double2(0.0, 0.0), dv = bv, x0 = bv;
  if (own) x0 = xtilde(r, c, bv, dv);
  xs[threadIdx.y + 1][threadIdx.x + 1] = x0;
  for (int h = threadIdx.y * TLX + threadIdx.x; h < TL_HALO; h += TLX * TLY) {
    int hy, hx;
    tile_halo_slot(h, hy, hx);
    const int rr = r0 + hy - 1, cc = c0 + hx - 1;
    double2 xv = make_double2(0.0, 0.0), b2, d2;
    if (rr >= 0 && rr <= L.ny && cc >= 0 && cc <= L.nx) xv = xtilde(rr, cc, b2, d2);
    xs[hy][hx] = xv;
  }
  __syncthreads();
  if (own) {
    double ax, ay;
    double2 xc;
    tile_apply<SYM>(L, xs, n, r, c, nn, ax, ay, xc);
    // row-slab mode: W b makes the distributed copy of b on interface rows (W = 1 elsewhere: fma(1, b, -ax) == b - ax).
    // An interface row of x2 is completed by x2 + x2(neighbour) - x~: here, by swapping the node with the neighbour
    // (in-kernel exchange), or by the host's exchange, for which those rows of x~ are stored
    const double wr = XCH ? row_weight(L, r) : 1.0;
    double2 v = make_double2(fma(omega * dv.x, fma(wr, bv.x, -ax), xc.x), fma(omega * dv.y, fma(wr, bv.y, -ay), xc.y)), q;
    if (XCH && wr != 1.0) {
      if (x.c == nullptr) L.x[n] = xc;
      else if (ll_row_swap(x, r == 0, c, seq, v, q)) { v.x = (v.x + q.x) - xc.x; v.y = (v.y + q.y) - xc.y; }
    }
    L.x2[n] = v;
  }
}
template <bool SYM, bool XCH>
__global__ void __launch_bounds__(TLX * TLY) k_lvl_up_t(const LevelView L, const double2 *__restrict__ xcoarse, int cnx, double omega,
                                                        const int *done, const __grid_constant__ RowXchg x) {
  if (done && *done) return;
  const unsigned long long seq = (XCH && x.c != nullptr) ? ll_rows_seq(x) : 0ull;
  __shared__ double2 xs[TLY + 2][TLX + 2];
  // row-slab mode: the tiles of the last tile row right after those of the first (see k_restrict_x)
  const int nby = gridDim.y;
  const int by = (!XCH || nby < 3) ? (int)blockIdx.y : (blockIdx.y == 0 ? 0 : (blockIdx.y == 1 ? nby - 1 : (int)blockIdx.y - 1));
  lvl_up_tile<SYM, XCH>(L, xcoarse, cnx, omega, x, seq, blockIdx.x, by, xs);
  if (XCH && x.c != nullptr && (by == 0 || by == nby - 1)) {      // (blocks that hold an interface row)
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) ll_rows_done(x, seq, gridDim.x * (gridDim.y > 1 ? 2u : 1u));
  }
}

// ---- row-slab mode: the SMALL slab-local levels in one launch per direction ------------------------------------------------
// On 8 GPUs the slab-local levels 3..5 of an 8192 x 4096 mesh have 66 000, 17 000 and 4 400 nodes; each of their kernels
// (restriction, sweep, up) takes 6-10 us whatever its size, ten launches and six in-kernel row swaps per V-cycle.  Here one
// persistent grid (one block per SM, all resident) walks through the phases with a grid barrier between them:
//   down:  for l = Lm..K-1:  b_l = P^T r_{l-1} (+ row swap)  |  r_l = W b_l - A_loc (w D^-1 b_l)  ;  then the level-K piece
//          P^T r_{K-1} is pushed to every rank (flag-in-data) and the global level-K right-hand side is assembled
//   up:    for l = K-1..Lm:  x2_l = tile up-kernel body (+ row swap)
// Row swaps of consecutive phases use consecutive sequence numbers (alternating slot parity); a rank cannot be two swaps
// ahead of its neighbour because every swap is followed by a grid barrier on both sides.
struct MidArgs {
  LevelView lv[8];
  int n;                                   // levels Lm .. K-1
  const double2 *r_parent;                 // residual of level Lm-1 (down) / result of level K on this slab's rows (up)
  int pnx, pny;                            // its size (down) ; pnx = cells per row of level K (up)
  double omega;
  const int *done;
  RowXchg x;
  unsigned int *bar;                       // [0] arrivals, [1] generation
  // level-K piece and gather (down only)
  int nxK, nyK;                            // local level-K piece: nxK x nyK cells
  PeerTable peers;
  int my_rank, world;
  char *region;
  size_t off_gather, gstride;
  double2 *bK;                             // assembled global right-hand side of level K
};
__device__ __forceinline__ void grid_barrier(unsigned int *bar, unsigned int nblocks) {
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    __threadfence();
    const unsigned int gen = *reinterpret_cast<volatile unsigned int *>(bar + 1);
    if (atomicAdd(bar, 1u) == nblocks - 1u) {
      bar[0] = 0u;
      __threadfence();
      atomicAdd(bar + 1, 1u);
    } else {
      while (*reinterpret_cast<volatile unsigned int *>(bar + 1) == gen) __nanosleep(32);
    }
    __threadfence();
  }
  __syncthreads();
}
__global__ void __launch_bounds__(TLX * TLY) k_slab_mid_down(const __grid_constant__ MidArgs a) {
  if (a.done && *a.done) return;
  P2PCounters *c = a.x.c;
  const int T = TLX * TLY, tid = threadIdx.y * TLX + threadIdx.x;
  const int gt = blockIdx.x * T + tid, GT = gridDim.x * T;
  const unsigned long long seq0 = ll_rows_seq(a.x), gseq = c->gather_push + 1;
  const double2 *rp = a.r_parent;
  int pnx = a.pnx, pny = a.pny;
  for (int k = 0; k < a.n; ++k) {
    const LevelView &L = a.lv[k];
    const int npr = L.nx + 1, nn = npr * (L.ny + 1);
    for (int n = gt; n < nn; n += GT) {
      int J, I;
      rc_of(n, npr, J, I);
      double2 b = restrict_node(rp, I, J, pnx, pny), q;
      if (J == 0 && a.x.peer_lo != nullptr && ll_row_swap(a.x, true, I, seq0 + k, b, q)) { b.x += q.x; b.y += q.y; }
      if (J == L.ny && a.x.peer_hi != nullptr && ll_row_swap(a.x, false, I, seq0 + k, b, q)) { b.x += q.x; b.y += q.y; }
      L.b[n] = b;
    }
    grid_barrier(a.bar, gridDim.x);
    if (L.sym) { for (int n = gt; n < nn; n += GT) node_down<true>(L, n, a.omega); }
    else { for (int n = gt; n < nn; n += GT) node_down<false>(L, n, a.omega); }
    grid_barrier(a.bar, gridDim.x);
    rp = L.r; pnx = L.nx; pny = L.ny;
  }
  // level-K piece: restricted, pushed to every rank's gather slot [my rank] (partial sums on the interface rows stay partial)
  {
    const unsigned s32 = (unsigned)gseq;
    const int nprK = a.nxK + 1;
    const long long npiece = (long long)nprK * (a.nyK + 1);
    for (long long i = gt; i < npiece; i += GT) {
      int J, I;
      rc_of(i, nprK, J, I);
      const double2 v = restrict_node(rp, I, J, pnx, pny);
      for (int r = 0; r < a.world; ++r) {
        uint4 *dst = reinterpret_cast<uint4 *>(a.peers.p[r] + a.off_gather + (size_t)(gseq & 1ull) * a.gstride) + 2 * ((size_t)a.my_rank * npiece + i);
        ll_store(dst, v.x, s32);
        ll_store(dst + 1, v.y, s32);
      }
    }
    // assemble the global right-hand side (every rank, identically): rank g holds coarse rows [g nyK, (g+1) nyK]
    const uint4 *__restrict__ recv = reinterpret_cast<const uint4 *>(a.region + a.off_gather + (size_t)(gseq & 1ull) * a.gstride);
    const long long nn = (long long)nprK * ((long long)a.nyK * a.world + 1);
    for (long long m = gt; m < nn; m += GT) {
      int J, I;
      rc_of(m, nprK, J, I);
      const int g = J / a.nyK, j = J - g * a.nyK;
      double2 acc = make_double2(0.0, 0.0);
      bool ok = true;
      if (g < a.world) {
        const uint4 *w = recv + 2 * ((long long)g * npiece + (long long)j * nprK + I);
        ok = ll_load(w, s32, acc.x, &c->error) && ll_load(w + 1, s32, acc.y, &c->error);
      }
      if (ok && j == 0 && g > 0) {
        const uint4 *w = recv + 2 * ((long long)(g - 1) * npiece + (long long)a.nyK * nprK + I);
        double qx, qy;
        ok = ll_load(w, s32, qx, &c->error) && ll_load(w + 1, s32, qy, &c->error);
        acc.x += qx; acc.y += qy;
      }
      if (ok) a.bK[m] = acc;
    }
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(&c->fticket[3], 1u) == gridDim.x - 1u) {
      c->fticket[3] = 0u;
      if (a.n > 0) { c->halo_push = seq0 + a.n - 1; c->halo_wait = seq0 + a.n - 1; }
      c->gather_push = gseq;
      c->gather_wait = gseq;
    }
  }
}
__global__ void __launch_bounds__(TLX * TLY) k_slab_mid_up(const __grid_constant__ MidArgs a) {
  if (a.done && *a.done) return;
  P2PCounters *c = a.x.c;
  __shared__ double2 xs[TLY + 2][TLX + 2];
  const unsigned long long seq0 = ll_rows_seq(a.x);
  const double2 *xc = a.r_parent;
  int cnx = a.pnx;
  for (int k = a.n - 1, ph = 0; k >= 0; --k, ++ph) {
    const LevelView &L = a.lv[k];
    const int tx = (L.nx + TLX) / TLX, ty = (L.ny + TLY) / TLY;
    for (int t = blockIdx.x; t < tx * ty; t += gridDim.x) {
      if (L.sym) lvl_up_tile<true, true>(L, xc, cnx, a.omega, a.x, seq0 + ph, t % tx, t / tx, xs);
      else lvl_up_tile<false, true>(L, xc, cnx, a.omega, a.x, seq0 + ph, t % tx, t / tx, xs);
      __syncthreads();
    }
    if (k > 0) grid_barrier(a.bar, gridDim.x);
    xc = L.x2; cnx = L.nx;
  }
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    __threadfence();
    if (atomicAdd(&c->fticket[2], 1u) == gridDim.x - 1u) {
      c->fticket[2] = 0u;
      if (a.n > 0) { c->halo_push = seq0 + a.n - 1; c->halo_wait = seq0 + a.n - 1; }
    }
  }
}

// ---- multi-block kernels (large coarse levels): the two-dimensional kernels above + the residual of an iterate ------
template <bool SYM>
__global__ void __launch_bounds__(128) k_lvl_resid(const LevelView L, const int *done) {
  if (done && *done) return;
  const int64_t nn = (int64_t)(L.nx + 1) * (L.ny + 1);
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n < nn) node_resid<SYM>(L, n);
}
// ---- tail: all levels >= first in ONE launch of a thread-block cluster ------------------------------------
// Levels with more than TAIL_SMALL nodes are worked on by every CTA of the cluster with a hardware
// cluster barrier (barrier.cluster, release/acquire: global writes of one CTA are visible to the others
// after it) between stages; the smallest levels are finished by CTA 0 alone with __syncthreads.
constexpr int TAIL_MAX = 16;
#ifndef TOPO_TAIL_THREADS
#define TOPO_TAIL_THREADS 768
#endif
constexpr int TAIL_THREADS = TOPO_TAIL_THREADS;     // 80 registers per thread: the small-level path keeps a node's operator in registers
constexpr int TAIL_SMALL = TAIL_THREADS;
struct TailArgs {
  LevelView lv[TAIL_MAX];
  int n;                       // number of levels in the tail (the last one is the coarsest)
  const double2 *r_parent;     // residual of the level above the tail (a float2 array when that level is the fine grid)
  int parent_f32;
  int pnx, pny;                // its size
  const double *coarse_inv;
  int coarse_n;
  double omega;
  int nu[TAIL_MAX];            // smoothing sweeps (pre = post) per tail level
  const int *done;
  int *err;                    // SolverScalars::tail_error
  unsigned long long *dbg;     // optional: globaltimer stamps at stage boundaries (thread 0 of CTA 0)
};

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long v;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
  return v;
}
#define STAMP() do { if (a.dbg && gt == 0) { a.dbg[1 + a.dbg[0]] = gtimer(); a.dbg[0] += 1; } } while (0)

// nu pre-smoothing sweeps from a zero guess, then the residual, then its restriction.  Sweep 1 is pointwise
// (x = w D^-1 b); for nu == 1 it is fused with the residual (node_down).  Sweeps ping-pong x <-> x2 through local
// pointer swaps; the iterate ends in x for odd nu and in x2 for even nu -- recorded in cur[l].
// does tail level l start its pre-smoothing from x0 = w D^-1 b stored by the stage that produced b?  (cluster
// levels with more than one sweep: the pointwise first sweep rides in the producer's stage instead of costing a
// cluster barrier of its own)
__device__ __forceinline__ bool tail_wants_x0(const TailArgs &a, int l) {
  return l + 1 < a.n && (a.lv[l].nx + 1) * (a.lv[l].ny + 1) > TAIL_SMALL && a.nu[l] > 1;
}
__device__ __forceinline__ void tail_store_b(const TailArgs &a, int l, int n, double2 bv, bool x0) {
  const LevelView &C = a.lv[l];
  C.b[n] = bv;
  if (x0) {
    const double2 dv = C.dinv[n];
    C.x[n] = make_double2(a.omega * dv.x * bv.x, a.omega * dv.y * bv.y);
  }
}
template <class Sync>
__device__ __forceinline__ void tail_down(const TailArgs &a, int l, double2 **cur, int gt, int GT, Sync sync, bool x0_ready) {
  LevelView L = a.lv[l];
  const int nn = (L.nx + 1) * (L.ny + 1);
  const int nu = a.nu[l];
  if (nu == 1) {
    for (int n = gt; n < nn; n += GT) node_down(L, n, a.omega);          // x = w dinv b ; r = b - A x
    sync();
    STAMP();
  } else {
    if (!x0_ready) {
      for (int n = gt; n < nn; n += GT) node_smooth0(L, n, a.omega);
      sync();
    }
    for (int s = 1; s < nu; ++s) {
      for (int n = gt; n < nn; n += GT) node_smooth(L, n, a.omega);      // L.x -> L.x2
      sync();
      double2 *t = L.x; L.x = L.x2; L.x2 = t;
    }
    for (int n = gt; n < nn; n += GT) node_resid(L, n);                  // r = b - A (current iterate)
    sync();
    STAMP();
  }
  cur[l] = L.x;
  const LevelView &C = a.lv[l + 1];
  const int nprc = C.nx + 1, nnc = nprc * (C.ny + 1);
  const bool x0 = tail_wants_x0(a, l + 1);
  for (int n = gt; n < nnc; n += GT) tail_store_b(a, l + 1, n, restrict_node(L.r, n % nprc, n / nprc, L.nx, L.ny), x0);
  sync();
  STAMP();
}

// coarse correction + nu post-smoothing sweeps; the final iterate pointer is left in cur[l]
template <class Sync>
__device__ __forceinline__ void tail_up(const TailArgs &a, int l, double2 **cur, int gt, int GT, Sync sync) {
  LevelView L = a.lv[l];
  const LevelView &C = a.lv[l + 1];
  if (cur[l] != L.x) { L.x2 = L.x; L.x = cur[l]; }                       // iterate left by tail_down
  const int nn = (L.nx + 1) * (L.ny + 1);
  for (int n = gt; n < nn; n += GT) node_prolong_add(L, cur[l + 1], C.nx, n);
  sync();
  STAMP();
  for (int s = 0; s < a.nu[l]; ++s) {
    for (int n = gt; n < nn; n += GT) node_smooth(L, n, a.omega);        // x -> x2
    sync();
    STAMP();
    double2 *t = L.x; L.x = L.x2; L.x2 = t;
  }
  cur[l] = L.x;
}

// ---- small levels (<= TAIL_THREADS nodes): one thread per node, operator in registers -------------------------------------
// A sweep of the generic path costs ~400 instructions per node (index arithmetic, bounds, 9 stencil + 9 vector
// loads, conversions) and the 18 warps of a 561-node level are ISSUE bound on one SM: ~1 us per stage.  Here a
// thread keeps its node's 9 stencil blocks (fp32, as stored), D^-1 and b in registers for all sweeps of a
// phase; the iterate travels between threads as an fp32 copy in shared memory, so a sweep is 9 LDS.64 + 36 FFMA
// + the fp64 update + one barrier.  The products A x are therefore formed in fp32 (operands rounded like the
// stencils already are); iterate, right-hand side and residual stay fp64.  Result of a level: always in L.x.
struct SmallNode {
  float4 st[9];
  double2 dinv, b, x;
  int n, npr, valid;          // valid = last node index (clamp for neighbours outside the level)
};
__device__ __forceinline__ void small_load(const LevelView &L, int t, SmallNode &N) {
  const int npr = L.nx + 1, nn = npr * (L.ny + 1);
  N.n = t; N.npr = npr; N.valid = nn - 1;
  if (t >= nn) return;
  const int r = t / npr, c = t - r * npr;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const int rr = r + k / 3 - 1, cc = c + k % 3 - 1;
    const bool ok = rr >= 0 && rr <= L.ny && cc >= 0 && cc <= L.nx;
    N.st[k] = ok ? __ldg(reinterpret_cast<const float4 *>(L.st) + (size_t)k * nn + t) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  N.dinv = L.dinv[t];
  N.b = L.b[t];
}
// (A x)(node) from the fp32 copy of the iterate
__device__ __forceinline__ void small_apply(const SmallNode &N, const float2 *__restrict__ xs, float &ax, float &ay) {
  ax = 0.f; ay = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const int m = min(max(N.n + (k / 3 - 1) * N.npr + (k % 3 - 1), 0), N.valid);   // missing neighbours: zero coefficient
    const float2 xv = xs[m];
    ax = fmaf(N.st[k].x, xv.x, ax); ax = fmaf(N.st[k].y, xv.y, ax);
    ay = fmaf(N.st[k].z, xv.x, ay); ay = fmaf(N.st[k].w, xv.y, ay);
  }
}
__device__ __forceinline__ void small_sweeps(SmallNode &N, bool active, float2 (*xs)[TAIL_THREADS], int &cur, int nsweeps, double omega) {
  for (int s = 0; s < nsweeps; ++s) {
    if (active) {
      float ax, ay;
      small_apply(N, xs[cur], ax, ay);
      N.x.x = fma(omega * N.dinv.x, N.b.x - (double)ax, N.x.x);
      N.x.y = fma(omega * N.dinv.y, N.b.y - (double)ay, N.x.y);
      xs[cur ^ 1][N.n] = make_float2((float)N.x.x, (float)N.x.y);
    }
    __syncthreads();
    cur ^= 1;
  }
}
// pre-smoothing from a zero guess (nu sweeps), residual, restriction into the next level's b
__device__ __noinline__ void small_down(const TailArgs &a, int l, float2 (*xs)[TAIL_THREADS], int t, int T) {
  const LevelView &L = a.lv[l];
  const int nn = (L.nx + 1) * (L.ny + 1);
  const bool active = t < nn;
  SmallNode N;
  small_load(L, t, N);
  int cur = 0;
  if (active) {
    N.x = make_double2(a.omega * N.dinv.x * N.b.x, a.omega * N.dinv.y * N.b.y);
    xs[0][t] = make_float2((float)N.x.x, (float)N.x.y);
  }
  __syncthreads();
  small_sweeps(N, active, xs, cur, a.nu[l] - 1, a.omega);
  if (active) {
    float ax, ay;
    small_apply(N, xs[cur], ax, ay);
    L.x[t] = N.x;
    L.r[t] = make_double2(N.dinv.x != 0.0 ? N.b.x - (double)ax : 0.0, N.dinv.y != 0.0 ? N.b.y - (double)ay : 0.0);
  }
  __syncthreads();
  const LevelView &C = a.lv[l + 1];
  const int nprc = C.nx + 1, nnc = nprc * (C.ny + 1);
  for (int n = t; n < nnc; n += T) C.b[n] = restrict_node(L.r, n % nprc, n / nprc, L.nx, L.ny);
  __syncthreads();
}
// coarse correction + nu post-smoothing sweeps
__device__ __noinline__ void small_up(const TailArgs &a, int l, float2 (*xs)[TAIL_THREADS], int t) {
  const LevelView &L = a.lv[l];
  const LevelView &C = a.lv[l + 1];
  const int npr = L.nx + 1, nn = npr * (L.ny + 1);
  const bool active = t < nn;
  SmallNode N;
  small_load(L, t, N);
  int cur = 0;
  if (active) {
    const double2 p = prolong_node(C.x, t % npr, t / npr, L.nx, L.ny, C.nx + 1);   // small levels leave their result in x
    N.x = L.x[t];
    if (N.dinv.x != 0.0) N.x.x += p.x;
    if (N.dinv.y != 0.0) N.x.y += p.y;
    xs[0][t] = make_float2((float)N.x.x, (float)N.x.y);
  }
  __syncthreads();
  small_sweeps(N, active, xs, cur, a.nu[l], a.omega);
  if (active) L.x[t] = N.x;
  __syncthreads();
}

__global__ void __launch_bounds__(TAIL_THREADS) k_tail(const __grid_constant__ TailArgs a) {
  if (a.done && *a.done) return;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank(), CS = (int)cluster.num_blocks();
  const int T = blockDim.x, t = threadIdx.x;
  const int gt = rank * T + t, GT = CS * T;
  auto csync = [&]() { cluster.sync(); };
  auto bsync = [&]() { __syncthreads(); };
  if (a.dbg && gt == 0) a.dbg[0] = 0;
  STAMP();
  double2 *cur[TAIL_MAX];                       // where each level's current iterate lives (x or x2)
  if (a.r_parent != nullptr) {                  // row-slab mode: lv[0].b was assembled from the gathered slabs
    const LevelView &C = a.lv[0];
    const int nprc = C.nx + 1, nnc = nprc * (C.ny + 1);
    const bool x0 = tail_wants_x0(a, 0);
    if (a.parent_f32) {
      const float2 *rp = reinterpret_cast<const float2 *>(a.r_parent);
      for (int n = gt; n < nnc; n += GT) tail_store_b(a, 0, n, restrict_node(rp, n % nprc, n / nprc, a.pnx, a.pny), x0);
    } else {
      for (int n = gt; n < nnc; n += GT) tail_store_b(a, 0, n, restrict_node(a.r_parent, n % nprc, n / nprc, a.pnx, a.pny), x0);
    }
  }
  csync();
  STAMP();
  int l = 0;
  // (row-slab mode: the first level's b was assembled by another kernel, so its x0 is not there yet)
  for (; l + 1 < a.n && (a.lv[l].nx + 1) * (a.lv[l].ny + 1) > TAIL_SMALL; ++l)
    tail_down(a, l, cur, gt, GT, csync, (l > 0 || a.r_parent != nullptr) && tail_wants_x0(a, l));
  const int lsmall = l;
  // every rank must know where the iterates of the big levels ended (same arithmetic on all ranks: cur[] is
  // computed identically); the small levels are finished by CTA 0, which publishes cur[lsmall] via the rule below
  if (rank == 0) {
    __shared__ float2 xs[2][TAIL_THREADS];
    for (l = lsmall; l + 1 < a.n; ++l) { small_down(a, l, xs, t, T); STAMP(); }
    {  // coarsest: dense inverse, one warp per row (lanes stride the columns, shuffle reduction)
      const LevelView &C = a.lv[a.n - 1];
      const double *bb = reinterpret_cast<const double *>(C.b);
      const double *dd = reinterpret_cast<const double *>(C.dinv);
      double *xx = reinterpret_cast<double *>(C.x);
      const int N = a.coarse_n;
      const int lane = t & 31, warp = t >> 5, nwarp = T >> 5;
      for (int i = warp; i < N; i += nwarp) {
        double sacc = 0.0;
        for (int j = lane; j < N; j += 32) sacc = fma(a.coarse_inv[i * N + j], bb[j], sacc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sacc += __shfl_down_sync(0xffffffffu, sacc, o);
        if (lane == 0) xx[i] = dd[i] != 0.0 ? sacc : 0.0;
      }
      cur[a.n - 1] = C.x;
    }
    __syncthreads();
    STAMP();
    for (l = a.n - 2; l >= lsmall; --l) { small_up(a, l, xs, t); STAMP(); }
  }
  // small levels (and the coarsest) leave their result in x
  if (lsmall < a.n) cur[lsmall] = a.lv[lsmall].x;
  csync();
  STAMP();
  for (l = lsmall - 1; l >= 0; --l) tail_up(a, l, cur, gt, GT, csync);
}

const char *lvl_name(const char *what, int l) {
  static char names[2][TOPO_MAX_LEVELS][24];
  static bool init = false;
  if (!init) {
    for (int i = 0; i < TOPO_MAX_LEVELS; ++i) {
      snprintf(names[0][i], 24, "k_lvl_down L%d", i);
      snprintf(names[1][i], 24, "k_lvl_up L%d", i);
    }
    init = true;
  }
  return names[what[0] == 'd' ? 0 : 1][l];
}

LevelView view_of(const topo_handle *h, const GmgLevel &G) {
  LevelView v;
  v.st = G.st; v.dinv = G.dinv; v.x = G.x; v.x2 = G.x2; v.b = G.b; v.r = G.r; v.nx = G.nx; v.ny = G.ny;
  v.wf = v.wl = 1.0;
  // levels whose stencil stream is what bounds them (TOPO_SYM_MIN nodes and more; default 100 000) read the symmetric
  // half only: 2048x1024 8.9 -> 8.6 ms per design iteration, 8192x4096 66.8 -> 64.3 ms, identical trajectories
  v.sym = G.nn >= h->knobs.sym_min ? 1 : 0;
  return v;
}

// sweeps per tail level: the deep, tiny levels are where the coarse solve loses accuracy on high-contrast
// designs and where a sweep costs next to nothing
int tail_nu(const topo_handle *h, int l) {
  return h->levels[l].nn > TAIL_SMALL ? h->nu_tail_big : h->nu_tail_small;
}
int nu_pre_of(const topo_handle *h, int l) {
  const int nuc_max = h->knobs.nu_coarse_maxlvl;
  return (h->nu_coarse > 0 && l <= nuc_max) ? h->nu_coarse : h->nu_pre;
}
int nu_post_of(const topo_handle *h, int l) {
  const int nuc_max = h->knobs.nu_coarse_maxlvl;
  return (h->nu_coarse > 0 && l <= nuc_max) ? h->nu_coarse : h->nu_post;
}
// first level handled by the tail kernel (row-slab mode: the first global level)
int tail_start(const topo_handle *h) {
  if (h->dist.on) return h->dist.K;
  int tail = h->n_levels - 1;
  while (tail > 1 && h->levels[tail - 1].nn <= h->tail_max_nodes) --tail;
  return tail;
}
// largest cluster the device schedules for the tail kernels (16 needs the non-portable size; else 8, 4, 2, 1)
template <class Kern>
int probe_cluster_size(Kern kern, size_t dyn_smem, int cs_max) {
  int cluster_size = 1;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
    for (int cs : {16, 8, 4, 2}) {
      if (cs > cs_max) continue;
      cudaLaunchConfig_t q = {};
      q.gridDim = dim3(cs); q.blockDim = dim3(TAIL_THREADS); q.dynamicSmemBytes = dyn_smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      q.attrs = at; q.numAttrs = 1;
      int ncl = 0;
      if (cudaOccupancyMaxActiveClusters(&ncl, kern, &q) == cudaSuccess && ncl >= 1) { cluster_size = cs; break; }
    }
  }
  cudaGetLastError();
  return cluster_size;
}
int tail_cluster_size(const topo_handle *h) {
  static int cs_dev[TOPO_MAX_DEVICES] = {};
  int &cs = cs_dev[topo_dev_slot(h)];
  if (cs == 0) cs = probe_cluster_size(k_tail, 0, h->knobs.tail_cluster_max);
  return cs;
}
// where the final iterate of level l lives after the cycle
const double2 *level_result(const topo_handle *h, int l, int tail) {
  const int L = h->n_levels;
  if (l == L - 1) return h->levels[l].x;          // coarsest: dense solve writes x
  if (l >= tail) {                                 // tail levels: parity of the pointer swaps (see k_tail)
    if (h->levels[l].nn <= TAIL_SMALL) return h->levels[l].x;      // small levels keep their iterate in x
    if (topo_tail2_applicable(h, tail)) return h->levels[l].x;     // k_tail2 writes the first tail level's result to x
    const int nu = tail_nu(h, l);
    const int swaps = (nu > 1 ? nu - 1 : 0) + nu;
    return (swaps & 1) ? h->levels[l].x2 : h->levels[l].x;
  }
  return (nu_post_of(h, l) & 1) ? h->levels[l].x2 : h->levels[l].x;
}

// all levels >= tail in one cluster launch; r_parent == nullptr: levels[tail].b is already in place
int launch_tail(topo_handle *h, int tail, const double2 *r_parent, int pnx, int pny, int parent_f32 = 0) {
  const int L = h->n_levels;
  TailArgs ta = {};
  ta.n = L - tail;
  if (ta.n > TAIL_MAX) {
    topo_set_error("multigrid tail has %d levels (max %d)", ta.n, TAIL_MAX);
    return TOPO_EINVAL;
  }
  for (int l = tail; l < L; ++l) ta.lv[l - tail] = view_of(h, h->levels[l]);
  ta.r_parent = r_parent; ta.pnx = pnx; ta.pny = pny; ta.parent_f32 = parent_f32;
  ta.coarse_inv = h->coarse_inv;
  ta.coarse_n = h->coarse_n;
  ta.omega = h->omega;
  for (int l = tail; l < L; ++l) ta.nu[l - tail] = tail_nu(h, l);
  ta.done = &h->sc->done;
  ta.err = &h->sc->tail_error;
  ta.dbg = h->tail_dbg;
  const int cluster_size = tail_cluster_size(h);
  if (h->inv_pending) {     // join the side stream that inverts the coarsest operator (first tail launch of a solve)
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(h->stream, &cap);
    if (cap == cudaStreamCaptureStatusNone) {
      CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_inv, 0));
      h->inv_pending = false;
    }
  }
  if (topo_tail2_applicable(h, tail)) return topo_tail2_launch(h, tail, r_parent, pnx, pny, parent_f32);
  const bool big_tail = h->levels[tail].nn > TAIL_SMALL;
  const int cs = big_tail ? cluster_size : 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cs);
  cfg.blockDim = dim3(TAIL_THREADS);
  cfg.stream = h->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  KL(h, "k_tail", CUDA_TRY(cudaLaunchKernelEx(&cfg, k_tail, ta)));
  return TOPO_OK;
}

// ---- row-slab mode helpers ---------------------------------------------------------------------------------
// global level-K right-hand side from the gathered per-rank pieces: rank g holds coarse rows
// [g*rows, (g+1)*rows] (rows+1 node rows, the shared ones are partial sums)
// (peer-memory transport: the pieces sit in one of two buffers, selected by the parity of the gather sequence number)
__global__ void k_dist_assemble_bK(const double2 *__restrict__ recv0, const double2 *__restrict__ recv1,
                                   const unsigned long long *seq, double2 *__restrict__ b, int npr, int rows,
                                   int world, const int *done) {
  if (done && *done) return;
  const double2 *__restrict__ recv = (seq != nullptr && (*seq & 1ull)) ? recv1 : recv0;
  const int64_t nn = (int64_t)npr * ((int64_t)rows * world + 1);
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nn) return;
  int J, I;
  rc_of(n, npr, J, I);
  const int g = J / rows, j = J - g * rows;
  const int64_t per = (int64_t)(rows + 1) * npr;
  double2 acc = make_double2(0.0, 0.0);
  if (g < world) acc = recv[(int64_t)g * per + (int64_t)j * npr + I];
  if (j == 0 && g > 0) {
    const double2 t = recv[(int64_t)(g - 1) * per + (int64_t)rows * npr + I];
    acc.x += t.x; acc.y += t.y;
  }
  b[n] = acc;
}

LevelView dist_view(const topo_handle *h, int l) {
  LevelView v = view_of(h, h->levels[l]);
  if (h->dist.on && l < h->dist.K) { v.wf = h->dist.w_first; v.wl = h->dist.w_last; }
  return v;
}

}  // namespace

// Restricts the fine residual, runs levels 1..L-1 and adds the prolongated level-1 correction into the
// fine-grid iterate x^ (constrained DOFs stay zero).  No host-side state changes (pointers are never
// swapped), so the launch sequence can be captured in a CUDA graph and replayed.
// Result of a multi-block level lives in its x2 (nu_post odd) or x (nu_post even); tail levels leave it in x.
int topo_gmg_coarse(topo_handle *h, const double2 *res_fine, double2 *xhat_fine, const double2 **x1_out) {
  cudaStream_t st = h->stream;
  const double om = h->omega;
  const int L = h->n_levels;
  if (L < 2) return TOPO_OK;
  if (h->dist.on) {
    topo_set_error("row-slab handle: the V-cycle is driven segment by segment (topo_dist_*)");
    return TOPO_ESTATE;
  }
  const int *done = &h->sc->done;
  const int tail = tail_start(h);
  auto result_of = [&](int l) { return level_result(h, l, tail); };
  // tile kernels (restriction inside the down kernel, prolongation inside the up kernel) where a level does exactly one
  // sweep each way; the generic kernels below remain for TOPO_NU_COARSE > 1 and for the row-slab drivers
  auto tiled = [&](int l) { return h->knobs.lvl_tiled && nu_pre_of(h, l) == 1 && nu_post_of(h, l) == 1; };
  if (tail > 1 && !tiled(1)) {
    GmgLevel &C = h->levels[1];
    KL(h, "k_restrict", CUDA_TRY(topo_launch(h, k_restrict2<float2>, lvl_grid(C.nx, C.ny), dim3(LB_X, LB_Y), 0, reinterpret_cast<const float2 *>(res_fine), C.b, h->nx, h->ny, C.nx, C.ny, done)));
  }
  for (int l = 1; l < tail; ++l) {
    GmgLevel &F = h->levels[l];
    LevelView V = view_of(h, F);
    if (tiled(l)) {
      // (a level after a non-tiled one finds its b already restricted -- only with mixed sweep counts; recomputing it is harmless)
      if (l == 1) {
        KL(h, lvl_name("down", l), CUDA_TRY(topo_launch(h, V.sym ? k_lvl_down_t<true, float2> : k_lvl_down_t<false, float2>, tile_grid(F.nx, F.ny), dim3(TLX, TLY), 0, V, reinterpret_cast<const float2 *>(res_fine), h->nx, h->ny, om, done)));
      } else {
        const GmgLevel &P = h->levels[l - 1];
        KL(h, lvl_name("down", l), CUDA_TRY(topo_launch(h, V.sym ? k_lvl_down_t<true, double2> : k_lvl_down_t<false, double2>, tile_grid(F.nx, F.ny), dim3(TLX, TLY), 0, V, (const double2 *)P.r, P.nx, P.ny, om, done)));
      }
      if (l + 1 < tail && !tiled(l + 1)) {
        GmgLevel &C = h->levels[l + 1];
        KL(h, "k_restrict", CUDA_TRY(topo_launch(h, k_restrict2<double2>, lvl_grid(C.nx, C.ny), dim3(LB_X, LB_Y), 0, F.r, C.b, F.nx, F.ny, C.nx, C.ny, done)));
      }
      continue;
    }
    KL(h, lvl_name("down", l), CUDA_TRY(topo_launch(h, V.sym ? k_lvl_down2<true> : k_lvl_down2<false>, lvl_grid(F.nx, F.ny), dim3(LB_X, LB_Y), 0, V, om, done)));
    for (int s = 1; s < nu_pre_of(h, l); ++s) {
      KL(h, "k_lvl_smooth", CUDA_TRY(topo_launch(h, V.sym ? k_lvl_smooth2<true> : k_lvl_smooth2<false>, lvl_grid(F.nx, F.ny), dim3(LB_X, LB_Y), 0, V, om, done)));
      CUDA_TRY(cudaMemcpyAsync(F.x, F.x2, sizeof(double2) * (size_t)F.nn, cudaMemcpyDeviceToDevice, st));
      KL(h, "k_lvl_resid", CUDA_TRY(topo_launch(h, V.sym ? k_lvl_resid<true> : k_lvl_resid<false>, dim3(blocks_for(F.nn)), dim3(128), 0, V, done)));
    }
    if (l + 1 < tail && !tiled(l + 1)) {
      GmgLevel &C = h->levels[l + 1];
      KL(h, "k_restrict", CUDA_TRY(topo_launch(h, k_restrict2<double2>, lvl_grid(C.nx, C.ny), dim3(LB_X, LB_Y), 0, F.r, C.b, F.nx, F.ny, C.nx, C.ny, done)));
    }
  }
  if (tail == 1) TOPO_TRY(launch_tail(h, tail, res_fine, h->nx, h->ny, 1));
  else TOPO_TRY(launch_tail(h, tail, h->levels[tail - 1].r, h->levels[tail - 1].nx, h->levels[tail - 1].ny));
  for (int l = tail - 1; l >= 1; --l) {
    GmgLevel &F = h->levels[l], &C = h->levels[l + 1];
    LevelView V = view_of(h, F);
    const int unfused_min = h->knobs.unfused_min;
    if (tiled(l)) {
      KL(h, lvl_name("up", l), CUDA_TRY(topo_launch(h, V.sym ? k_lvl_up_t<true, false> : k_lvl_up_t<false, false>, tile_grid(F.nx, F.ny), dim3(TLX, TLY), 0, V, result_of(l + 1), C.nx, om, done, RowXchg{})));
      continue;
    }
    if (F.nn > unfused_min) {
      // big level: the stencil stream dominates, a separate cheap prolongation pass beats recomputing
      // P x_c for all 9 neighbours inside the smoothing kernel
      KL(h, "k_lvl_prolong", CUDA_TRY(topo_launch(h, k_lvl_prolong2, lvl_grid(F.nx, F.ny), dim3(LB_X, LB_Y), 0, V, result_of(l + 1), C.nx, done)));
      KL(h, lvl_name("up", l), CUDA_TRY(topo_launch(h, V.sym ? k_lvl_smooth2<true> : k_lvl_smooth2<false>, lvl_grid(F.nx, F.ny), dim3(LB_X, LB_Y), 0, V, om, done)));
    } else {
      KL(h, lvl_name("up", l), CUDA_TRY(topo_launch(h, V.sym ? k_lvl_up2<true> : k_lvl_up2<false>, lvl_grid(F.nx, F.ny), dim3(LB_X, LB_Y), 0, V, result_of(l + 1), C.nx, om, done)));
    }
    for (int s = 1; s < nu_post_of(h, l); ++s) {     // extra sweeps alternate x2 -> x -> x2 ...
      LevelView W = V;
      if (s & 1) { W.x = F.x2; W.x2 = F.x; }
      KL(h, "k_lvl_smooth", CUDA_TRY(topo_launch(h, W.sym ? k_lvl_smooth2<true> : k_lvl_smooth2<false>, lvl_grid(F.nx, F.ny), dim3(LB_X, LB_Y), 0, W, om, done)));
    }
  }
  if (x1_out != nullptr) {
    *x1_out = result_of(1);          // the caller's smoothing kernel prolongates on the fly
  } else {
    GmgLevel &C = h->levels[1];
    KL(h, "k_prolong_add_fine", CUDA_TRY(topo_launch(h, k_prolong_add_fine2, lvl_grid(h->nx, h->ny), dim3(LB_X, LB_Y), 0, result_of(1), reinterpret_cast<float2 *>(xhat_fine), h->fixed, h->nx, h->ny, C.nx, done)));
  }
  CUDA_TRY(cudaGetLastError());
  return TOPO_OK;
}

// ---- row-slab mode: the V-cycle in segments; the host completes interface rows between them -----------------
// (levels 1..K-1 are local to the slab, one sweep before/after; levels >= K are global and run in the tail kernel
// on every rank redundantly)

// b of the next level from a distributed residual: local level l+1 < K -> its b (distributed, the host halo-sums
// it); l+1 == K -> the send buffer of the gather
static int dist_restrict_from(topo_handle *h, int l, const double2 *r) {
  const DistState &D = h->dist;
  const int nxf = h->lvl_nx[l], nyf = (l == 0) ? h->ny : h->levels[l].ny;
  const int *done = &h->sc->done;
  // the fine-grid residual (l == 0) is a float2 array, level residuals are double2
  const float2 *rf32 = reinterpret_cast<const float2 *>(r);
  double2 *dst = (l + 1 < D.K) ? h->levels[l + 1].b : D.bK_send;
  const int nxc = h->lvl_nx[l + 1], nyc = nyf / 2;
  const int nb = blocks_for((int64_t)(nxc + 1) * (nyc + 1));
  const RowXchg x = topo_dist_row_xchg(h);
  if (x.c != nullptr && l + 1 < D.K) {                     // the interface rows of b_{l+1} are summed inside the kernel
    if (l == 0) KL(h, "k_restrict_x", k_restrict_x<<<lvl_grid(nxc, nyc), dim3(LB_X, LB_Y), 0, h->stream>>>(rf32, dst, nxf, nyf, nxc, nyc, done, x));
    else KL(h, "k_restrict_x", k_restrict_x<<<lvl_grid(nxc, nyc), dim3(LB_X, LB_Y), 0, h->stream>>>(r, dst, nxf, nyf, nxc, nyc, done, x));
    CUDA_TRY(cudaGetLastError());
    return TOPO_OK;
  }
  if (l == 0) KL(h, "k_restrict", k_restrict2<<<lvl_grid(nxc, nyc), dim3(LB_X, LB_Y), 0, h->stream>>>(rf32, dst, nxf, nyf, nxc, nyc, done));
  else KL(h, "k_restrict", k_restrict2<<<lvl_grid(nxc, nyc), dim3(LB_X, LB_Y), 0, h->stream>>>(r, dst, nxf, nyf, nxc, nyc, done));
  CUDA_TRY(cudaGetLastError());
  return TOPO_OK;
}

int topo_dist_restrict_fine(topo_handle *h, const double2 *res_fine) { return dist_restrict_from(h, 0, res_fine); }

// requires b_l accumulated: x = w D^-1 b ; r = W b - A_loc x (distributed) ; restrict
int topo_dist_level_down(topo_handle *h, int l) {
  if (!h->dist.on || l < 1 || l >= h->dist.K) return TOPO_EINVAL;
  const DistState &D = h->dist;
  if (D.mid_from > 0 && l >= D.mid_from) return l == D.mid_from ? topo_dist_mid_down(h) : TOPO_OK;
  GmgLevel &F = h->levels[l];
  const LevelView V = dist_view(h, l);
  KL(h, lvl_name("down", l), CUDA_TRY(topo_launch(h, V.sym ? k_lvl_down2<true> : k_lvl_down2<false>, lvl_grid(F.nx, F.ny), dim3(LB_X, LB_Y), 0, V,
                                                  h->omega, &h->sc->done)));
  if (D.mid_from > 0 && l + 1 == D.mid_from) return TOPO_OK;        // (the persistent kernel restricts this residual itself)
  return dist_restrict_from(h, l, F.r);
}

int topo_dist_tail(topo_handle *h, int p2p) {
  const DistState &D = h->dist;
  if (!D.on) return TOPO_EINVAL;
  GmgLevel &G = h->levels[D.K];
  const int rows = (h->ny >> D.K);
  if (p2p) {
    const double2 *r0 = reinterpret_cast<const double2 *>(D.p2p_region + D.p2p_off_gather);
    const double2 *r1 = reinterpret_cast<const double2 *>(D.p2p_region + D.p2p_off_gather + D.p2p_gather_stride);
    KL(h, "k_dist_assemble_bK", k_dist_assemble_bK<<<blocks_for(G.nn), 128, 0, h->stream>>>(r0, r1, &D.p2p_cnt->gather_wait, G.b, G.nx + 1,
                                                                                            rows, D.world, &h->sc->done));
  } else {
    KL(h, "k_dist_assemble_bK", k_dist_assemble_bK<<<blocks_for(G.nn), 128, 0, h->stream>>>(D.bK_recv, nullptr, nullptr, G.b, G.nx + 1, rows,
                                                                                            D.world, &h->sc->done));
  }
  return launch_tail(h, D.K, nullptr, 0, 0);
}

// levels >= K when the global right-hand side is already in place (topo_dist_p2p_gather_tail)
int topo_dist_tail_after_assemble(topo_handle *h) {
  if (!h->dist.on) return TOPO_EINVAL;
  return launch_tail(h, h->dist.K, nullptr, 0, 0);
}

// this slab's rows of the level-(l+1) result (global levels are indexed with the slab's row offset)
static const double2 *dist_coarse_result(const topo_handle *h, int l_coarse) {
  const DistState &D = h->dist;
  const double2 *x = level_result(h, l_coarse, D.K);
  if (l_coarse == D.K) x += (size_t)D.rank * (size_t)(h->ny >> D.K) * (size_t)(h->levels[D.K].nx + 1);
  return x;
}

// x += P x_{l+1} ; x2 = x + w D^-1 (W b - A_loc x): complete on interior rows, partial on interface rows
// (the host adds the neighbour's x2 row and subtracts x there: topo_dist_xchg_finish)
int topo_dist_level_up(topo_handle *h, int l) {
  if (!h->dist.on || l < 1 || l >= h->dist.K) return TOPO_EINVAL;
  if (h->dist.mid_from > 0 && l >= h->dist.mid_from) return l == h->dist.K - 1 ? topo_dist_mid_up(h) : TOPO_OK;
  GmgLevel &F = h->levels[l];
  const LevelView V = dist_view(h, l);
  const int *done = &h->sc->done;
  if (h->knobs.lvl_tiled) {
    // (x = w D^-1 b is recomputed from the accumulated b; the interface rows of x~ are stored for the exchange)
    KL(h, lvl_name("up", l), CUDA_TRY(topo_launch(h, V.sym ? k_lvl_up_t<true, true> : k_lvl_up_t<false, true>, tile_grid(F.nx, F.ny), dim3(TLX, TLY), 0, V,
                                                  dist_coarse_result(h, l + 1), h->lvl_nx[l + 1], h->omega, done, topo_dist_row_xchg(h))));
    return TOPO_OK;
  }
  KL(h, "k_lvl_prolong", k_lvl_prolong2<<<lvl_grid(F.nx, F.ny), dim3(LB_X, LB_Y), 0, h->stream>>>(V, dist_coarse_result(h, l + 1), h->lvl_nx[l + 1], done));
  KL(h, lvl_name("up", l), CUDA_TRY(topo_launch(h, V.sym ? k_lvl_smooth2<true> : k_lvl_smooth2<false>, lvl_grid(F.nx, F.ny), dim3(LB_X, LB_Y), 0, V,
                                                h->omega, done)));
  CUDA_TRY(cudaGetLastError());
  return TOPO_OK;
}

int topo_dist_prolong_fine(topo_handle *h, double2 *xhat_fine) {
  if (!h->dist.on) return TOPO_EINVAL;
  KL(h, "k_prolong_add_fine", k_prolong_add_fine2<<<lvl_grid(h->nx, h->ny), dim3(LB_X, LB_Y), 0, h->stream>>>(dist_coarse_result(h, 1), reinterpret_cast<float2 *>(xhat_fine), h->fixed, h->nx,
                                                                                           h->ny, h->lvl_nx[1], &h->sc->done));
  CUDA_TRY(cudaGetLastError());
  return TOPO_OK;
}

// ---- row-slab mode: the small slab-local levels in one persistent kernel per direction (k_slab_mid_down / _up) ------------
// plan: mid_from = first level l >= 2 (l < K) whose tiles AND nodes fit one wave of resident blocks -- every block then
// owns at most one tile / every thread at most one node per phase, so a thread waiting for a neighbour's row never holds
// back work the other neighbour waits for
int topo_dist_mid_plan(topo_handle *h, int enable) {
  DistState &D = h->dist;
  D.mid_from = 0;
  if (!enable || !D.on || !D.p2p_on || !D.p2p_inkernel || !(h->knobs.lvl_tiled & 1) || D.K < 3) return TOPO_OK;
  CUDA_TRY(cudaSetDevice(h->device));
  int sms = 0, occ_d = 0, occ_u = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_d, k_slab_mid_down, TLX * TLY, 0));
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_u, k_slab_mid_up, TLX * TLY, 0));
  const int cap = sms * std::max(1, std::min(std::min(occ_d, occ_u), 3));
  for (int l = 2; l < D.K; ++l) {
    const GmgLevel &L = h->levels[l];
    const int tiles = div_up(L.nx + 1, TLX) * div_up(L.ny + 1, TLY);
    if (tiles <= cap && (int64_t)L.nn <= (int64_t)cap * TLX * TLY && D.K - l <= 8) {
      D.mid_from = l;
      D.mid_grid = std::max(tiles, std::min(cap, div_up(L.nn, TLX * TLY)));
      break;
    }
  }
  return TOPO_OK;
}
static void mid_args(topo_handle *h, MidArgs &a) {
  const DistState &D = h->dist;
  a = MidArgs{};
  a.n = D.K - D.mid_from;
  for (int l = D.mid_from; l < D.K; ++l) a.lv[l - D.mid_from] = dist_view(h, l);
  a.omega = h->omega;
  a.done = &h->sc->done;
  a.x = topo_dist_row_xchg(h);
  a.bar = D.mid_bar;
}
int topo_dist_mid_down(topo_handle *h) {
  const DistState &D = h->dist;
  if (!D.on || D.mid_from < 2) return TOPO_EINVAL;
  MidArgs a;
  mid_args(h, a);
  const GmgLevel &P = h->levels[D.mid_from - 1];
  a.r_parent = P.r; a.pnx = P.nx; a.pny = P.ny;
  a.nxK = h->lvl_nx[D.K]; a.nyK = h->ny >> D.K;
  a.peers = topo_dist_peer_table(h);
  a.my_rank = D.rank; a.world = D.world;
  a.region = D.p2p_region; a.off_gather = D.p2p_off_ll_gather; a.gstride = 2 * D.p2p_gather_stride;
  a.bK = h->levels[D.K].b;
  KL(h, "k_slab_mid_down", k_slab_mid_down<<<D.mid_grid, dim3(TLX, TLY), 0, h->stream>>>(a));
  CUDA_TRY(cudaGetLastError());
  return TOPO_OK;
}
int topo_dist_mid_up(topo_handle *h) {
  const DistState &D = h->dist;
  if (!D.on || D.mid_from < 2) return TOPO_EINVAL;
  MidArgs a;
  mid_args(h, a);
  a.r_parent = dist_coarse_result(h, D.K);
  a.pnx = h->lvl_nx[D.K];
  KL(h, "k_slab_mid_up", k_slab_mid_up<<<D.mid_grid, dim3(TLX, TLY), 0, h->stream>>>(a));
  CUDA_TRY(cudaGetLastError());
  return TOPO_OK;
}
