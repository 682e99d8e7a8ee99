// Stage 1: matrix-free application of K(scale) on the structured quad4 mesh (fine grid).
//
// Replaces reference solver.py:364-410 (sparse_assem) + the CSR products inside spsolve and
// stiff_mat.dot (optimize.py:437,455): K is never formed.
//
// Kernel shape (node-centred marching, scatter-free):
//   * a warp owns 32 consecutive node columns and marches over a chunk of node rows;
//   * per row every lane loads ONE double2 (its node's (ux,uy): 512 B per warp, fully coalesced) and
//     ONE element scale; the left/right neighbours come from warp shuffles, the two strip-edge
//     neighbours from one predicated halo load (L1/L2 hit: the neighbouring strip reads that line);
//   * a 3-row register window (9 double2) and a 2-row scale window (4 doubles) give every u value
//     loaded from HBM exactly once per strip; the 8x8 element matrix is a kernel parameter, so its
//     entries are constant-bank operands of the DFMAs (no registers, no shared memory);
//   * each lane gathers the 4 element contributions of ITS node, so no atomics and no scatter.
// Algorithmic traffic: 16 B (u) + 16 B (y) per node + 8 B per element + 1 B/node of BC flags.
#include "topo_internal.cuh"

namespace {

enum { MV_APPLY = 0, MV_RESID = 1, MV_APPLY_DOT = 2, MV_SMOOTH = 3 };

struct MvArgs {
  const double2 *u;      // input vector
  const double *scale;   // element scale
  const uint8_t *fixed;  // per node: bit0 = ux constrained, bit1 = uy constrained
  double2 *y;            // output
  const double2 *b;      // MV_RESID / MV_SMOOTH: right-hand side
  const double2 *dinv;   // MV_SMOOTH: inverse diagonal
  double omega;          // MV_SMOOTH
  double *partials;      // MV_APPLY_DOT
  unsigned int *ticket;
  double *dot_out;       // MV_APPLY_DOT: receives sum u.y
  const int *done;       // optional early-exit flag (PCG converged)
  int nx, ny, rows_per_chunk;
};

__device__ __forceinline__ void row_terms(const KeParam &ke, int a, const double2 (&q)[4], double &tx,
                                          double &ty) {
  // (rows 2a, 2a+1 of ke) . [q0.x q0.y q1.x q1.y q2.x q2.y q3.x q3.y]
  const double *r0 = &ke.k[(2 * a) * 8];
  const double *r1 = &ke.k[(2 * a + 1) * 8];
  tx = r0[0] * q[0].x;
  ty = r1[0] * q[0].x;
  tx = fma(r0[1], q[0].y, tx);
  ty = fma(r1[1], q[0].y, ty);
#pragma unroll
  for (int b = 1; b < 4; ++b) {
    tx = fma(r0[2 * b], q[b].x, tx);
    ty = fma(r1[2 * b], q[b].x, ty);
    tx = fma(r0[2 * b + 1], q[b].y, tx);
    ty = fma(r1[2 * b + 1], q[b].y, ty);
  }
}

template <int MODE>
__global__ void __launch_bounds__(128) k_fine_apply(const __grid_constant__ KeParam ke, const MvArgs a) {
  if (a.done != nullptr && *a.done) return;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int npr = a.nx + 1;                              // nodes per row
  const int c0 = (blockIdx.x * 4 + warp) * 32;           // first node column of this warp's strip
  const int c = c0 + lane;
  const int r0 = blockIdx.y * a.rows_per_chunk;
  const int r1 = min(r0 + a.rows_per_chunk, a.ny + 1);
  const bool strip_active = c0 < npr;                    // warp-uniform
  const bool col_ok = c < npr;
  const bool ecol_ok = c < a.nx;                         // element column c exists
  const double2 zero2 = make_double2(0.0, 0.0);
  double dot = 0.0;

  if (strip_active) {
    // halo column handled by lane 0 (left, c0-1) and lane 31 (right, c0+32)
    const int hc = (lane == 0) ? c0 - 1 : c0 + 32;
    const bool h_ok = (lane == 0) ? (c0 > 0) : (lane == 31 && hc < npr);
    const bool he_ok = (lane == 0) && (c0 > 0);          // element column c0-1 for lane 0's left element

    // window: U[k][j]: k = 0,1,2 -> rows r-1,r,r+1 ; j = 0,1,2 -> cols c-1,c,c+1
    double2 U[3][3];
    double El[2], Er[2];                                 // El: element col c-1, Er: element col c; [0]=row r-1,[1]=row r

    auto load_urow = [&](int row, double2 &l, double2 &m, double2 &rgt) {
      double2 own = zero2, halo = zero2;
      if (row >= 0 && row <= a.ny) {
        const double2 *p = a.u + (size_t)row * npr;
        if (col_ok) own = __ldg(p + c);
        if (h_ok) halo = __ldg(p + hc);
      }
      l = shfl_up2(own, 1);
      rgt = shfl_down2(own, 1);
      if (lane == 0) l = halo;
      if (lane == 31) rgt = halo;
      m = own;
    };
    auto load_erow = [&](int row, double &l, double &rgt) {
      double own = 0.0, halo = 0.0;
      if (row >= 0 && row < a.ny) {
        const double *p = a.scale + (size_t)row * a.nx;
        if (ecol_ok) own = __ldg(p + c);
        if (he_ok) halo = __ldg(p + c0 - 1);
      }
      l = __shfl_up_sync(0xffffffffu, own, 1);
      if (lane == 0) l = halo;
      rgt = own;
    };

    load_urow(r0 - 1, U[0][0], U[0][1], U[0][2]);
    load_urow(r0, U[1][0], U[1][1], U[1][2]);
    load_erow(r0 - 1, El[0], Er[0]);

#pragma unroll 2
    for (int r = r0; r < r1; ++r) {
      load_urow(r + 1, U[2][0], U[2][1], U[2][2]);
      load_erow(r, El[1], Er[1]);

      // element NE = (r, c): this node is its local node 0
      double yx, yy, tx, ty;
      {
        const double2 q[4] = {U[1][1], U[1][2], U[2][2], U[2][1]};
        row_terms(ke, 0, q, tx, ty);
        yx = Er[1] * tx;
        yy = Er[1] * ty;
      }
      {  // NW = (r, c-1): local node 1
        const double2 q[4] = {U[1][0], U[1][1], U[2][1], U[2][0]};
        row_terms(ke, 1, q, tx, ty);
        yx = fma(El[1], tx, yx);
        yy = fma(El[1], ty, yy);
      }
      {  // SW = (r-1, c-1): local node 2
        const double2 q[4] = {U[0][0], U[0][1], U[1][1], U[1][0]};
        row_terms(ke, 2, q, tx, ty);
        yx = fma(El[0], tx, yx);
        yy = fma(El[0], ty, yy);
      }
      {  // SE = (r-1, c): local node 3
        const double2 q[4] = {U[0][1], U[0][2], U[1][2], U[1][1]};
        row_terms(ke, 3, q, tx, ty);
        yx = fma(Er[0], tx, yx);
        yy = fma(Er[0], ty, yy);
      }

      if (col_ok) {
        const size_t n = (size_t)r * npr + c;
        const unsigned fx = a.fixed[n];
        if (MODE == MV_APPLY || MODE == MV_APPLY_DOT) {
          if (fx & 1u) yx = 0.0;
          if (fx & 2u) yy = 0.0;
          a.y[n] = make_double2(yx, yy);
          if (MODE == MV_APPLY_DOT) dot = fma(U[1][1].x, yx, fma(U[1][1].y, yy, dot));
        } else if (MODE == MV_RESID) {
          const double2 bv = __ldg(a.b + n);
          double rx = bv.x - yx, ry = bv.y - yy;
          if (fx & 1u) rx = 0.0;
          if (fx & 2u) ry = 0.0;
          a.y[n] = make_double2(rx, ry);
        } else {  // MV_SMOOTH: y = u + omega * dinv * (b - A u)   (dinv is 0 at constrained DOFs)
          const double2 bv = __ldg(a.b + n);
          const double2 dv = __ldg(a.dinv + n);
          a.y[n] = make_double2(fma(a.omega * dv.x, bv.x - yx, U[1][1].x),
                                fma(a.omega * dv.y, bv.y - yy, U[1][1].y));
        }
      }
      // rotate windows
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        U[0][j] = U[1][j];
        U[1][j] = U[2][j];
      }
      El[0] = El[1];
      Er[0] = Er[1];
    }
  }
  if (MODE == MV_APPLY_DOT) {
    double v[1] = {dot};
    grid_reduce<1>(v, a.partials, a.ticket, a.dot_out);
  }
}

// diag(K) per node from the 4 adjacent element scales; stores 1/diag (0 at constrained DOFs) and the
// max diagonal over free DOFs (= K.max() of the reference's equilibrium guard, optimize.py:455: the
// largest entry of an SPD matrix sits on its diagonal).
__global__ void k_build_diag(const __grid_constant__ KeParam ke, const double *__restrict__ scale,
                             const uint8_t *__restrict__ fixed, double2 *__restrict__ dinv, int nx, int ny,
                             double *partials, unsigned int *ticket, double *kmax_out) {
  const int64_t nn = (int64_t)(nx + 1) * (ny + 1);
  double mx = 0.0;
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < nn; n += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(n / (nx + 1)), c = (int)(n - (int64_t)r * (nx + 1));
    const double ne = (r < ny && c < nx) ? scale[(size_t)r * nx + c] : 0.0;
    const double nw = (r < ny && c > 0) ? scale[(size_t)r * nx + c - 1] : 0.0;
    const double sw = (r > 0 && c > 0) ? scale[(size_t)(r - 1) * nx + c - 1] : 0.0;
    const double se = (r > 0 && c < nx) ? scale[(size_t)(r - 1) * nx + c] : 0.0;
    double dx = ne * ke.k[0 * 8 + 0] + nw * ke.k[2 * 8 + 2] + sw * ke.k[4 * 8 + 4] + se * ke.k[6 * 8 + 6];
    double dy = ne * ke.k[1 * 8 + 1] + nw * ke.k[3 * 8 + 3] + sw * ke.k[5 * 8 + 5] + se * ke.k[7 * 8 + 7];
    const unsigned fx = fixed[n];
    double ix = 0.0, iy = 0.0;
    if (!(fx & 1u)) {
      mx = fmax(mx, dx);
      ix = dx > 0.0 ? 1.0 / dx : 0.0;
    }
    if (!(fx & 2u)) {
      mx = fmax(mx, dy);
      iy = dy > 0.0 ? 1.0 / dy : 0.0;
    }
    dinv[n] = make_double2(ix, iy);
  }
  double v[1] = {mx};
  grid_reduce<1, true>(v, partials, ticket, kmax_out);
}

template <int MODE>
int launch_apply(topo_handle *h, MvArgs &a) {
  a.scale = h->scale;
  a.fixed = h->fixed;
  a.nx = h->nx;
  a.ny = h->ny;
  a.rows_per_chunk = h->matvec_rows_per_chunk;
  a.partials = h->partials;
  a.ticket = h->ticket;
  dim3 grid(div_up(h->nx + 1, 128), div_up(h->ny + 1, a.rows_per_chunk));
  if ((int64_t)grid.x * grid.y > TOPO_MAX_PARTIAL_BLOCKS) {
    topo_set_error("matvec grid too large for the reduction buffer");
    return TOPO_EINVAL;
  }
  const bool prof = h->prof_on && (size_t)(2 * h->prof_n + 1) < h->prof_ev.size();
  if (prof) cudaEventRecord(h->prof_ev[2 * h->prof_n], h->stream);
  k_fine_apply<MODE><<<grid, 128, 0, h->stream>>>(h->ke, a);
  if (prof) cudaEventRecord(h->prof_ev[2 * h->prof_n++ + 1], h->stream);
  h->timing.kernel_launches++;
  CUDA_TRY(cudaGetLastError());
  return TOPO_OK;
}

}  // namespace

int topo_fine_matvec(topo_handle *h, const double2 *u, double2 *y) {
  MvArgs a = {};
  a.u = u;
  a.y = y;
  return launch_apply<MV_APPLY>(h, a);
}

int topo_fine_matvec_dot(topo_handle *h, const double2 *u, double2 *y, double *dot_out, const int *done) {
  MvArgs a = {};
  a.u = u;
  a.y = y;
  a.dot_out = dot_out;
  a.done = done;
  return launch_apply<MV_APPLY_DOT>(h, a);
}

int topo_fine_residual(topo_handle *h, const double2 *u, const double2 *b, double2 *r) {
  MvArgs a = {};
  a.u = u;
  a.b = b;
  a.y = r;
  return launch_apply<MV_RESID>(h, a);
}

int topo_fine_smooth(topo_handle *h, const double2 *x, const double2 *b, double2 *x_out, double omega,
                     const int *done) {
  MvArgs a = {};
  a.u = x;
  a.b = b;
  a.y = x_out;
  a.dinv = h->dinv;
  a.omega = omega;
  a.done = done;
  return launch_apply<MV_SMOOTH>(h, a);
}

int topo_build_diag(topo_handle *h) {
  const int blocks = (int)std::min<int64_t>(div_up(h->nn, 256), 1184);
  k_build_diag<<<blocks, 256, 0, h->stream>>>(h->ke, h->scale, h->fixed, h->dinv, h->nx, h->ny, h->partials,
                                              h->ticket, h->kmax);
  h->timing.kernel_launches++;
  CUDA_TRY(cudaGetLastError());
  h->diag_valid = true;
  return TOPO_OK;
}
