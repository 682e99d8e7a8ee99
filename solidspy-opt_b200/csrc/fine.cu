// Stage 1 + the fine-grid half of stage 2: matrix-free application of K(scale) on the structured
// quad4 mesh, fused with the PCG / multigrid vector work that surrounds it.
//
// Replaces reference solver.py:364-410 (sparse_assem) and the CSR products inside spsolve /
// stiff_mat.dot (optimize.py:437,455): K is never formed.
//
// Kernel shape ("staged marching", warp-private):
//   * a WARP owns 32 node columns (+1 halo column each side) and marches over a chunk of node rows; the
//     four warps of a CTA are independent (no block barrier anywhere in the row loop);
//   * every raw input vector (1..5 per kernel) is streamed row by row into the warp's shared-memory ring
//     with cp.async (LDGSTS) several rows ahead of use, so the bytes in flight per SM are set by the ring
//     depth, not by registers or occupancy; element scales ride in the same ring;
//   * stage B ("derive"): each lane turns the raw values of ITS node of the newest row into the vector the
//     stencil is applied to (e.g. p = z + beta p), stores side outputs, and publishes the value in a
//     2-slot shared row (__syncwarp);
//   * stage C: each lane gathers the 4 element contributions of ITS node from a 3x3 register window
//     (new row: 3 LDS.128), so there are no atomics and no scatter; the element matrix is 16 constants
//     in registers (mirror symmetry of the rectangle);
//   * epilogue: fused axpy / smoothing / masking and warp+block+ordered-grid dot products.
// Ring slots and the register window are compile-time (row loop unrolled by the ring depth).
// Inputs that are also outputs (p, u, r) are written to a second buffer: neighbouring warps still read
// the old halo rows/columns.
// Algorithmic traffic of the plain product: 16 B (u) + 16 B (y) per node + 8 B per element.
#include "topo_internal.cuh"

namespace {

constexpr int BW = 128;        // threads per CTA = 4 warps = 4 independent strips of 32 columns
// NC = node columns per lane: a warp owns 32*NC columns (+1 halo column each side)
template <int NC> struct Geo {
  static constexpr int WC = 32 * NC;     // own columns per warp
  static constexpr int SW = WC + 2;      // staged node columns per warp
  static constexpr int EW = WC + 2;      // staged element columns per warp (c0-1 .. c0+WC-1, padded to even)
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, bool valid) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem, bool valid) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int sz = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

enum {
  FK_APPLY = 0,        // y = mask(A u)                                  raw: u
  FK_RESID = 1,        // y = mask(b - A u)                              raw: u, b
  FK_PCG_MATVEC = 2,   // p' = z + beta p ; Ap = A p' ; pAp              raw: z, p
  FK_PCG_MATVEC_J = 3, // p' = dinv*r + beta p ; Ap ; pAp  (Jacobi)      raw: r, p, dinv
  FK_UPDATE_RESID0 = 4,// u' = u + alpha p ; r' = r - alpha Ap ; rr ;    raw: p, Ap, u, r, dinv
                       // x^ = omega dinv r' ; res = r' - A x^
  FK_PROLONG_SMOOTH = 5,// x~ = x^ + P xc ; z = x~ + omega dinv (r - A x~) ; rz   raw: x^, r, dinv
  FK_SMOOTH = 6        // y = u + omega dinv (b - A u)                   raw: u, b, dinv
};

// ring depth per kind (S slots, S-2 rows in flight per warp); the row loop is unrolled by S and the
// register window cycles through 4 row images, so slots and window roles are compile-time
// F32MASK: bit k set = raw input k is a float2 array in global memory (the vectors that exist only inside the
// preconditioner -- x^, res, and the fp32 copy of D^-1 -- are stored in fp32: they carry 18 % of the fine-grid
// traffic of a PCG iteration and the multigrid correction is accurate to a few per cent anyway).  Such an input
// still owns a double2 slot per node in the ring; only its low 8 bytes are copied and read.
template <int K> struct FK;
template <> struct FK<FK_APPLY> { static constexpr int NIN = 1, S = 8, F32MASK = 0; static constexpr bool DERIVE = false; };
template <> struct FK<FK_RESID> { static constexpr int NIN = 2, S = 8, F32MASK = 0; static constexpr bool DERIVE = false; };
template <> struct FK<FK_SMOOTH> { static constexpr int NIN = 3, S = 4, F32MASK = 0; static constexpr bool DERIVE = false; };
template <> struct FK<FK_PCG_MATVEC> { static constexpr int NIN = 2, S = 8, F32MASK = 0; static constexpr bool DERIVE = true; };
template <> struct FK<FK_PCG_MATVEC_J> { static constexpr int NIN = 3, S = 4, F32MASK = 0; static constexpr bool DERIVE = true; };
template <> struct FK<FK_UPDATE_RESID0> { static constexpr int NIN = 5, S = 4, F32MASK = 1 << 4; static constexpr bool DERIVE = true; };
template <> struct FK<FK_PROLONG_SMOOTH> { static constexpr int NIN = 3, S = 4, F32MASK = (1 << 0) | (1 << 2); static constexpr bool DERIVE = true; };

// value of raw input k at a ring position (fp32 inputs sit in the low half of their slot)
template <int KIND, int K>
__device__ __forceinline__ double2 raw_in(const double2 *rw, int SW) {
  if ((FK<KIND>::F32MASK >> K) & 1) {
    const float2 f = *reinterpret_cast<const float2 *>(rw + K * SW);
    return make_double2((double)f.x, (double)f.y);
  }
  return rw[K * SW];
}

struct FineArgs {
  const double2 *in[5];
  double2 *out[4];
  const double *scale;
  const uint8_t *fixed;
  const double2 *coarse;     // FK_PROLONG_SMOOTH: level-1 correction (nullptr: already added)
  int cnx;                   // its cells per row
  double omega;
  double *partials;
  unsigned int *ticket;
  double *dot_out;           // FK_APPLY with want_dot
  SolverScalars *sc;
  int nx, ny, rows_per_chunk;
  int want_dot;              // FK_APPLY: reduce u.y into dot_out
  int check_done;            // early exit when sc->done
  int init;                  // FK_UPDATE_RESID0 / FK_PROLONG_SMOOTH: first pass (alpha = 0, beta = 0)
  // row-slab mode (DistState): weights of node rows 0 and ny that turn an accumulated vector into a distributed
  // one; dist != 0: dot products are left as per-rank partial sums in sc->aux[0..2] (pAp, rr, rz) for the host's
  // all-reduce, the scalar bookkeeping happens afterwards (k_dist_pcg_finish)
  double w_first, w_last;
  int dist;
};

// Rows 0 and 1 of the element matrix (the row block of local node 0).  For a rectangular element of an
// isotropic material the row blocks of the other three local nodes follow from it by the element's
// mirror symmetries (node permutation + sign flips, checked on the host in topo_create), so the 72
// DFMAs of a node use 16 constants that live in ordinary registers.  (On sm_100a DFMA cannot take a
// constant-bank operand: with all 64 entries as kernel-parameter constants ptxas spends two R2UR per
// DFMA and 128 registers on them.)
struct KeRow {
  double r0[8], r1[8];
};

// FLIP = false: t = R0 . q      (NE element, node is local 0; SW element with nodes permuted [2,3,0,1])
// FLIP = true : mirrored once   (NW element: x-mirror; SE element: y-mirror) -- same sign pattern
template <bool FLIP>
__device__ __forceinline__ void row_terms(const KeRow &k, const double2 &q0, const double2 &q1, const double2 &q2,
                                          const double2 &q3, double &tx, double &ty) {
  if (!FLIP) {
    tx = k.r0[0] * q0.x;            ty = k.r1[0] * q0.x;
    tx = fma(k.r0[1], q0.y, tx);    ty = fma(k.r1[1], q0.y, ty);
    tx = fma(k.r0[2], q1.x, tx);    ty = fma(k.r1[2], q1.x, ty);
    tx = fma(k.r0[3], q1.y, tx);    ty = fma(k.r1[3], q1.y, ty);
    tx = fma(k.r0[4], q2.x, tx);    ty = fma(k.r1[4], q2.x, ty);
    tx = fma(k.r0[5], q2.y, tx);    ty = fma(k.r1[5], q2.y, ty);
    tx = fma(k.r0[6], q3.x, tx);    ty = fma(k.r1[6], q3.x, ty);
    tx = fma(k.r0[7], q3.y, tx);    ty = fma(k.r1[7], q3.y, ty);
  } else {
    tx = k.r0[0] * q0.x;            ty = k.r1[1] * q0.y;
    tx = fma(-k.r0[1], q0.y, tx);   ty = fma(-k.r1[0], q0.x, ty);
    tx = fma(k.r0[2], q1.x, tx);    ty = fma(-k.r1[2], q1.x, ty);
    tx = fma(-k.r0[3], q1.y, tx);   ty = fma(k.r1[3], q1.y, ty);
    tx = fma(k.r0[4], q2.x, tx);    ty = fma(-k.r1[4], q2.x, ty);
    tx = fma(-k.r0[5], q2.y, tx);   ty = fma(k.r1[5], q2.y, ty);
    tx = fma(k.r0[6], q3.x, tx);    ty = fma(-k.r1[6], q3.x, ty);
    tx = fma(-k.r0[7], q3.y, tx);   ty = fma(k.r1[7], q3.y, ty);
  }
}

__device__ __forceinline__ void prolong_taps1(int i, int nf, int &I0, int &I1, double &w1) {
  if ((i & 1) == 0) { I0 = I1 = i >> 1; w1 = 0.0; }
  else if (i == nf) { I0 = I1 = (i + 1) >> 1; w1 = 0.0; }
  else { I0 = (i - 1) >> 1; I1 = (i + 1) >> 1; w1 = 0.5; }
}

// register image of one node row as seen by a thread: stencil input at columns c-1, c, c+1 and the
// scales of the two elements BELOW the row (element row q-1, columns c-1 and c)
template <int NC>
struct RowWin {
  double2 u[NC + 2];     // columns c-1 .. c+NC of this lane's NC own columns c .. c+NC-1
  double e[NC + 1];      // scales of element columns c-1 .. c+NC-1 of the element row below
};

template <int V> struct IC { static constexpr int value = V; };

template <int KIND, int NC>
__host__ __device__ constexpr int warp_smem_bytes() {
  return (int)(sizeof(double2) * (FK<KIND>::S * FK<KIND>::NIN * Geo<NC>::SW + 2 * Geo<NC>::SW) +
               sizeof(double) * FK<KIND>::S * Geo<NC>::EW + 16 * ((FK<KIND>::S * ((Geo<NC>::WC + 6) / 4) * 4 + 15) / 16));
}

template <int KIND, int NC>
__global__ void __launch_bounds__(BW, NC == 1 ? 4 : 3) k_fine(const __grid_constant__ KeRow kparam,
                                                              const __grid_constant__ FineArgs a) {
  constexpr int NIN = FK<KIND>::NIN, S = FK<KIND>::S, D = S - 2;
  constexpr int WC = Geo<NC>::WC, SW = Geo<NC>::SW, EW = Geo<NC>::EW;
  constexpr bool DERIVE = FK<KIND>::DERIVE;
  // BC flags are needed only by the kinds whose output is defined as masked and that have no D^-1 input; they ride
  // in the ring as aligned 4-byte words (a direct byte load prefetched one row ahead was the dominant stall)
  constexpr bool NEEDS_FIXED = (KIND == FK_APPLY || KIND == FK_RESID);
  constexpr int FWORDS = (WC + 3 + 3) / 4;               // aligned words covering the warp's WC flag bytes
  if (a.check_done && a.sc->done) return;
  const KeRow ke = kparam;                                // 16 constants -> (uniform) registers
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char *wbase = smem_raw + warp * warp_smem_bytes<KIND, NC>();
  double2 *raw = reinterpret_cast<double2 *>(wbase);                     // [S][NIN][SW]
  double2 *der = raw + S * NIN * SW;                                     // [2][SW] (DERIVE kinds)
  double *ering = reinterpret_cast<double *>(der + 2 * SW);              // [S][EW]
  unsigned *fring = reinterpret_cast<unsigned *>(ering + S * EW);        // [S][FWORDS] (NEEDS_FIXED kinds)

  const int npr = a.nx + 1;
  const int c0 = (blockIdx.x * 4 + warp) * WC;            // first own node column of this warp
  const int cl = c0 + NC * lane;                          // first own node column of this lane (staged NC*lane+1)
  const int sl = NC * lane + 1;
  const int r0 = blockIdx.y * a.rows_per_chunk;
  const int r1 = min(r0 + a.rows_per_chunk, a.ny + 1);
  const bool active = c0 < npr;                           // warp-uniform
  // halo duty: lane 0 -> staged col 0 (node col c0-1), lane 1 -> staged col SW-1 (node col c0+WC)
  const int hs = (lane == 0) ? 0 : SW - 1;
  const int hc = (lane == 0) ? c0 - 1 : c0 + WC;
  const bool has_halo = lane < 2;
  const bool hcol_ok = has_halo && hc >= 0 && hc < npr;
  const bool ehalo_ok = (lane == 0) && c0 > 0;
  // column offsets clamped into the row so that masked (zero-fill) copies still carry a valid address
  const int hoff = hcol_ok ? hc : 0, ehoff = ehalo_ok ? c0 - 1 : 0;

  double alpha = 0.0, beta = 0.0;
  if (KIND == FK_UPDATE_RESID0 && !a.init) alpha = a.sc->rz[0] / (a.dist ? a.sc->aux[0] : a.sc->pAp);
  auto wrow = [&](int row) -> double { return row == 0 ? a.w_first : (row == a.ny ? a.w_last : 1.0); };
  if (KIND == FK_PCG_MATVEC || KIND == FK_PCG_MATVEC_J) beta = a.sc->rz[1];     // rz[1] holds beta (pcg.cu)
  const double omega = a.omega;
  double dots[1] = {0.0};

  if (active) {
    // ---- copies of group G_q = {raw row q, element row q-1} into ring slot SL = (q - (r0-1)) mod S ----
    int q_issue = r0 - 1;
    auto issue = [&](auto slc) {
      constexpr int SL = decltype(slc)::value;
      const int q = q_issue++;
      const bool row_ok = q >= 0 && q <= a.ny;
      const size_t rowoff = (size_t)(row_ok ? q : 0) * npr;
#pragma unroll
      for (int k = 0; k < NIN; ++k) {
        double2 *dst = raw + (SL * NIN + k) * SW;
        if ((FK<KIND>::F32MASK >> k) & 1) {
          const float2 *src = reinterpret_cast<const float2 *>(a.in[k]) + rowoff;
#pragma unroll
          for (int j = 0; j < NC; ++j) {
            const bool ok = cl + j < npr;
            cp_async8(dst + sl + j, src + (ok ? cl + j : 0), row_ok && ok);
          }
          if (has_halo) cp_async8(dst + hs, src + hoff, row_ok && hcol_ok);
        } else {
          const double2 *src = a.in[k] + rowoff;
#pragma unroll
          for (int j = 0; j < NC; ++j) {
            const bool ok = cl + j < npr;
            cp_async16(dst + sl + j, src + (ok ? cl + j : 0), row_ok && ok);
          }
          if (has_halo) cp_async16(dst + hs, src + hoff, row_ok && hcol_ok);
        }
      }
      const bool erow_ok = q >= 1 && q <= a.ny;
      const double *esrc = a.scale + (size_t)(erow_ok ? q - 1 : 0) * a.nx;
      double *edst = ering + SL * EW;
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        const bool ok = cl + j < a.nx;
        cp_async8(edst + sl + j, esrc + (ok ? cl + j : 0), erow_ok && ok);
      }
      if (lane == 0) cp_async8(edst, esrc + ehoff, erow_ok && ehalo_ok);
      if (NEEDS_FIXED && lane < FWORDS) {
        const size_t a0 = ((size_t)(row_ok ? q : 0) * npr + c0) & ~(size_t)3;
        const unsigned sdst = (unsigned)__cvta_generic_to_shared(fring + SL * FWORDS + lane);
        const int sz = row_ok ? 4 : 0;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(sdst), "l"(a.fixed + a0 + 4 * lane), "r"(sz) : "memory");
      }
      cp_async_commit();
    };

    // ---- stage B (DERIVE kinds): stencil input of (row, column) from this lane's own raw copies ----
    auto derive_at = [&](const double2 *rw, int row, int ncol, bool own) -> double2 {
      const bool in_dom = row >= 0 && row <= a.ny && ncol >= 0 && ncol < npr;
      const size_t n = (size_t)(in_dom ? row : 0) * npr + (in_dom ? ncol : 0);
      const bool store = own && in_dom && row >= r0 && row < r1;
      if (KIND == FK_PCG_MATVEC) {
        const double2 z = rw[0], p = rw[SW];
        const double2 v = make_double2(fma(beta, p.x, z.x), fma(beta, p.y, z.y));
        if (store) a.out[0][n] = v;
        return v;
      } else if (KIND == FK_PCG_MATVEC_J) {
        const double2 r = rw[0], p = rw[SW], d = rw[2 * SW];
        const double2 v = make_double2(fma(beta, p.x, d.x * r.x), fma(beta, p.y, d.y * r.y));
        if (store) a.out[0][n] = v;
        return v;
      } else if (KIND == FK_UPDATE_RESID0) {
        const double2 p = rw[0], ap = rw[SW], u = rw[2 * SW], r = rw[3 * SW], d = raw_in<KIND, 4>(rw, SW);
        const double2 rn = make_double2(d.x != 0.0 ? fma(-alpha, ap.x, r.x) : 0.0, d.y != 0.0 ? fma(-alpha, ap.y, r.y) : 0.0);
        const double2 xh = make_double2(omega * d.x * rn.x, omega * d.y * rn.y);
        if (store) {
          a.out[0][n] = make_double2(fma(alpha, p.x, u.x), fma(alpha, p.y, u.y));
          a.out[1][n] = rn;
          reinterpret_cast<float2 *>(a.out[2])[n] = make_float2((float)xh.x, (float)xh.y);
          const double wr = wrow(row);
          dots[0] = fma(wr * rn.x, rn.x, fma(wr * rn.y, rn.y, dots[0]));
        }
        return xh;
      } else {  // FK_PROLONG_SMOOTH
        const double2 xh = raw_in<KIND, 0>(rw, SW), d = raw_in<KIND, 2>(rw, SW);
        double px = 0.0, py = 0.0;
        if (in_dom && a.coarse != nullptr) {
          int I0, I1, J0, J1;
          double wx1, wy1;
          prolong_taps1(ncol, a.nx, I0, I1, wx1);
          prolong_taps1(row, a.ny, J0, J1, wy1);
          const double wx0 = 1.0 - wx1, wy0 = 1.0 - wy1;
          const int nprc = a.cnx + 1;
          const double2 v00 = __ldg(a.coarse + (size_t)J0 * nprc + I0);
          px = wy0 * wx0 * v00.x;
          py = wy0 * wx0 * v00.y;
          if (wx1 != 0.0) {
            const double2 v = __ldg(a.coarse + (size_t)J0 * nprc + I1);
            px = fma(wy0 * wx1, v.x, px); py = fma(wy0 * wx1, v.y, py);
          }
          if (wy1 != 0.0) {
            const double2 v = __ldg(a.coarse + (size_t)J1 * nprc + I0);
            px = fma(wy1 * wx0, v.x, px); py = fma(wy1 * wx0, v.y, py);
            if (wx1 != 0.0) {
              const double2 w = __ldg(a.coarse + (size_t)J1 * nprc + I1);
              px = fma(wy1 * wx1, w.x, px); py = fma(wy1 * wx1, w.y, py);
            }
          }
        }
        return make_double2(d.x != 0.0 ? xh.x + px : 0.0, d.y != 0.0 ? xh.y + py : 0.0);
      }
    };
    auto stage_b = [&](auto slc, int row) {
      constexpr int SL = decltype(slc)::value;
      if (DERIVE) {
        double2 *dr = der + (SL & 1) * SW;
        const double2 *rw = raw + SL * NIN * SW;
#pragma unroll
        for (int j = 0; j < NC; ++j) dr[sl + j] = derive_at(rw + sl + j, row, cl + j, true);
        if (has_halo) dr[hs] = derive_at(rw + hs, row, hc, false);
      }
    };
    // register image of the node row held in ring slot SL (after the __syncwarp that publishes it)
    auto fill = [&](auto slc, RowWin<NC> &w) {
      constexpr int SL = decltype(slc)::value;
      const double2 *src = DERIVE ? (der + (SL & 1) * SW + sl - 1) : (raw + SL * NIN * SW + sl - 1);
#pragma unroll
      for (int i = 0; i < NC + 2; ++i) w.u[i] = src[i];
      const double *e = ering + SL * EW + sl - 1;
#pragma unroll
      for (int i = 0; i < NC + 1; ++i) w.e[i] = e[i];
    };

    // ---- prologue: S groups in flight (rows r0-1 .. r0+D), rows r0-1 and r0 resident ----
    issue(IC<0>()); issue(IC<1>()); issue(IC<2>()); issue(IC<3>());
    if (S == 8) { issue(IC<4 % S>()); issue(IC<5 % S>()); issue(IC<6 % S>()); issue(IC<7 % S>()); }
    cp_async_wait<D>();
    RowWin<NC> W0, W1, W2, W3;
    if (DERIVE) {
      stage_b(IC<0>(), r0 - 1);
      stage_b(IC<1>(), r0);
    }
    __syncwarp();
    fill(IC<0>(), W0);
    fill(IC<1>(), W1);
    size_t n_out = (size_t)r0 * npr + cl;

    // one node row r held in slot K+1: Wm = row r-1, Wc = row r, Wp receives row r+1 (slot K+2);
    // the slot of row r-1 (K) is refilled with row r-1+S
    auto step = [&](auto kc, int r, const RowWin<NC> &Wm, const RowWin<NC> &Wc, RowWin<NC> &Wp) {
      constexpr int K = decltype(kc)::value;
      constexpr int S_prev = K % S, S_cur = (K + 1) % S, S_next = (K + 2) % S;
      cp_async_wait<D - 1>();                              // group G_{r+1} has landed (this lane's part)
      stage_b(IC<S_next>(), r + 1);
      __syncwarp();                                        // row r+1 visible; every lane is done with row r-1's slot
      issue(IC<S_prev>());
      fill(IC<S_next>(), Wp);
      unsigned fx[NC];
#pragma unroll
      for (int j = 0; j < NC; ++j) fx[j] = 0u;
      if (NEEDS_FIXED) {
        const unsigned char *fb = reinterpret_cast<const unsigned char *>(fring + S_cur * FWORDS);
        const unsigned off = (unsigned)((n_out - (size_t)(NC * lane)) & 3);      // (r*npr + c0) mod 4
#pragma unroll
        for (int j = 0; j < NC; ++j) fx[j] = fb[off + NC * lane + j];
      }
      const double wr = (KIND == FK_RESID || KIND == FK_UPDATE_RESID0 || KIND == FK_PROLONG_SMOOTH) ? wrow(r) : 1.0;
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        double yx, yy, tx, ty;
        row_terms<false>(ke, Wc.u[j + 1], Wc.u[j + 2], Wp.u[j + 2], Wp.u[j + 1], tx, ty);   // NE element (r, c): node is local 0
        yx = Wp.e[j + 1] * tx; yy = Wp.e[j + 1] * ty;
        row_terms<true>(ke, Wc.u[j + 1], Wc.u[j], Wp.u[j], Wp.u[j + 1], tx, ty);            // NW (r, c-1): x-mirror image
        yx = fma(Wp.e[j], tx, yx); yy = fma(Wp.e[j], ty, yy);
        row_terms<false>(ke, Wc.u[j + 1], Wc.u[j], Wm.u[j], Wm.u[j + 1], tx, ty);           // SW (r-1, c-1): point-mirror image
        yx = fma(Wc.e[j], tx, yx); yy = fma(Wc.e[j], ty, yy);
        row_terms<true>(ke, Wc.u[j + 1], Wc.u[j + 2], Wm.u[j + 2], Wm.u[j + 1], tx, ty);    // SE (r-1, c): y-mirror image
        yx = fma(Wc.e[j + 1], tx, yx); yy = fma(Wc.e[j + 1], ty, yy);
        if (cl + j < npr) {
          const size_t n = n_out + j;
          const double2 *rw = raw + S_cur * NIN * SW + sl + j;             // raw inputs of this node
          const double2 v = Wc.u[j + 1];
          if (KIND == FK_APPLY) {
            if (fx[j] & 1u) yx = 0.0;
            if (fx[j] & 2u) yy = 0.0;
            a.out[0][n] = make_double2(yx, yy);
            if (a.want_dot) dots[0] = fma(v.x, yx, fma(v.y, yy, dots[0]));
          } else if (KIND == FK_RESID) {
            const double2 b = rw[SW];
            a.out[0][n] = make_double2((fx[j] & 1u) ? 0.0 : fma(wr, b.x, -yx), (fx[j] & 2u) ? 0.0 : fma(wr, b.y, -yy));
          } else if (KIND == FK_SMOOTH) {
            const double2 b = rw[SW], d = rw[2 * SW];
            a.out[0][n] = make_double2(fma(omega * d.x, b.x - yx, v.x), fma(omega * d.y, b.y - yy, v.y));
          } else if (KIND == FK_PCG_MATVEC || KIND == FK_PCG_MATVEC_J) {
            // Ap is NOT masked here: p vanishes at constrained DOFs (so p.Ap is unaffected) and the update
            // kernel masks r' = r - alpha Ap through D^-1 == 0
            if (KIND == FK_PCG_MATVEC_J) {
              const double2 d = rw[2 * SW];
              if (d.x == 0.0) yx = 0.0;
              if (d.y == 0.0) yy = 0.0;
            }
            a.out[1][n] = make_double2(yx, yy);
            dots[0] = fma(v.x, yx, fma(v.y, yy, dots[0]));
          } else if (KIND == FK_UPDATE_RESID0) {
            const double2 ap = rw[SW], rr = rw[3 * SW], d = raw_in<KIND, 4>(rw, SW);
            const double rnx = fma(-alpha, ap.x, rr.x), rny = fma(-alpha, ap.y, rr.y);
            reinterpret_cast<float2 *>(a.out[3])[n] =
                make_float2(d.x != 0.0 ? (float)fma(wr, rnx, -yx) : 0.f, d.y != 0.0 ? (float)fma(wr, rny, -yy) : 0.f);
          } else {  // FK_PROLONG_SMOOTH
            const double2 rr = rw[SW], d = raw_in<KIND, 2>(rw, SW);
            const double zx = fma(omega * d.x, fma(wr, rr.x, -yx), v.x), zy = fma(omega * d.y, fma(wr, rr.y, -yy), v.y);
            a.out[0][n] = make_double2(zx, zy);
            // interface rows: z is completed by z_lo + z_hi - x~, so each side contributes r.(z - (1-w) x~)
            const double wc = 1.0 - wr;
            dots[0] = fma(rr.x, zx - wc * v.x, fma(rr.y, zy - wc * v.y, dots[0]));
          }
        }
      }
      n_out += npr;
    };
    (void)wrow;
    // S rows per trip; window roles cycle through the 4 row images (4 divides S)
    for (int r = r0; r < r1; r += S) {
      step(IC<0>(), r, W0, W1, W2);
      if (r + 1 < r1) step(IC<1>(), r + 1, W1, W2, W3);
      if (r + 2 < r1) step(IC<2>(), r + 2, W2, W3, W0);
      if (r + 3 < r1) step(IC<3>(), r + 3, W3, W0, W1);
      if (S == 8) {
        if (r + 4 < r1) step(IC<4>(), r + 4, W0, W1, W2);
        if (r + 5 < r1) step(IC<5>(), r + 5, W1, W2, W3);
        if (r + 6 < r1) step(IC<6>(), r + 6, W2, W3, W0);
        if (r + 7 < r1) step(IC<7>(), r + 7, W3, W0, W1);
      }
    }
    cp_async_wait<0>();
  }

  // ---- reductions + scalar bookkeeping (last block) ----
  const int t = threadIdx.x;
  if (KIND == FK_APPLY) {
    if (a.want_dot) grid_reduce<1>(dots, a.partials, a.ticket, a.dot_out);
  } else if (KIND == FK_PCG_MATVEC || KIND == FK_PCG_MATVEC_J) {
    if (grid_reduce<1>(dots, a.partials, a.ticket, a.sc->aux) && t == 0 && !a.dist) a.sc->pAp = dots[0];
  } else if (KIND == FK_UPDATE_RESID0) {
    if (grid_reduce<1>(dots, a.partials, a.ticket, a.sc->aux + (a.dist ? 1 : 0)) && t == 0 && !a.dist) {
      SolverScalars *sc = a.sc;
      sc->rr = dots[0];
      if (!a.init) sc->iters += 1;
      if (dots[0] <= sc->tol2 * sc->bb || !(dots[0] == dots[0])) sc->done = 1;
      else if (sc->iters >= sc->maxit) sc->done = 2;
    }
  } else if (KIND == FK_PROLONG_SMOOTH) {
    if (grid_reduce<1>(dots, a.partials, a.ticket, a.sc->aux + (a.dist ? 2 : 0)) && t == 0 && !a.dist) {
      SolverScalars *sc = a.sc;
      sc->rz[1] = a.init ? 0.0 : dots[0] / sc->rz[0];      // beta for the next direction update
      sc->rz[0] = dots[0];
    }
  }
}

template <int KIND, int NC>
int launch_fine_nc(topo_handle *h, FineArgs &a) {
  static int resident_dev[TOPO_MAX_DEVICES] = {};          // CTAs of this variant that fit one SM (0: not set up yet)
  int &resident = resident_dev[topo_dev_slot(h)];
  const size_t smem = 4 * (size_t)warp_smem_bytes<KIND, NC>();
  if (resident == 0) {
    CUDA_TRY(cudaFuncSetAttribute(k_fine<KIND, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_fine<KIND, NC>, BW, smem));
    resident = std::max(1, occ);
  }
  // chunk rows so that one wave of CTAs covers the mesh (TOPO_FINE_WAVES: more, shorter chunks -- experiment knob)
  const double waves = h->knobs.fine_waves;
  const int bx = div_up(h->nx + 1, 4 * Geo<NC>::WC);
  const int chunks = std::max(1, (int)(waves * 148 * resident) / bx);
  a.rows_per_chunk = std::max(8, div_up(h->ny + 1, chunks));
  dim3 grid(bx, div_up(h->ny + 1, a.rows_per_chunk));
  if ((int64_t)grid.x * grid.y > TOPO_MAX_PARTIAL_BLOCKS) {
    topo_set_error("fine-grid launch too large for the reduction buffer");
    return TOPO_EINVAL;
  }
  KeRow kr;
  for (int j = 0; j < 8; ++j) { kr.r0[j] = h->ke.k[j]; kr.r1[j] = h->ke.k[8 + j]; }
  static const char *const names[7] = {"k_fine<apply>", "k_fine<resid>", "k_fine<pcg_matvec>", "k_fine<pcg_matvec_jacobi>",
                                       "k_fine<update_resid0>", "k_fine<smooth_dot>", "k_fine<smooth>"};
  KL(h, names[KIND], CUDA_TRY(topo_launch(h, k_fine<KIND, NC>, grid, dim3(BW), smem, kr, a)));
  CUDA_TRY(cudaGetLastError());
  return TOPO_OK;
}

template <int KIND>
int launch_fine(topo_handle *h, FineArgs &a) {
  a.scale = h->scale;
  a.fixed = h->fixed;
  a.nx = h->nx;
  a.ny = h->ny;
  a.partials = h->partials;
  a.ticket = h->ticket;
  a.sc = h->sc;
  a.w_first = h->dist.on ? h->dist.w_first : 1.0;
  a.w_last = h->dist.on ? h->dist.w_last : 1.0;
  a.dist = h->dist.on ? 1 : 0;
  // node columns per lane: 2 for the plain product on wide meshes (overhead instructions amortised over two
  // nodes), 1 for the fused kinds (measured: they lose more from the lower occupancy of 64-column strips than
  // they gain); TOPO_FINE_NC_MASK (bit KIND set = two columns for that kind) is the experiment knob
  const int nc_mask = h->knobs.fine_nc_mask;
  const int nc = (KIND == FK_APPLY) ? h->fine_nc : (((nc_mask >> KIND) & 1) && h->nx + 1 >= 1024 ? 2 : 1);
  if (nc == 2) return launch_fine_nc<KIND, 2>(h, a);
  return launch_fine_nc<KIND, 1>(h, a);
}

// diag(K) per node from the 4 adjacent element scales; stores 1/diag (0 at constrained DOFs) and the
// max diagonal over free DOFs (= K.max() of the reference's equilibrium guard, optimize.py:455: the
// largest entry of an SPD matrix sits on its diagonal).
__global__ void k_build_diag(const __grid_constant__ KeParam ke, const double *__restrict__ scale,
                             const uint8_t *__restrict__ fixed, double2 *__restrict__ dinv,
                             float2 *__restrict__ dinv_f, int nx, int ny, double *partials, unsigned int *ticket,
                             double *kmax_out) {
  const int64_t nn = (int64_t)(nx + 1) * (ny + 1);
  double mx = 0.0;
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < nn; n += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(n / (nx + 1)), c = (int)(n - (int64_t)r * (nx + 1));
    const double ne = (r < ny && c < nx) ? scale[(size_t)r * nx + c] : 0.0;
    const double nw = (r < ny && c > 0) ? scale[(size_t)r * nx + c - 1] : 0.0;
    const double sw = (r > 0 && c > 0) ? scale[(size_t)(r - 1) * nx + c - 1] : 0.0;
    const double se = (r > 0 && c < nx) ? scale[(size_t)(r - 1) * nx + c] : 0.0;
    const double dx = ne * ke.k[0 * 8 + 0] + nw * ke.k[2 * 8 + 2] + sw * ke.k[4 * 8 + 4] + se * ke.k[6 * 8 + 6];
    const double dy = ne * ke.k[1 * 8 + 1] + nw * ke.k[3 * 8 + 3] + sw * ke.k[5 * 8 + 5] + se * ke.k[7 * 8 + 7];
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
    dinv_f[n] = make_float2((float)ix, (float)iy);
  }
  double v[1] = {mx};
  grid_reduce<1, true>(v, partials, ticket, kmax_out);
}

}  // namespace

// ---- drivers -------------------------------------------------------------------------------------------
int topo_fine_matvec(topo_handle *h, const double2 *u, double2 *y) {
  FineArgs a = {};
  a.in[0] = u;
  a.out[0] = y;
  return launch_fine<FK_APPLY>(h, a);
}

int topo_fine_residual(topo_handle *h, const double2 *u, const double2 *b, double2 *r) {
  FineArgs a = {};
  a.in[0] = u;
  a.in[1] = b;
  a.out[0] = r;
  return launch_fine<FK_RESID>(h, a);
}

int topo_fine_smooth(topo_handle *h, const double2 *x, const double2 *b, double2 *x_out, double omega) {
  FineArgs a = {};
  a.in[0] = x;
  a.in[1] = b;
  a.in[2] = h->dinv;
  a.out[0] = x_out;
  a.omega = omega;
  return launch_fine<FK_SMOOTH>(h, a);
}

// p_out = z + beta p_in (beta from sc) ; Ap = A p_out ; sc->pAp
int topo_fine_pcg_matvec(topo_handle *h, const double2 *z, const double2 *p_in, double2 *p_out, double2 *Ap) {
  FineArgs a = {};
  a.in[0] = z;
  a.in[1] = p_in;
  a.out[0] = p_out;
  a.out[1] = Ap;
  a.check_done = 1;
  return launch_fine<FK_PCG_MATVEC>(h, a);
}

// Jacobi flavour: p_out = dinv*r + beta p_in
int topo_fine_pcg_matvec_jacobi(topo_handle *h, const double2 *r, const double2 *p_in, double2 *p_out, double2 *Ap) {
  FineArgs a = {};
  a.in[0] = r;
  a.in[1] = p_in;
  a.in[2] = h->dinv;
  a.out[0] = p_out;
  a.out[1] = Ap;
  a.check_done = 1;
  return launch_fine<FK_PCG_MATVEC_J>(h, a);
}

// u_out = u_in + alpha p ; r_out = r_in - alpha Ap ; rr, convergence flag ; xhat = omega dinv r_out ;
// res = r_out - A xhat
int topo_fine_update_resid0(topo_handle *h, const double2 *p, const double2 *Ap, const double2 *u_in,
                            const double2 *r_in, double2 *u_out, double2 *r_out, double2 *xhat, double2 *res,
                            double omega, int init) {
  FineArgs a = {};
  a.in[0] = p;
  a.in[1] = Ap;
  a.in[2] = u_in;
  a.in[3] = r_in;
  a.in[4] = reinterpret_cast<const double2 *>(h->dinv_f);     // fp32 copy (FK<FK_UPDATE_RESID0>::F32MASK)
  a.out[0] = u_out;
  a.out[1] = r_out;
  a.out[2] = xhat;
  a.out[3] = res;
  a.omega = omega;
  a.init = init;
  a.check_done = init ? 0 : 1;
  return launch_fine<FK_UPDATE_RESID0>(h, a);
}

// z = x~ + omega dinv (r - A x~), x~ = xhat + P xc ; sc->rz, beta
int topo_fine_prolong_smooth(topo_handle *h, const double2 *xhat, const double2 *r, const double2 *xc, int cnx,
                             double2 *z, double omega, int init) {
  FineArgs a = {};
  a.in[0] = xhat;                                             // float2 array
  a.in[1] = r;
  a.in[2] = reinterpret_cast<const double2 *>(h->dinv_f);     // fp32 copy
  a.out[0] = z;
  a.coarse = xc;
  a.cnx = cnx;
  a.omega = omega;
  a.init = init;
  a.check_done = init ? 0 : 1;
  return launch_fine<FK_PROLONG_SMOOTH>(h, a);
}

int topo_build_diag(topo_handle *h) {
  const int blocks = (int)std::min<int64_t>(div_up(h->nn, 256), 1184);
  KL(h, "k_build_diag", k_build_diag<<<blocks, 256, 0, h->stream>>>(h->ke, h->scale, h->fixed, h->dinv, h->dinv_f, h->nx, h->ny,
                                                                      h->partials, h->ticket, h->kmax));
  CUDA_TRY(cudaGetLastError());
  h->diag_valid = true;
  return TOPO_OK;
}
