// Device helpers shared by the multigrid V-cycle translation units (gmg_cycle.cu: level kernels + generic tail;
// gmg_tail.cu: the distributed-shared-memory tail): transfer rules, node indexing, the stencil product, LevelView.
#pragma once
#include "topo_internal.cuh"

namespace {

// ---- transfers: 1-D rules --------------------------------------------------------------------------------------
// Coarse node I sits at fine index X = min(2I, nf).  Fine node i is a coarse node if i is even or
// i == nf; otherwise it interpolates (1/2, 1/2) between coarse (i-1)/2 and (i+1)/2.
__device__ __forceinline__ void restrict_taps(int I, int nf, int &X, double &wl, double &wr) {
  X = min(2 * I, nf);
  const bool even = (X == 2 * I);
  wl = (even && I >= 1) ? 0.5 : 0.0;
  wr = (even && X + 1 < nf) ? 0.5 : 0.0;
}
__device__ __forceinline__ void prolong_taps(int i, int nf, int &I0, int &I1, double &w0, double &w1) {
  if ((i & 1) == 0) { I0 = I1 = i >> 1; w0 = 1.0; w1 = 0.0; }
  else if (i == nf) { I0 = I1 = (i + 1) >> 1; w0 = 1.0; w1 = 0.0; }
  else { I0 = (i - 1) >> 1; I1 = (i + 1) >> 1; w0 = 0.5; w1 = 0.5; }
}

template <class V2>
__device__ __forceinline__ double2 restrict_node(const V2 *__restrict__ rf, int I, int J, int nxf, int nyf) {
  const int nprf = nxf + 1;
  int X, Y;
  double wxl, wxr, wyl, wyr;
  restrict_taps(I, nxf, X, wxl, wxr);
  restrict_taps(J, nyf, Y, wyl, wyr);
  const double wx[3] = {wxl, 1.0, wxr}, wy[3] = {wyl, 1.0, wyr};
  double sx = 0.0, sy = 0.0;
#pragma unroll
  for (int dj = -1; dj <= 1; ++dj) {
    if (wy[dj + 1] == 0.0) continue;
#pragma unroll
    for (int di = -1; di <= 1; ++di) {
      const double w = wy[dj + 1] * wx[di + 1];
      if (w == 0.0) continue;
      const V2 v = rf[(size_t)(Y + dj) * nprf + (X + di)];
      sx = fma(w, (double)v.x, sx);
      sy = fma(w, (double)v.y, sy);
    }
  }
  return make_double2(sx, sy);
}

// (P xc) at fine node (i, j)
__device__ __forceinline__ double2 prolong_node(const double2 *__restrict__ xc, int i, int j, int nxf, int nyf,
                                                int nprc) {
  int I0, I1, J0, J1;
  double wx0, wx1, wy0, wy1;
  prolong_taps(i, nxf, I0, I1, wx0, wx1);
  prolong_taps(j, nyf, J0, J1, wy0, wy1);
  const double2 v00 = xc[(size_t)J0 * nprc + I0];
  double px = wy0 * wx0 * v00.x, py = wy0 * wx0 * v00.y;
  if (wx1 != 0.0) {
    const double2 v = xc[(size_t)J0 * nprc + I1];
    px = fma(wy0 * wx1, v.x, px); py = fma(wy0 * wx1, v.y, py);
  }
  if (wy1 != 0.0) {
    const double2 v = xc[(size_t)J1 * nprc + I0];
    px = fma(wy1 * wx0, v.x, px); py = fma(wy1 * wx0, v.y, py);
    if (wx1 != 0.0) {
      const double2 w = xc[(size_t)J1 * nprc + I1];
      px = fma(wy1 * wx1, w.x, px); py = fma(wy1 * wx1, w.y, py);
    }
  }
  return make_double2(px, py);
}

// (row, column) of node n on a level with npr nodes per row.  Node counts stay below 2^30 (topo_create refuses
// larger meshes), so the 32-bit division is exact -- the 64-bit one is a ~60-instruction subroutine, a third of the
// dynamic instructions of the transfer kernels.
__device__ __forceinline__ void rc_of(int64_t n, int npr, int &r, int &c) {
  const unsigned un = (unsigned)n, q = un / (unsigned)npr;
  r = (int)q;
  c = (int)(un - q * (unsigned)npr);
}

// A x at node n of a stencil level, x of each neighbour supplied by `get(rr, cc)`.
// sym: the operator is symmetric, so the block towards a "backward" neighbour m (W, SE, S, SW) is the transpose of
// the block node m stores towards n.  Reading planes 4..8 only (self, E, NW, N, NE -- at n and at the four backward
// neighbours, whose entries sit in cache lines the neighbouring threads fetch anyway) cuts the stencil stream of a
// level from 144 to 80 B per node; the big levels are bound by exactly that stream.
template <bool SYM, class Get>
__device__ __forceinline__ void stencil_row(const stencil_t *__restrict__ st, int64_t nn, int64_t n, int r, int c, int nx,
                                            int ny, Get get, double &ax, double &ay) {
  ax = 0.0;
  ay = 0.0;
  const float4 *__restrict__ pl = reinterpret_cast<const float4 *>(st);
  const int npr = nx + 1;
#pragma unroll
  for (int dr = -1; dr <= 1; ++dr) {
#pragma unroll
    for (int dc = -1; dc <= 1; ++dc) {
      const int rr = r + dr, cc = c + dc;
      if (rr < 0 || rr > ny || cc < 0 || cc > nx) continue;
      const int nb = 3 * (dr + 1) + (dc + 1);
      const double2 xv = get(rr, cc);
      if (SYM && nb < 4) {
        // A[n, m] = A[m, n]^T ; node m keeps A[m, n] in plane 8 - nb
        const float4 cf = pl[(size_t)(8 - nb) * nn + (n + (int64_t)dr * npr + dc)];
        ax = fma((double)cf.x, xv.x, ax);
        ax = fma((double)cf.z, xv.y, ax);
        ay = fma((double)cf.y, xv.x, ay);
        ay = fma((double)cf.w, xv.y, ay);
      } else {
        // one float4 per neighbour block: (a00, a01, a10, a11), plane nb, consecutive nodes contiguous
        const float4 cf = pl[(size_t)nb * nn + n];
        ax = fma((double)cf.x, xv.x, ax);
        ax = fma((double)cf.y, xv.y, ax);
        ay = fma((double)cf.z, xv.x, ay);
        ay = fma((double)cf.w, xv.y, ay);
      }
    }
  }
}

struct LevelView {
  const stencil_t *st;
  const double2 *dinv;
  double2 *x, *x2, *b, *r;
  int nx, ny;
  double wf, wl;       // row-slab mode: weights of node rows 0 / ny that make a distributed copy of b (else 1)
  int sym;             // read the backward half of the stencil as transposes of the neighbours' forward blocks
};
__device__ __forceinline__ double row_weight(const LevelView &L, int r) { return r == 0 ? L.wf : (r == L.ny ? L.wl : 1.0); }

}  // namespace
