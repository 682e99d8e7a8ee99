// Geometric multigrid V-cycle used as the PCG preconditioner (stage 2; no counterpart in the
// reference, which calls SuperLU at optimize.py:437).
//
// Hierarchy: level 0 is the matrix-free fine grid; level l+1 halves level l in both directions
// (ceil, so any nx, ny coarsens down to <= 2x2 cells: a trailing single child is allowed).  Coarse
// operators are exact Galerkin products of the CONSTRAINED fine operator, A_{l+1} = P^T A_l P with
// bilinear P, built cell by cell: K_C = sum_children P_c^T K_c P_c (bilinear spaces nest, so the sum
// over the <=4 children is the whole product).  Level 1 is built straight from the element scales
// (K_c = scale_c * ke; children touching constrained DOFs get their rows/cols masked).  Cell matrices
// are then gathered into 3x3-block node stencils (36 doubles per node, SoA planes) which is what the
// smoother / residual kernels stream.  Smoother: damped Jacobi.  Coarsest level: dense inverse.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "topo_internal.cuh"

namespace {


// ---- child geometry ---------------------------------------------------------------------------
// 1-D child type t: 0 = first child of a 2-cell parent, 1 = second child, 2 = only child (parent == child)
__host__ __device__ inline double l1d(int t, int fine_local, int coarse_local) {
  // weight of coarse local node (0 left, 1 right) at fine local node (0 left, 1 right)
  if (t == 2) return fine_local == coarse_local ? 1.0 : 0.0;
  if (t == 0) return fine_local == 0 ? (coarse_local == 0 ? 1.0 : 0.0) : 0.5;
  return fine_local == 1 ? (coarse_local == 1 ? 1.0 : 0.0) : 0.5;
}
// local node a (CCW from lower-left) -> (ix, iy)
__host__ __device__ inline int lx(int a) { return (a == 1 || a == 2) ? 1 : 0; }
__host__ __device__ inline int ly(int a) { return (a >= 2) ? 1 : 0; }

struct PwTable {
  double w[9][4][4];   // [variant = 3*ty+tx][fine local a][coarse local A]
};
struct GTable {
  double g[9][64];     // P_v^T ke P_v
};

__constant__ PwTable c_pw;
__constant__ GTable c_g;

// ---- level-1 cells straight from element scales -------------------------------------------------
__device__ __noinline__ void masked_child(const KeParam &ke, int v, double s, unsigned freebits, double *acc) {
  // acc = s * (m o P)^T ke (m o P): freebits bit (2a+d) set = fine DOF free (acc is a scratch array)
  for (int i = 0; i < 8; ++i) {
    const int A = i >> 1, d = i & 1;
    for (int j = 0; j < 8; ++j) {
      const int B = j >> 1, e = j & 1;
      double t = 0.0;
      for (int a = 0; a < 4; ++a) {
        const double wa = c_pw.w[v][a][A];
        if (wa == 0.0 || !((freebits >> (2 * a + d)) & 1u)) continue;
        for (int b = 0; b < 4; ++b) {
          const double wb = c_pw.w[v][b][B];
          if (wb == 0.0 || !((freebits >> (2 * b + e)) & 1u)) continue;
          t = fma(wa * wb, ke.k[(2 * a + d) * 8 + 2 * b + e], t);
        }
      }
      acc[i * 8 + j] = s * t;
    }
  }
}

__global__ void __launch_bounds__(128) k_cells_level1(const __grid_constant__ KeParam ke,
                                                      const double *__restrict__ scale,
                                                      const uint8_t *__restrict__ fixed, double *__restrict__ cell,
                                                      int nxf, int nyf, int nxc, int nyc) {
  const int64_t ncell = (int64_t)nxc * nyc;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ncell) return;
  const int J = (int)(t / nxc), I = (int)(t - (int64_t)J * nxc);
  double acc[64];
#pragma unroll
  for (int k = 0; k < 64; ++k) acc[k] = 0.0;
  const int sx = min(2, nxf - 2 * I), sy = min(2, nyf - 2 * J);
  for (int cy = 0; cy < sy; ++cy) {
    for (int cx = 0; cx < sx; ++cx) {
      const int ex = 2 * I + cx, ey = 2 * J + cy;
      const int v = 3 * (sy == 2 ? cy : 2) + (sx == 2 ? cx : 2);
      const double s = scale[(size_t)ey * nxf + ex];
      const size_t n0 = (size_t)ey * (nxf + 1) + ex;
      const unsigned f0 = fixed[n0], f1 = fixed[n0 + 1], f2 = fixed[n0 + nxf + 2], f3 = fixed[n0 + nxf + 1];
      const unsigned fixedbits = (f0 & 3u) | ((f1 & 3u) << 2) | ((f2 & 3u) << 4) | ((f3 & 3u) << 6);
      if (fixedbits == 0u) {
#pragma unroll
        for (int k = 0; k < 64; ++k) acc[k] = fma(s, c_g.g[v][k], acc[k]);
      } else if (fixedbits != 0xffu) {
        double tmpm[64];                       // rare path (children touching constrained DOFs)
        masked_child(ke, v, s, (~fixedbits) & 0xffu, tmpm);
#pragma unroll
        for (int k = 0; k < 64; ++k) acc[k] += tmpm[k];
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 64; ++k) cell[(size_t)k * ncell + t] = acc[k];
}

// ---- level 1 and level 2 WITHOUT the level-1 cell array --------------------------------------------------------
// k_cells_level1 writes 64 doubles per level-1 cell (268 MB on 2048x1024, 4.3 GB on 8192x4096) that are read exactly
// twice: by the level-1 stencil gather and by the level-2 coarsening.  The two kernels below produce those two results
// straight from the element scales, evaluating each level-1 cell entry they need with the SAME fma chain over the
// cell's children (so stencils and level-2 cells are bit-identical to the two-step path): 0.45 -> 0.1 ms of the
// multigrid set-up per design iteration, and the array is no longer allocated.
// entry (i, j) of (m o P_v)^T ke (m o P_v) (the rare children that touch constrained DOFs)
__device__ __noinline__ double masked_entry(const KeParam &ke, int v, unsigned freebits, int i, int j) {
  const int A = i >> 1, d = i & 1, B = j >> 1, e = j & 1;
  double t = 0.0;
  for (int a = 0; a < 4; ++a) {
    const double wa = c_pw.w[v][a][A];
    if (wa == 0.0 || !((freebits >> (2 * a + d)) & 1u)) continue;
    for (int b = 0; b < 4; ++b) {
      const double wb = c_pw.w[v][b][B];
      if (wb == 0.0 || !((freebits >> (2 * b + e)) & 1u)) continue;
      t = fma(wa * wb, ke.k[(2 * a + d) * 8 + 2 * b + e], t);
    }
  }
  return t;
}
// NE entries of level-1 cell (I, J), entry q at matrix index idx(q): acc[q] = sum over the cell's children in
// (cy, cx) order, exactly as k_cells_level1 accumulates them
// scales and constraint bits of the (up to) four children of level-1 cell (I, J): ALL loads are issued before anything
// depends on them (with the loads inside the child loop a thread paid one L2 round trip per child, 140 us per launch)
struct L1Children {
  double s[4];          // child (cy, cx) at 2*cy + cx; 0 where the child does not exist
  unsigned fb[4];       // fixed bits of its 8 DOFs (0xff: skip -- missing child or fully constrained)
  int v[4];             // interpolation variant
  bool full;            // 2 x 2 children: variant of child c is 3*(c>>1) + (c&1), a compile-time constant-bank address
};
__device__ __forceinline__ void level1_load(const double *__restrict__ scale, const uint8_t *__restrict__ fixed, int nxf,
                                            int nyf, int I, int J, bool cell_ok, L1Children &ch) {
  const int sx = min(2, nxf - 2 * I), sy = min(2, nyf - 2 * J);
  ch.full = (sx == 2 && sy == 2);
  unsigned f[3][3];
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const bool ok = cell_ok && i <= sx && j <= sy;
      f[j][i] = ok ? (unsigned)__ldg(fixed + (size_t)(2 * J + j) * (nxf + 1) + (2 * I + i)) : 3u;
    }
#pragma unroll
  for (int cy = 0; cy < 2; ++cy)
#pragma unroll
    for (int cx = 0; cx < 2; ++cx) {
      const bool ok = cell_ok && cx < sx && cy < sy;
      ch.s[2 * cy + cx] = ok ? __ldg(scale + (size_t)(2 * J + cy) * nxf + (2 * I + cx)) : 0.0;
      ch.fb[2 * cy + cx] = ok ? ((f[cy][cx] & 3u) | ((f[cy][cx + 1] & 3u) << 2) | ((f[cy + 1][cx + 1] & 3u) << 4) |
                                 ((f[cy + 1][cx] & 3u) << 6)) : 0xffu;
      ch.v[2 * cy + cx] = 3 * (sy == 2 ? cy : 2) + (sx == 2 ? cx : 2);
    }
}
// NE entries of a level-1 cell, entry q at matrix index idx(q): acc[q] = sum over the cell's children in (cy, cx)
// order, exactly as k_cells_level1 accumulates them
template <int NE, class Idx>
__device__ __forceinline__ void level1_entries(const KeParam &ke, const L1Children &ch, Idx idx, double (&acc)[NE]) {
#pragma unroll
  for (int q = 0; q < NE; ++q) acc[q] = 0.0;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const double s = ch.s[c];
    const unsigned fixedbits = ch.fb[c];
    const int v = ch.v[c];
    if (fixedbits == 0u) {
      // (a run-time variant index makes every coefficient an indexed constant load: 256 per thread, the whole cost of
      // the first version; cells with all four children address the table at compile time)
      if (ch.full) {
#pragma unroll
        for (int q = 0; q < NE; ++q) acc[q] = fma(s, c_g.g[3 * (c >> 1) + (c & 1)][idx(q)], acc[q]);
      } else {
#pragma unroll
        for (int q = 0; q < NE; ++q) acc[q] = fma(s, c_g.g[v][idx(q)], acc[q]);
      }
    } else if (fixedbits != 0xffu) {
#pragma unroll
      for (int q = 0; q < NE; ++q) {                   // (unrolled: a dynamic index would move acc[] to local memory)
        const int k = idx(q);
        acc[q] += s * masked_entry(ke, v, (~fixedbits) & 0xffu, k >> 3, k & 7);
      }
    }
  }
}

// ---- branch-free path of the two kernels below ----------------------------------------------------------------------------
// ncu (profiles/r2_ncu_kernels.txt): k_stencil_level1 and k_cells_level2 execute ~2000 instructions per thread for 256-768
// FMAs -- the rest is the bookkeeping of cells cut by the mesh edge and of children that touch constrained DOFs (9 flag
// bytes per level-1 cell, bit packing, run-time interpolation variants: indexed constant loads), at 144-154 registers.
// Away from the mesh edge and from the supports none of it is needed: every cell has its four children, every variant
// index is a compile-time constant and the node weights of a child are the constants 1, 1/2, 1/4, 0.  k_interior_flags
// marks those level-1 nodes / level-2 cells once per set-up; the FAST instantiations handle them with the SAME fma chains
// in the same order (terms with an exactly zero weight dropped: fma(m, 0, s) == s), so the operators are bit-identical;
// the general instantiations handle the rest (a strip along the mesh edge and the supports).
__device__ __forceinline__ void level1_load_fast(const double *__restrict__ scale, int nxf, int I, int J, double (&s)[4]) {
#pragma unroll
  for (int cy = 0; cy < 2; ++cy)
#pragma unroll
    for (int cx = 0; cx < 2; ++cx) s[2 * cy + cx] = __ldg(scale + (size_t)(2 * J + cy) * nxf + (2 * I + cx));
}
template <int NE, class Idx>
__device__ __forceinline__ void level1_entries_fast(const double (&s)[4], Idx idx, double (&acc)[NE]) {
#pragma unroll
  for (int q = 0; q < NE; ++q) acc[q] = 0.0;
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int q = 0; q < NE; ++q) acc[q] = fma(s[c], c_g.g[3 * (c >> 1) + (c & 1)][idx(q)], acc[q]);
}
__device__ __forceinline__ bool fixed_block_clear(const uint8_t *__restrict__ fixed, int npr0, int r0, int c0, int rows, int cols) {
  unsigned any = 0u;
  for (int j = 0; j < rows; ++j)
    for (int i = 0; i < cols; ++i) any |= (unsigned)__ldg(fixed + (size_t)(r0 + j) * npr0 + (c0 + i));
  return (any & 3u) == 0u;
}
// fine mesh nx0 x ny0, level 1 nx1 x ny1 cells, level 2 nx2 x ny2 cells
__global__ void k_interior_flags(const uint8_t *__restrict__ fixed, int nx0, int ny0, int nx1, int ny1, int nx2, int ny2,
                                 uint8_t *__restrict__ node1, uint8_t *__restrict__ cell2, int enable) {
  const int64_t nn1 = (int64_t)(nx1 + 1) * (ny1 + 1), nc2 = (int64_t)nx2 * ny2;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nn1) {
    const int r = (int)(t / (nx1 + 1)), c = (int)(t - (int64_t)r * (nx1 + 1));
    // the four cells around the node exist and have 2 x 2 children; no constrained DOF among the 5 x 5 fine nodes around
    bool ok = enable && r >= 1 && r <= ny1 - 1 && c >= 1 && c <= nx1 - 1 && 2 * c + 2 <= nx0 && 2 * r + 2 <= ny0;
    if (ok) ok = fixed_block_clear(fixed, nx0 + 1, 2 * r - 2, 2 * c - 2, 5, 5);
    node1[t] = ok ? 1 : 0;
  }
  if (cell2 != nullptr && t < nc2) {
    const int J = (int)(t / nx2), I = (int)(t - (int64_t)J * nx2);
    // 2 x 2 children with 2 x 2 children each; no constrained DOF among the cell's 5 x 5 fine nodes
    bool ok = enable && 2 * I + 2 <= nx1 && 2 * J + 2 <= ny1 && 4 * I + 4 <= nx0 && 4 * J + 4 <= ny0;
    if (ok) ok = fixed_block_clear(fixed, nx0 + 1, 4 * J, 4 * I, 5, 5);
    cell2[t] = ok ? 1 : 0;
  }
}

// level-1 node stencil (and D^-1) from the element scales: per adjacent cell the two matrix rows of this node
template <bool FAST>
__global__ void __launch_bounds__(128) k_stencil_level1(const __grid_constant__ KeParam ke, const double *__restrict__ scale,
                                                        const uint8_t *__restrict__ fixed, stencil_t *__restrict__ st,
                                                        double2 *__restrict__ dinv, int nxf, int nyf, int nx, int ny,
                                                        const uint8_t *__restrict__ interior) {
  const int npr = nx + 1;
  const int64_t nn = (int64_t)npr * (ny + 1);
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nn) return;
  if ((interior[n] != 0) != FAST) return;                // (the other instantiation handles this node)
  const int r = (int)(n / npr), c = (int)(n - (int64_t)r * npr);
  double s[36];
#pragma unroll
  for (int k = 0; k < 36; ++k) s[k] = 0.0;
  const int cro[4] = {0, 0, -1, -1}, cco[4] = {0, -1, -1, 0};
  L1Children ch[FAST ? 1 : 4];
  double sc[FAST ? 4 : 1][4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int cr = r + cro[q], cc = c + cco[q];
    if (FAST) level1_load_fast(scale, nxf, cc, cr, sc[FAST ? q : 0]);
    else level1_load(scale, fixed, nxf, nyf, cc, cr, !(cr < 0 || cr >= ny || cc < 0 || cc >= nx), ch[FAST ? 0 : q]);
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int cr = r + cro[q], cc = c + cco[q];
    if (!FAST && (cr < 0 || cr >= ny || cc < 0 || cc >= nx)) continue;
    const int a = q;                                     // local index of this node in that cell
    double part[16];                                     // rows 2a, 2a+1 of the cell matrix
    if (FAST) level1_entries_fast<16>(sc[FAST ? q : 0], [a](int t) { return (2 * a + (t >> 3)) * 8 + (t & 7); }, part);
    else level1_entries<16>(ke, ch[FAST ? 0 : q], [a](int t) { return (2 * a + (t >> 3)) * 8 + (t & 7); }, part);
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int dr = cr + ly(b) - r, dc = cc + lx(b) - c;
      const int nb = 3 * (dr + 1) + (dc + 1);
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) s[4 * nb + 2 * i + j] += part[i * 8 + 2 * b + j];
    }
  }
#pragma unroll
  for (int nb = 0; nb < 9; ++nb)
    reinterpret_cast<float4 *>(st)[(size_t)nb * nn + n] =
        make_float4((float)s[4 * nb], (float)s[4 * nb + 1], (float)s[4 * nb + 2], (float)s[4 * nb + 3]);
  const double dx = (double)(stencil_t)s[4 * 4 + 0], dy = (double)(stencil_t)s[4 * 4 + 3];
  dinv[n] = make_double2(dx > 0.0 ? 1.0 / dx : 0.0, dy > 0.0 ? 1.0 / dy : 0.0);
}

// level-2 cells from the element scales: k_cells_coarsen with the child (level-1) cell entries computed on the fly
// FAST (cells marked by k_interior_flags): the component pair (d, e) is a compile-time parameter (the matrix indices of
// the Galerkin table become immediates), the child weights are constants and their zero terms are not executed
template <bool FAST, int DE>
__global__ void __launch_bounds__(128) k_cells_level2(const __grid_constant__ KeParam ke, const double *__restrict__ scale,
                                                      const uint8_t *__restrict__ fixed, double *__restrict__ ccell, int nx0,
                                                      int ny0, int nxf, int nyf, int nxc, int nyc,
                                                      const uint8_t *__restrict__ interior) {
  const int64_t ncc = (int64_t)nxc * nyc;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ncc) return;
  if ((interior[t] != 0) != FAST) return;                // (the other instantiation handles this cell)
  const int d = FAST ? (DE >> 1) : (int)(blockIdx.y >> 1), e = FAST ? (DE & 1) : (int)(blockIdx.y & 1);
  const int J = (int)(t / nxc), I = (int)(t - (int64_t)J * nxc);
  const int sx = FAST ? 2 : min(2, nxf - 2 * I), sy = FAST ? 2 : min(2, nyf - 2 * J);
  double acc[4][4];
#pragma unroll
  for (int A = 0; A < 4; ++A)
#pragma unroll
    for (int B = 0; B < 4; ++B) acc[A][B] = 0.0;
  L1Children ch[FAST ? 1 : 4];
  double sc[FAST ? 4 : 1][4];
#pragma unroll
  for (int cy = 0; cy < 2; ++cy)
#pragma unroll
    for (int cx = 0; cx < 2; ++cx) {
      if (FAST) level1_load_fast(scale, nx0, 2 * I + cx, 2 * J + cy, sc[FAST ? 2 * cy + cx : 0]);
      else level1_load(scale, fixed, nx0, ny0, 2 * I + cx, 2 * J + cy, cx < sx && cy < sy, ch[FAST ? 0 : 2 * cy + cx]);
    }
#pragma unroll
  for (int cy = 0; cy < 2; ++cy) {
#pragma unroll
    for (int cx = 0; cx < 2; ++cx) {
      if (cx >= sx || cy >= sy) continue;
      const int v = 3 * (sy == 2 ? cy : 2) + (sx == 2 ? cx : 2);
      double m16[16];                                    // M[a][b] = child cell entry (2a+d, 2b+e)
      if (FAST) level1_entries_fast<16>(sc[FAST ? 2 * cy + cx : 0], [d, e](int q) { return (2 * (q >> 2) + d) * 8 + 2 * (q & 3) + e; }, m16);
      else level1_entries<16>(ke, ch[FAST ? 0 : 2 * cy + cx], [d, e](int q) { return (2 * (q >> 2) + d) * 8 + 2 * (q & 3) + e; }, m16);
      double W[4][4], T[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int A = 0; A < 4; ++A) W[a][A] = FAST ? l1d(cx, lx(a), lx(A)) * l1d(cy, ly(a), ly(A)) : c_pw.w[v][a][A];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int B = 0; B < 4; ++B) {
          double s = 0.0;
#pragma unroll
          for (int b = 0; b < 4; ++b)
            if (!FAST || W[b][B] != 0.0) s = fma(m16[4 * a + b], W[b][B], s);
          T[a][B] = s;
        }
#pragma unroll
      for (int A = 0; A < 4; ++A)
#pragma unroll
        for (int B = 0; B < 4; ++B) {
          double s = acc[A][B];
#pragma unroll
          for (int a = 0; a < 4; ++a)
            if (!FAST || W[a][A] != 0.0) s = fma(W[a][A], T[a][B], s);
          acc[A][B] = s;
        }
    }
  }
#pragma unroll
  for (int A = 0; A < 4; ++A)
#pragma unroll
    for (int B = 0; B < 4; ++B) ccell[(size_t)((2 * A + d) * 8 + 2 * B + e) * ncc + t] = acc[A][B];
}

// ---- generic Galerkin coarsening of cell matrices ---------------------------------------------------
// K_C = sum_children P^T K_c P.  P couples equal displacement components only, so for each component pair
// (d,e) the 4x4 node block M_c[a][b] = K_c[2a+d][2b+e] transforms independently: K_C^{de} += W^T M_c W with
// the 4x4 node-weight matrix W of the child.  thread = (coarse cell, (d,e)): 16 accumulators in registers,
// 16 coalesced plane loads per child.
__global__ void __launch_bounds__(128) k_cells_coarsen(const double *__restrict__ fcell, double *__restrict__ ccell,
                                                       int nxf, int nyf, int nxc, int nyc) {
  const int64_t ncc = (int64_t)nxc * nyc, ncf = (int64_t)nxf * nyf;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ncc) return;
  const int d = blockIdx.y >> 1, e = blockIdx.y & 1;
  const int J = (int)(t / nxc), I = (int)(t - (int64_t)J * nxc);
  const int sx = min(2, nxf - 2 * I), sy = min(2, nyf - 2 * J);
  double acc[4][4];
#pragma unroll
  for (int A = 0; A < 4; ++A)
#pragma unroll
    for (int B = 0; B < 4; ++B) acc[A][B] = 0.0;
  for (int cy = 0; cy < sy; ++cy) {
    for (int cx = 0; cx < sx; ++cx) {
      const int v = 3 * (sy == 2 ? cy : 2) + (sx == 2 ? cx : 2);
      const double *base = fcell + (size_t)(2 * J + cy) * nxf + (2 * I + cx);
      double M[4][4], T[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) M[a][b] = __ldg(base + (size_t)((2 * a + d) * 8 + 2 * b + e) * ncf);
      double W[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int A = 0; A < 4; ++A) W[a][A] = c_pw.w[v][a][A];
      // T = M W ; acc += W^T T
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int B = 0; B < 4; ++B) {
          double s = 0.0;
#pragma unroll
          for (int b = 0; b < 4; ++b) s = fma(M[a][b], W[b][B], s);
          T[a][B] = s;
        }
#pragma unroll
      for (int A = 0; A < 4; ++A)
#pragma unroll
        for (int B = 0; B < 4; ++B) {
          double s = acc[A][B];
#pragma unroll
          for (int a = 0; a < 4; ++a) s = fma(W[a][A], T[a][B], s);
          acc[A][B] = s;
        }
    }
  }
#pragma unroll
  for (int A = 0; A < 4; ++A)
#pragma unroll
    for (int B = 0; B < 4; ++B) ccell[(size_t)((2 * A + d) * 8 + 2 * B + e) * ncc + t] = acc[A][B];
}

// ---- cells -> node stencils ---------------------------------------------------------------------------
// stencil plane index: 4*nb + 2*i + j with nb = 3*(dr+1) + (dc+1)
// (all levels of a set-up in ONE launch, blockIdx.y = entry: the levels are independent of each other once their cells
// exist, and seven launches of 11 us each were latency, not work)
struct ToStencilArgs {
  const double *cell[TOPO_MAX_LEVELS];
  stencil_t *st[TOPO_MAX_LEVELS];
  double2 *dinv[TOPO_MAX_LEVELS];
  int nx[TOPO_MAX_LEVELS], ny[TOPO_MAX_LEVELS];
};
__global__ void __launch_bounds__(128) k_cells_to_stencil(const __grid_constant__ ToStencilArgs args) {
  const int lv = blockIdx.y;
  const double *__restrict__ cell = args.cell[lv];
  stencil_t *__restrict__ st = args.st[lv];
  double2 *__restrict__ dinv = args.dinv[lv];
  const int nx = args.nx[lv], ny = args.ny[lv];
  const int npr = nx + 1;
  const int64_t nn = (int64_t)npr * (ny + 1), ncell = (int64_t)nx * ny;
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nn) return;
  const int r = (int)(n / npr), c = (int)(n - (int64_t)r * npr);
  double s[36];
#pragma unroll
  for (int k = 0; k < 36; ++k) s[k] = 0.0;
  // adjacent cells: (cell row offset, cell col offset, local index of this node in that cell)
  const int cro[4] = {0, 0, -1, -1}, cco[4] = {0, -1, -1, 0}, me[4] = {0, 1, 2, 3};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int cr = r + cro[q], cc = c + cco[q];
    if (cr < 0 || cr >= ny || cc < 0 || cc >= nx) continue;
    const size_t ci = (size_t)cr * nx + cc;
    const int a = me[q];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      // node b of that cell sits at (cr + ly(b), cc + lx(b)); offset from (r, c):
      const int dr = cr + ly(b) - r, dc = cc + lx(b) - c;
      const int nb = 3 * (dr + 1) + (dc + 1);
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
          s[4 * nb + 2 * i + j] += __ldg(cell + (size_t)((2 * a + i) * 8 + 2 * b + j) * ncell + ci);
    }
  }
#pragma unroll
  // layout: 9 planes of float4 (one 2x2 block per neighbour), plane nb, node n
#pragma unroll
  for (int nb = 0; nb < 9; ++nb)
    reinterpret_cast<float4 *>(st)[(size_t)nb * nn + n] =
        make_float4((float)s[4 * nb], (float)s[4 * nb + 1], (float)s[4 * nb + 2], (float)s[4 * nb + 3]);
  const double dx = (double)(stencil_t)s[4 * 4 + 0], dy = (double)(stencil_t)s[4 * 4 + 3];
  dinv[n] = make_double2(dx > 0.0 ? 1.0 / dx : 0.0, dy > 0.0 ? 1.0 / dy : 0.0);
}

// ---- coarsest level: dense inverse --------------------------------------------------------------------------------
// One block of 96 x 8 threads.  Builds the dense operator from the node stencil, replaces dead rows/cols (zero
// diagonal) by identity and inverts IN PLACE by Gauss-Jordan (SPD => no pivoting): thread (j, g) owns column j of
// the rows g, g+8, ...; per elimination step it reads its <= 12 entries and the saved pivot column into registers,
// then writes -- loads before stores, so the shared-memory stores do not serialise the next loads (the first
// version walked a whole column per thread, 2.7 us per step).
constexpr int CI_MAXN = 96, CI_RG = 8, CI_ROWS = CI_MAXN / CI_RG;
__global__ void __launch_bounds__(CI_MAXN * CI_RG) k_coarse_invert(const stencil_t *__restrict__ st, int nx, int ny,
                                                                  double *__restrict__ inv) {
  extern __shared__ double A[];                 // N x N
  __shared__ double colk[CI_MAXN], rowk[CI_MAXN], diag0[CI_MAXN];
  const int npr = nx + 1, nn = npr * (ny + 1), N = 2 * nn;
  const int tid = threadIdx.x, nthr = blockDim.x;
  for (int t = tid; t < N * N; t += nthr) A[t] = 0.0;
  __syncthreads();
  for (int t = tid; t < nn * 9; t += nthr) {
    const int n = t / 9, nb = t % 9;
    const int r = n / npr, c = n % npr;
    const int rr = r + nb / 3 - 1, cc = c + nb % 3 - 1;
    if (rr < 0 || rr > ny || cc < 0 || cc > nx) continue;
    const int m = rr * npr + cc;
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b) A[(2 * n + a) * N + 2 * m + b] = st[((size_t)nb * nn + n) * 4 + 2 * a + b];
  }
  __syncthreads();
  for (int t = tid; t < N; t += nthr) {
    if (!(A[t * N + t] > 0.0)) {
      for (int q = 0; q < N; ++q) { A[t * N + q] = 0.0; A[q * N + t] = 0.0; }
    }
  }
  __syncthreads();
  for (int t = tid; t < N; t += nthr) {
    if (!(A[t * N + t] > 0.0)) A[t * N + t] = 1.0;
    diag0[t] = A[t * N + t];
  }
  __syncthreads();
  const int j = tid % CI_MAXN, g = tid / CI_MAXN;
  for (int k = 0; k < N; ++k) {
    // Perforated designs (ESO: elements removed, orphaned nodes pinned) can leave the coarsest operator singular
    // -- floating islands, hinges -- without any zero diagonal.  A direction whose pivot has dropped below the
    // accuracy of the fp32 stencil the matrix was read from is removed from the inverse (row and column zeroed)
    // instead of being amplified by 1/pivot; a consistent right-hand side has no component along it.
    const double piv = A[k * N + k];
    const double pinv = (piv > 1e-6 * diag0[k]) ? 1.0 / piv : 0.0;
    if (tid < N) {
      colk[tid] = A[tid * N + k];
      rowk[tid] = (tid == k) ? pinv : A[k * N + tid] * pinv;
    }
    __syncthreads();
    if (j < N) {
      const double rk = rowk[j];
      double a[CI_ROWS], c[CI_ROWS];
#pragma unroll
      for (int q = 0; q < CI_ROWS; ++q) {
        const int i = min(g + CI_RG * q, N - 1);
        a[q] = A[i * N + j];
        c[q] = colk[i];
      }
#pragma unroll
      for (int q = 0; q < CI_ROWS; ++q) {
        const int i = g + CI_RG * q;
        if (i < N) {
          double v;
          if (i == k) v = rk;                         // scaled pivot row (1/p on the diagonal)
          else if (j == k) v = -c[q] * pinv;          // pivot column
          else v = fma(-c[q], rk, a[q]);
          A[i * N + j] = v;
        }
      }
    }
    __syncthreads();
  }
  for (int t = tid; t < N * N; t += nthr) inv[t] = A[t];
}

inline int blocks_for(int64_t n, int threads = 128) { return (int)((n + threads - 1) / threads); }

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
int topo_gmg_plan(topo_handle *h) {
  h->n_levels = 1;
  h->lvl_nx[0] = h->nx;
  h->lvl_ny[0] = h->ny;
  int nx = h->nx, ny = h->ny;
  // coarsen (ceil halving) until a level has at most COARSEST_MAX_NODES nodes: that level is solved exactly with
  // a dense inverse (<= 96 DOFs), which is both cheaper and more accurate than smoothing three more tiny levels.
  // Row-slab mode: levels < K halve the slab (its height is a multiple of 2^K), level K and deeper have the
  // dimensions of the GLOBAL hierarchy (they are replicated on every rank).
  const int coarsest_max_nodes = h->knobs.coarsest_nodes;
  const bool slab = h->dist.on;
  while (h->n_levels < TOPO_MAX_LEVELS) {
    const int l = h->n_levels;
    if (!(slab && l <= h->dist.K) && !((nx + 1) * (ny + 1) > coarsest_max_nodes && std::max(nx, ny) > 1)) break;
    nx = (nx + 1) / 2;
    if (slab && l < h->dist.K) ny = ny / 2;
    else if (slab && l == h->dist.K) ny = h->dist.ny_global >> h->dist.K;
    else ny = (ny + 1) / 2;
    h->lvl_nx[h->n_levels] = nx;
    h->lvl_ny[h->n_levels] = ny;
    h->n_levels++;
  }
  h->levels.assign(h->n_levels, GmgLevel{});
  for (int l = 1; l < h->n_levels; ++l) {
    GmgLevel &L = h->levels[l];
    L.nx = h->lvl_nx[l];
    L.ny = h->lvl_ny[l];
    L.nn = (L.nx + 1) * (L.ny + 1);
    L.ncell = L.nx * L.ny;
    // level-1 cells are never stored (k_stencil_level1 / k_cells_level2 work from the element scales)
    // (row-slab mode with K == 1: level 1 is the first global level and its gathered cells live here)
    if (l >= 2 || (h->dist.on && h->dist.K == 1)) CUDA_TRY(cudaMalloc(&L.cell, sizeof(double) * 64 * (size_t)L.ncell));
    if (l == 1) CUDA_TRY(cudaMalloc(&L.interior, (size_t)L.nn));          // per level-1 node
    if (l == 2) CUDA_TRY(cudaMalloc(&L.interior, (size_t)L.ncell));       // per level-2 cell
  }
  // Per-cycle data of the levels: level 1 gets its own allocation; levels >= 2 (and the coarsest dense
  // inverse) share ONE slab so that a single L2 persisting access-policy window can cover them -- the
  // fine grid streams ~1 GB per PCG iteration through the 126 MB L2 and would otherwise evict them,
  // leaving the small, latency-bound kernels to pay DRAM latency on every load.
  auto level_bytes = [](const GmgLevel &L) {
    return (size_t)L.nn * (sizeof(stencil_t) * 36 + sizeof(double2) * 5);
  };
  auto carve = [](GmgLevel &L, char *base) {
    size_t off = 0;
    L.st = reinterpret_cast<stencil_t *>(base + off); off += sizeof(stencil_t) * 36 * (size_t)L.nn;
    L.dinv = reinterpret_cast<double2 *>(base + off); off += sizeof(double2) * (size_t)L.nn;
    L.x = reinterpret_cast<double2 *>(base + off); off += sizeof(double2) * (size_t)L.nn;
    L.x2 = reinterpret_cast<double2 *>(base + off); off += sizeof(double2) * (size_t)L.nn;
    L.b = reinterpret_cast<double2 *>(base + off); off += sizeof(double2) * (size_t)L.nn;
    L.r = reinterpret_cast<double2 *>(base + off);
  };
  h->lvl1_slab = nullptr;
  h->coarse_slab = nullptr;
  h->coarse_slab_bytes = 0;
  h->coarse_n = 0;
  h->coarse_inv = nullptr;
  if (h->n_levels > 1) {
    char *p1 = nullptr;
    CUDA_TRY(cudaMalloc(&p1, level_bytes(h->levels[1])));
    h->lvl1_slab = p1;
    carve(h->levels[1], p1);
    h->coarse_n = 2 * h->levels[h->n_levels - 1].nn;
    size_t tot = 256 + sizeof(double) * (size_t)h->coarse_n * h->coarse_n;
    for (int l = 2; l < h->n_levels; ++l) tot += (level_bytes(h->levels[l]) + 255) & ~(size_t)255;
    char *p = nullptr;
    CUDA_TRY(cudaMalloc(&p, tot));
    h->coarse_slab = p;
    h->coarse_slab_bytes = tot;
    size_t off = 0, persist_off = (size_t)-1;
    for (int l = 2; l < h->n_levels; ++l) {
      carve(h->levels[l], p + off);
      // L2 residency (below) only for the small, latency-bound levels at the end of the slab
      if (persist_off == (size_t)-1 && h->levels[l].nn <= 40000) persist_off = off;
      off += (level_bytes(h->levels[l]) + 255) & ~(size_t)255;
    }
    if (persist_off == (size_t)-1) persist_off = off;
    h->coarse_inv = reinterpret_cast<double *>(p + off);
    if (h->n_levels == 2) {   // the only coarse level is level 1: its data is not in the slab
    }
    // L2 residency for the levels of <= 40 000 nodes + the dense inverse (best effort: ignored if the device refuses).
    // The set-aside must stay SMALL: it is taken from the L2 of every other kernel of the process, and with the
    // whole slab in the window (tens of MB on 2048x1024, the device maximum on 8192x4096) the fine-grid kernels
    // lost half of their throughput beyond the L2 capacity (plain product on 8192x4096: 660 us with, 293 us without).
    int dev_max_persist = 0, dev_max_window = 0;
    cudaDeviceGetAttribute(&dev_max_persist, cudaDevAttrMaxPersistingL2CacheSize, h->device);
    cudaDeviceGetAttribute(&dev_max_window, cudaDevAttrMaxAccessPolicyWindowSize, h->device);
    const size_t win_bytes = tot - persist_off;
    const size_t want = std::min<size_t>(std::min<size_t>(win_bytes, (size_t)16 << 20), (size_t)std::max(dev_max_persist, 0));
    if (want > 0 && dev_max_window > 0 && !h->knobs.no_l2persist) {
      cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
      cudaStreamAttrValue attr = {};
      attr.accessPolicyWindow.base_ptr = p + persist_off;
      attr.accessPolicyWindow.num_bytes = std::min<size_t>(win_bytes, (size_t)dev_max_window);
      attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)want / (double)attr.accessPolicyWindow.num_bytes);
      attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
      cudaGetLastError();
    }
  }
  // constant tables (per-process; identical for every handle except G, which depends on ke)
  PwTable pw;
  for (int ty = 0; ty < 3; ++ty)
    for (int tx = 0; tx < 3; ++tx)
      for (int a = 0; a < 4; ++a)
        for (int A = 0; A < 4; ++A) pw.w[3 * ty + tx][a][A] = l1d(tx, lx(a), lx(A)) * l1d(ty, ly(a), ly(A));
  CUDA_TRY(cudaMemcpyToSymbol(c_pw, &pw, sizeof(pw)));
  return TOPO_OK;
}

static int upload_gtable(topo_handle *h) {
  // G_v = P_v^T ke P_v (host, once per setup: ke is per handle but the symbol is per process)
  GTable g;
  for (int v = 0; v < 9; ++v) {
    const int tx = v % 3, ty = v / 3;
    double P[8][8];
    std::memset(P, 0, sizeof(P));
    for (int a = 0; a < 4; ++a)
      for (int A = 0; A < 4; ++A) {
        const double w = l1d(tx, lx(a), lx(A)) * l1d(ty, ly(a), ly(A));
        P[2 * a][2 * A] = w;
        P[2 * a + 1][2 * A + 1] = w;
      }
    for (int i = 0; i < 8; ++i)
      for (int j = 0; j < 8; ++j) {
        double s = 0.0;
        for (int a = 0; a < 8; ++a)
          for (int b = 0; b < 8; ++b) s += P[a][i] * h->ke.k[a * 8 + b] * P[b][j];
        g.g[v][i * 8 + j] = s;
      }
  }
  CUDA_TRY(cudaMemcpyToSymbolAsync(c_g, &g, sizeof(g), 0, cudaMemcpyHostToDevice, h->stream));
  return TOPO_OK;
}

static int invert_coarsest(topo_handle *h) {
  GmgLevel &L = h->levels[h->n_levels - 1];
  const int N = h->coarse_n;
  if (N > 96) {
    topo_set_error("coarsest multigrid level too large for the dense inverse (N=%d)", N);
    return TOPO_EINVAL;
  }
  static bool attr_done_dev[TOPO_MAX_DEVICES] = {};
  bool &attr_done = attr_done_dev[topo_dev_slot(h)];
  if (!attr_done) {
    CUDA_TRY(cudaFuncSetAttribute(k_coarse_invert, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * CI_MAXN * CI_MAXN)));
    attr_done = true;
  }
  // Single-GPU path outside profile mode: the inversion runs on the side stream, forked here (the coarsest stencil
  // is in place) and joined by the first tail launch of the solve (launch_tail) -- everything in between (the other
  // levels' stencils, residual, first fine-grid pass, levels 1..3 going down) does not need the inverse.
  const bool overlap = h->inv_overlap && !h->prof_on && !h->dist.on;
  cudaStream_t st = h->stream;
  if (overlap) {
    CUDA_TRY(cudaEventRecord(h->ev_fork, h->stream));
    CUDA_TRY(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
    st = h->side;
  }
  KL(h, "k_coarse_invert", k_coarse_invert<<<1, CI_MAXN * CI_RG, sizeof(double) * N * N, st>>>(L.st, L.nx, L.ny, h->coarse_inv));
  CUDA_TRY(cudaGetLastError());
  if (overlap) {
    CUDA_TRY(cudaEventRecord(h->ev_inv, h->side));
    h->inv_pending = true;
  }
  return TOPO_OK;
}

// level-1 stencil / level-2 cells straight from the element scales of an nx0 x ny0 mesh (the whole mesh, or a slab):
// interior flags first, then the branch-free instantiations for the marked nodes / cells and the general ones for the rest
// node stencils + D^-1 of levels [l0, l1) from their cells, one launch
static int launch_to_stencil(topo_handle *h, int l0, int l1) {
  if (l1 <= l0) return TOPO_OK;
  ToStencilArgs a = {};
  int64_t nmax = 0;
  for (int l = l0; l < l1; ++l) {
    const GmgLevel &L = h->levels[l];
    a.cell[l - l0] = L.cell; a.st[l - l0] = L.st; a.dinv[l - l0] = L.dinv; a.nx[l - l0] = L.nx; a.ny[l - l0] = L.ny;
    nmax = std::max<int64_t>(nmax, L.nn);
  }
  KL(h, "k_cells_to_stencil", k_cells_to_stencil<<<dim3(blocks_for(nmax), l1 - l0), 128, 0, h->stream>>>(a));
  CUDA_TRY(cudaGetLastError());
  return TOPO_OK;
}

// Order: flags | side stream: level-2 edge cells, level-1 edge nodes | main stream: level-2 cells, level-1 stencil, then
// the join with the level-2 edge cells (the coarsening chain that follows reads all level-2 cells).  The level-1 edge
// nodes are joined by join_level1_edge at the end of the set-up.  Profile mode (event brackets on the main stream):
// everything on the main stream.
static int launch_level12(topo_handle *h, int nx1, int ny1, double *cell2, int nx2, int ny2) {
  cudaStream_t st = h->stream;
  GmgLevel &L1 = h->levels[1];
  uint8_t *f1 = L1.interior, *f2 = (cell2 != nullptr) ? h->levels[2].interior : nullptr;
  const int64_t nn1 = (int64_t)(nx1 + 1) * (ny1 + 1), nc2 = (int64_t)nx2 * ny2;
  KL(h, "k_interior_flags", k_interior_flags<<<blocks_for(std::max<int64_t>(nn1, cell2 ? nc2 : 0), 256), 256, 0, st>>>(h->fixed, h->nx, h->ny, nx1, ny1, nx2, ny2, f1, f2, h->knobs.setup_fast));
  const bool side = !h->prof_on;
  cudaStream_t se = side ? h->side : st;
  if (side) {
    CUDA_TRY(cudaEventRecord(h->ev_fork, st));
    CUDA_TRY(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
  }
  const dim3 g1(blocks_for(nc2), 1), g4(blocks_for(nc2), 4);
#define TOPO_CL2(FAST, DE, GRID, STREAM) k_cells_level2<FAST, DE><<<GRID, 128, 0, STREAM>>>(h->ke, h->scale, h->fixed, cell2, h->nx, h->ny, nx1, ny1, nx2, ny2, f2)
#define TOPO_ST1(FAST, STREAM) k_stencil_level1<FAST><<<blocks_for(nn1), 128, 0, STREAM>>>(h->ke, h->scale, h->fixed, L1.st, L1.dinv, h->nx, h->ny, nx1, ny1, f1)
  if (cell2 != nullptr) KL(h, "k_cells_level2 edge", TOPO_CL2(false, 0, g4, se));
  if (side) CUDA_TRY(cudaEventRecord(h->ev_edge[0], h->side));
  KL(h, "k_stencil_level1 edge", TOPO_ST1(false, se));
  if (side) {
    CUDA_TRY(cudaEventRecord(h->ev_edge[1], h->side));
    h->edge_pending = true;
  }
  if (cell2 != nullptr)
    KL(h, "k_cells_level2", (TOPO_CL2(true, 0, g1, st), TOPO_CL2(true, 1, g1, st), TOPO_CL2(true, 2, g1, st), TOPO_CL2(true, 3, g1, st)));
  KL(h, "k_stencil_level1", TOPO_ST1(true, st));
#undef TOPO_CL2
#undef TOPO_ST1
  if (side) CUDA_TRY(cudaStreamWaitEvent(st, h->ev_edge[0], 0));
  CUDA_TRY(cudaGetLastError());
  return TOPO_OK;
}
static int join_level1_edge(topo_handle *h) {
  if (h->edge_pending) {
    CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_edge[1], 0));
    h->edge_pending = false;
  }
  return TOPO_OK;
}

int topo_gmg_setup(topo_handle *h) {
  if (h->n_levels < 2) return TOPO_OK;
  if (h->dist.on) {
    topo_set_error("row-slab handle: use topo_dist_setup_local / topo_dist_setup_global");
    return TOPO_ESTATE;
  }
  cudaStream_t st = h->stream;
  TOPO_TRY(upload_gtable(h));
  {
    GmgLevel &F = h->levels[1];
    if (h->n_levels > 2) TOPO_TRY(launch_level12(h, F.nx, F.ny, h->levels[2].cell, h->levels[2].nx, h->levels[2].ny));
    else TOPO_TRY(launch_level12(h, F.nx, F.ny, nullptr, 0, 0));
  }
  for (int l = 3; l < h->n_levels; ++l) {
    GmgLevel &F = h->levels[l - 1], &C = h->levels[l];
    dim3 grid(blocks_for(C.ncell), 4);
    KL(h, "k_cells_coarsen", k_cells_coarsen<<<grid, 128, 0, st>>>(F.cell, C.cell, F.nx, F.ny, C.nx, C.ny));
  }
  TOPO_TRY(launch_to_stencil(h, 2, h->n_levels));
  if (h->n_levels > 2) TOPO_TRY(invert_coarsest(h));       // (side stream: overlaps the start of the solve)
  if (h->n_levels == 2) {                                  // level 1 is the coarsest: its stencil was written above
    TOPO_TRY(join_level1_edge(h));
    TOPO_TRY(invert_coarsest(h));
  }
  TOPO_TRY(join_level1_edge(h));
  h->gmg_valid = true;
  return TOPO_OK;
}

// ---- row-slab mode ------------------------------------------------------------------------------------------
namespace {
// raw diagonal (1/dinv, 0 where dinv == 0) of node rows 0 and ny of one level -> packed send buffers
__global__ void k_dist_diag_pack(const double2 *__restrict__ dinv, int npr, int ny, double2 *__restrict__ lo,
                                 double2 *__restrict__ hi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npr) return;
  const double2 a = dinv[i], b = dinv[(size_t)ny * npr + i];
  lo[i] = make_double2(a.x != 0.0 ? 1.0 / a.x : 0.0, a.y != 0.0 ? 1.0 / a.y : 0.0);
  hi[i] = make_double2(b.x != 0.0 ? 1.0 / b.x : 0.0, b.y != 0.0 ? 1.0 / b.y : 0.0);
}
// dinv = 1 / (own diagonal + the neighbour's) on one interface row
__global__ void k_dist_diag_unpack(double2 *__restrict__ dinv_row, float2 *__restrict__ dinv_f_row,
                                   const double2 *__restrict__ mine, const double2 *__restrict__ theirs, int npr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npr) return;
  const double dx = mine[i].x + theirs[i].x, dy = mine[i].y + theirs[i].y;
  const double2 v = make_double2(mine[i].x > 0.0 && dx > 0.0 ? 1.0 / dx : 0.0, mine[i].y > 0.0 && dy > 0.0 ? 1.0 / dy : 0.0);
  dinv_row[i] = v;
  if (dinv_f_row != nullptr) dinv_f_row[i] = make_float2((float)v.x, (float)v.y);     // fine grid: fp32 copy
}
// gathered level-K cells [world][64][ncl] -> global SoA planes [64][world*ncl]
__global__ void k_dist_cells_reorder(const double *__restrict__ recv, double *__restrict__ cell, int64_t ncl, int world) {
  const int64_t tot = 64 * ncl * world;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = t % ncl, q = t / ncl;
    const int plane = (int)(q % 64), g = (int)(q / 64);
    cell[(int64_t)plane * (ncl * world) + (int64_t)g * ncl + c] = recv[t];
  }
}
}  // namespace

// local part of the set-up: fine diagonal, Galerkin cells of the slab down to level K (the level-K cells go to the
// gather send buffer), node stencils of the local levels, interface diagonals packed for the exchange
int topo_dist_setup_local(topo_handle *h) {
  DistState &D = h->dist;
  if (!D.on) return TOPO_ESTATE;
  cudaStream_t st = h->stream;
  TOPO_TRY(topo_build_diag(h));
  TOPO_TRY(upload_gtable(h));
  const int K = D.K;
  auto local_cells = [&](int l) -> double * { return l == K ? D.cellK_send : h->levels[l].cell; };
  auto local_ny = [&](int l) { return h->ny >> l; };
  if (K == 1) {                                   // the level-1 cells ARE the gathered block
    const int nxc = h->lvl_nx[1], nyc = local_ny(1);
    KL(h, "k_cells_level1", k_cells_level1<<<blocks_for((int64_t)nxc * nyc), 128, 0, st>>>(h->ke, h->scale, h->fixed, local_cells(1), h->nx, h->ny,
                                                                                            nxc, nyc));
  } else {
    TOPO_TRY(launch_level12(h, h->lvl_nx[1], local_ny(1), local_cells(2), h->lvl_nx[2], local_ny(2)));
  }
  for (int l = 3; l <= K; ++l) {
    const int nxf = h->lvl_nx[l - 1], nyf = local_ny(l - 1), nxc = h->lvl_nx[l], nyc = local_ny(l);
    dim3 grid(blocks_for((int64_t)nxc * nyc), 4);
    KL(h, "k_cells_coarsen", k_cells_coarsen<<<grid, 128, 0, st>>>(local_cells(l - 1), local_cells(l), nxf, nyf, nxc, nyc));
  }
  TOPO_TRY(launch_to_stencil(h, 2, K));
  TOPO_TRY(join_level1_edge(h));
  long long off = 0;
  for (int l = 0; l < K; ++l) {
    const int npr = h->lvl_nx[l] + 1, ny = local_ny(l);
    const double2 *dinv = (l == 0) ? h->dinv : h->levels[l].dinv;
    KL(h, "k_dist_diag_pack", k_dist_diag_pack<<<blocks_for(npr), 128, 0, st>>>(dinv, npr, ny, D.diag_send_lo + off, D.diag_send_hi + off));
    off += npr;
  }
  CUDA_TRY(cudaGetLastError());
  return TOPO_OK;
}

// after the host exchanged the interface diagonals and gathered the level-K cells
int topo_dist_setup_global(topo_handle *h) {
  DistState &D = h->dist;
  if (!D.on) return TOPO_ESTATE;
  cudaStream_t st = h->stream;
  const int K = D.K;
  long long off = 0;
  for (int l = 0; l < K; ++l) {
    const int npr = h->lvl_nx[l] + 1, ny = h->ny >> l;
    double2 *dinv = (l == 0) ? h->dinv : h->levels[l].dinv;
    float2 *dinv_f = (l == 0) ? h->dinv_f : nullptr;
    const size_t last = (size_t)ny * npr;
    if (D.has_lo)
      KL(h, "k_dist_diag_unpack", k_dist_diag_unpack<<<blocks_for(npr), 128, 0, st>>>(dinv, dinv_f, D.diag_send_lo + off, D.diag_recv_lo + off,
                                                                                       npr));
    if (D.has_hi)
      KL(h, "k_dist_diag_unpack", k_dist_diag_unpack<<<blocks_for(npr), 128, 0, st>>>(dinv + last, dinv_f ? dinv_f + last : nullptr,
                                                                                       D.diag_send_hi + off, D.diag_recv_hi + off, npr));
    off += npr;
  }
  {
    GmgLevel &G = h->levels[K];
    KL(h, "k_dist_cells_reorder", k_dist_cells_reorder<<<std::min(blocks_for(64 * D.ncellK_loc * D.world, 256), 148 * 8), 256, 0, st>>>(
           D.cellK_recv, G.cell, D.ncellK_loc, D.world));
  }
  for (int l = K + 1; l < h->n_levels; ++l) {
    GmgLevel &F = h->levels[l - 1], &C = h->levels[l];
    dim3 grid(blocks_for(C.ncell), 4);
    KL(h, "k_cells_coarsen", k_cells_coarsen<<<grid, 128, 0, st>>>(F.cell, C.cell, F.nx, F.ny, C.nx, C.ny));
  }
  TOPO_TRY(launch_to_stencil(h, K, h->n_levels));
  TOPO_TRY(invert_coarsest(h));
  h->gmg_valid = true;
  return TOPO_OK;
}

void topo_tail2_forget(topo_handle *h);

void topo_gmg_free(topo_handle *h) {
  topo_tail2_forget(h);
  for (size_t l = 1; l < h->levels.size(); ++l) {
    cudaFree(h->levels[l].cell);
    cudaFree(h->levels[l].interior);
  }
  h->levels.clear();
  cudaFree(h->lvl1_slab);
  cudaFree(h->coarse_slab);
  h->lvl1_slab = h->coarse_slab = nullptr;
  h->n_levels = 1;
  h->gmg_valid = false;
}

extern "C" int topo_gmg_info(topo_t *h, int *n_levels, int *nx_levels, int *ny_levels) {
  if (!h || !n_levels) return TOPO_EINVAL;
  *n_levels = h->n_levels;
  for (int l = 0; l < h->n_levels; ++l) {
    if (nx_levels) nx_levels[l] = h->lvl_nx[l];
    if (ny_levels) ny_levels[l] = h->lvl_ny[l];
  }
  return TOPO_OK;
}

extern "C" int topo_set_gmg(topo_t *h, double omega, int nu_pre, int nu_post) {
  if (!h || !(omega > 0.0) || nu_pre < 1 || nu_post < 1) {
    topo_set_error("topo_set_gmg: omega > 0 and nu_pre, nu_post >= 1 required");
    return TOPO_EINVAL;
  }
  h->omega = omega;
  h->nu_pre = nu_pre;
  h->nu_post = nu_post;
  h->pcg_graph_valid = false;
  return TOPO_OK;
}
