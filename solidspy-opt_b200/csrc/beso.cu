// Stages 3+4 of the BESO / ESO iterations: strain-energy sensitivity numbers, the BESO
// element->node->element smoothing, device radix select of the removal threshold, strict-threshold
// masks with load/support protection, and orphan-node pinning.
// Reference: optimize.py:86-97,188-196,304-331; solver.py:6-58,90-129,177-362.
//
// Arithmetic whose ROUNDING decides a strict comparison later on (the weighted averages feeding
// `sens > alpha`) is written with explicit non-fused __dmul_rn/__dadd_rn in the reference's operation
// order, so that exactly tied values stay exactly tied.
#include <algorithm>
#include <cmath>

#include "topo_internal.cuh"

int topo_mask_vector(topo_handle *h, double2 *v);

namespace {

constexpr int ET = 256;
inline int el_blocks(int64_t n) { return (int)std::min<int64_t>(div_up(n, ET), 148 * 8); }

// a_e = 0.5 u_e^T (ke u_e) * keep_e ; reduces max
__global__ void __launch_bounds__(ET) k_energy(const __grid_constant__ KeParam ke, const double2 *__restrict__ u,
                                               const uint8_t *__restrict__ keep, double *__restrict__ a, int nx,
                                               int ny, double *partials, unsigned int *ticket, double *out) {
  const int64_t ne = (int64_t)nx * ny;
  const int npr = nx + 1;
  double v[1] = {-1.0 / 0.0};
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x) {
    double val = 0.0;
    if (keep[e]) {
      const int r = (int)(e / nx), c = (int)(e - (int64_t)r * nx);
      const double2 *p = u + (size_t)r * npr + c;
      const double2 q0 = __ldg(p), q1 = __ldg(p + 1), q2 = __ldg(p + npr + 1), q3 = __ldg(p + npr);
      const double w[8] = {q0.x, q0.y, q1.x, q1.y, q2.x, q2.y, q3.x, q3.y};
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {        // u^T (ke u): row dot products first (solver.py:124)
        double t = 0.0;
#pragma unroll
        for (int j = 0; j < 8; ++j) t = fma(ke.k[i * 8 + j], w[j], t);
        s = fma(w[i], t, s);
      }
      val = 0.5 * s;
      v[0] = fmax(v[0], val);
    }
    a[e] = val;
  }
  grid_reduce<1, true>(v, partials, ticket, out);
}

struct BesoW {
  double w4[4], w2x[2], w2y[2], we[4], wesum;
};

__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }

// sensitivity_nodes (solver.py:177-210): node <- adjacent elements (ascending id), input a/amax
__global__ void __launch_bounds__(ET) k_beso_nodes(const __grid_constant__ BesoW W, const double *__restrict__ a,
                                                   const double *__restrict__ amax, double *__restrict__ nodal,
                                                   int nx, int ny) {
  const int npr = nx + 1;
  const int64_t nn = (int64_t)npr * (ny + 1);
  const double mx = *amax;
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < nn; n += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(n / npr), c = (int)(n - (int64_t)r * npr);
    const bool hasS = r > 0, hasN = r < ny, hasW = c > 0, hasE = c < nx;
    auto el = [&](int er, int ec) { return a[(size_t)er * nx + ec] / mx; };
    double out;
    if (hasS && hasN && hasW && hasE) {
      out = add(add(add(mul(W.w4[0], el(r - 1, c - 1)), mul(W.w4[1], el(r - 1, c))), mul(W.w4[2], el(r, c - 1))),
                mul(W.w4[3], el(r, c)));
    } else if ((hasS != hasN) && hasW && hasE) {          // bottom / top edge: two elements side by side
      const int er = hasN ? r : r - 1;
      out = add(mul(W.w2x[0], el(er, c - 1)), mul(W.w2x[1], el(er, c)));
    } else if ((hasW != hasE) && hasS && hasN) {          // left / right edge: two elements stacked
      const int ec = hasE ? c : c - 1;
      out = add(mul(W.w2y[0], el(r - 1, ec)), mul(W.w2y[1], el(r, ec)));
    } else {                                              // corner: single element, plain copy
      out = el(hasN ? r : r - 1, hasE ? c : c - 1);
    }
    nodal[n] = out;
  }
}

// sensitivity_filter (solver.py:212-243) with r_min selecting the 4 corner nodes (ascending node id)
__global__ void __launch_bounds__(ET) k_beso_els(const __grid_constant__ BesoW W, const double *__restrict__ nodal,
                                                 double *__restrict__ s, int nx, int ny, double *partials,
                                                 unsigned int *ticket, double *out) {
  const int64_t ne = (int64_t)nx * ny;
  const int npr = nx + 1;
  double v[1] = {-1.0 / 0.0};
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(e / nx), c = (int)(e - (int64_t)r * nx);
    const double *p = nodal + (size_t)r * npr + c;
    const double num = add(add(add(mul(W.we[0], p[0]), mul(W.we[1], p[1])), mul(W.we[2], p[npr])), mul(W.we[3], p[npr + 1]));
    const double val = num / W.wesum;
    s[e] = val;
    v[0] = fmax(v[0], val);
  }
  grid_reduce<1, true>(v, partials, ticket, out);
}

// sensitivity_filter (solver.py:212-243) for ANY radius: the nodes within r_min of an element centre are a fixed set of
// offsets on the uniform grid, cut off at the mesh boundary.  Per element, over the offsets that exist:
//   m = count, R = sum r, w = (1 - r/R)/(m - 1), out = sum(w s)/sum(w)        (the reference's expressions, :236-238)
constexpr int BR_MAX = 9;              // node offsets -BR_MAX+1 .. BR_MAX around the element's lower-left node
struct BesoTaps {
  int lo, hi;                          // offsets lo..hi in both directions
  double r[2 * BR_MAX][2 * BR_MAX];    // distance node (dj, di) -> element centre; < 0: outside the radius
};
__global__ void __launch_bounds__(ET) k_beso_els_radius(const __grid_constant__ BesoTaps T, const double *__restrict__ nodal,
                                                        double *__restrict__ s, int nx, int ny, double *partials,
                                                        unsigned int *ticket, double *out) {
  const int64_t ne = (int64_t)nx * ny;
  const int npr = nx + 1;
  double v[1] = {-1.0 / 0.0};
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(e / nx), c = (int)(e - (int64_t)r * nx);
    int m = 0;
    double R = 0.0;
    for (int dj = T.lo; dj <= T.hi; ++dj) {
      if (r + dj < 0 || r + dj > ny) continue;
      for (int di = T.lo; di <= T.hi; ++di) {
        const double d = T.r[dj - T.lo][di - T.lo];
        if (d < 0.0 || c + di < 0 || c + di > nx) continue;
        ++m;
        R += d;
      }
    }
    const double f = 1.0 / (double)(m - 1);
    double num = 0.0, den = 0.0;
    for (int dj = T.lo; dj <= T.hi; ++dj) {
      if (r + dj < 0 || r + dj > ny) continue;
      for (int di = T.lo; di <= T.hi; ++di) {
        const double d = T.r[dj - T.lo][di - T.lo];
        if (d < 0.0 || c + di < 0 || c + di > nx) continue;
        const double w = f * (1.0 - d / R);
        num = fma(w, nodal[(size_t)(r + dj) * npr + (c + di)], num);
        den += w;
      }
    }
    const double val = num / den;
    s[e] = val;
    v[0] = fmax(v[0], val);
  }
  grid_reduce<1, true>(v, partials, ticket, out);
}

// s <- s / max1 ; if avg: s <- (s + prev)/2 ; reduce max of the result
__global__ void __launch_bounds__(ET) k_beso_hist(double *__restrict__ s, const double *__restrict__ prev,
                                                  const double *__restrict__ max1, int avg, int64_t ne,
                                                  double *partials, unsigned int *ticket, double *out) {
  const double m1 = *max1;
  double v[1] = {-1.0 / 0.0};
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x) {
    double val = s[e] / m1;
    if (avg) val = add(val, prev[e]) / 2.0;
    s[e] = val;
    v[0] = fmax(v[0], val);
  }
  grid_reduce<1, true>(v, partials, ticket, out);
}

__global__ void k_div_scalar(double *__restrict__ s, const double *__restrict__ d, int64_t ne) {
  const double m = *d;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x)
    s[e] = s[e] / m;
}

// ---- radix select (k-th largest, 0-based, duplicates counted) -----------------------------------------
// order-preserving map double -> uint64
__device__ __forceinline__ unsigned long long dkey(double x) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dunkey(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}
// sel layout: [0..255] histogram, [256] prefix, [257] remaining k, [258] pass
__global__ void k_sel_init(unsigned long long *sel, unsigned long long k) {
  for (int i = threadIdx.x; i < 256; i += blockDim.x) sel[i] = 0ull;
  if (threadIdx.x == 0) { sel[256] = 0ull; sel[257] = k; sel[258] = 0ull; }
}
// One pass = histogram of the current 8-bit digit over the keys that match the prefix found so far; the block that
// draws the last ticket then picks the digit (bins walked from the largest down, as np.sort(s)[::-1][k] counts) with
// all 256 threads -- the separate one-thread scan kernel of round 1 cost 10.9 us per pass, more than the histogram.
__global__ void __launch_bounds__(ET) k_sel_hist(const double *__restrict__ s, int64_t ne, unsigned long long *sel,
                                                 unsigned int *ticket, double *alpha_out) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned long long cnt[256];
  __shared__ bool s_last;
  for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0u;
  __syncthreads();
  const int pass = (int)sel[258];
  const int shift = 56 - 8 * pass;
  const unsigned long long prefix = sel[256];
  const unsigned long long himask = pass == 0 ? 0ull : (~0ull << (shift + 8));
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long key = dkey(s[e]);
    if ((key & himask) == prefix) atomicAdd(&hist[(key >> shift) & 0xffull], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 256; i += blockDim.x)
    if (hist[i]) atomicAdd(&sel[i], (unsigned long long)hist[i]);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int b = threadIdx.x;                       // blockDim.x == 256 == number of bins
  cnt[b] = __ldcg(&sel[b]);
  __syncthreads();
  unsigned long long above = 0ull;                 // keys in larger digits
  for (int j = b + 1; j < 256; ++j) above += cnt[j];
  const unsigned long long k = __ldcg(&sel[257]);
  const bool mine = (b >= 1) ? (k >= above && k < above + cnt[b]) : (k >= above);
  __syncthreads();
  if (mine) {
    const unsigned long long pre = prefix | (((unsigned long long)b) << shift);
    sel[256] = pre;
    sel[257] = k - above;
    sel[258] = (unsigned long long)(pass + 1);
    if (pass == 7) *alpha_out = dunkey(pre);
  }
  sel[b] = 0ull;
  if (b == 0) *ticket = 0u;
}

// ---- masks, protection, node pinning ------------------------------------------------------------------------
__global__ void k_mark_nodes(uint8_t *prot, const int64_t *ids, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) prot[ids[i]] = 1;
}

__device__ __forceinline__ bool touches(const uint8_t *prot, int r, int c, int npr) {
  const uint8_t *p = prot + (size_t)r * npr + c;
  return (p[0] | p[1] | p[npr] | p[npr + 1]) != 0;
}

// BESO: keep = (s > alpha) | (removed & touches protected node); counts kept and re-added
__global__ void __launch_bounds__(ET) k_beso_mask(const double *__restrict__ s, const double *__restrict__ alpha,
                                                  const uint8_t *__restrict__ prot, uint8_t *__restrict__ keep, int nx,
                                                  int ny, double *partials, unsigned int *ticket, double *out) {
  const int64_t ne = (int64_t)nx * ny;
  const double al = *alpha;
  double v[2] = {0.0, 0.0};
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(e / nx), c = (int)(e - (int64_t)r * nx);
    bool k = s[e] > al;
    if (!k && touches(prot, r, c, nx + 1)) {
      k = true;
      v[1] += 1.0;
    }
    keep[e] = k ? 1 : 0;
    v[0] += k ? 1.0 : 0.0;
  }
  grid_reduce<2>(v, partials, ticket, out);
}

// ESO: delete present elements with s < rr unless protected (optimize.py:189-195); deleted => scale 0
__global__ void __launch_bounds__(ET) k_eso_mask(const double *__restrict__ s, double rr,
                                                 const uint8_t *__restrict__ prot, uint8_t *__restrict__ keep,
                                                 double *__restrict__ scale, int nx, int ny, double *partials,
                                                 unsigned int *ticket, double *out) {
  const int64_t ne = (int64_t)nx * ny;
  double v[1] = {0.0};
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x) {
    if (!keep[e]) continue;
    const int r = (int)(e / nx), c = (int)(e - (int64_t)r * nx);
    if (s[e] < rr && !touches(prot, r, c, nx + 1)) {
      keep[e] = 0;
      scale[e] = 0.0;
    } else {
      v[0] += 1.0;
    }
  }
  grid_reduce<1>(v, partials, ticket, out);
}

// del_node (solver.py:35-58) / del_nodeESO (:345-362): node in no kept element -> both flags
// constrained; BESO additionally frees used, unprotected nodes.
__global__ void __launch_bounds__(ET) k_del_node(const uint8_t *__restrict__ keep, const uint8_t *__restrict__ prot,
                                                 uint8_t *__restrict__ fixed, int nx, int ny, int beso) {
  const int npr = nx + 1;
  const int64_t nn = (int64_t)npr * (ny + 1);
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < nn; n += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(n / npr), c = (int)(n - (int64_t)r * npr);
    bool used = false;
    if (r > 0 && c > 0) used |= keep[(size_t)(r - 1) * nx + c - 1] != 0;
    if (r > 0 && c < nx) used |= keep[(size_t)(r - 1) * nx + c] != 0;
    if (r < ny && c > 0) used |= keep[(size_t)r * nx + c - 1] != 0;
    if (r < ny && c < nx) used |= keep[(size_t)r * nx + c] != 0;
    if (!used) fixed[n] = 3;
    else if (beso && !prot[n]) fixed[n] = 0;
  }
}

// ---- ESO_stress criterion ----------------------------------------------------------------------------------------
struct StressTab {
  double B[4][3][8];     // strain-displacement matrices at the 4 corners
  double f1, f2, shear;
};

// nodal stresses: corner stresses of the present adjacent elements (ascending element id), / count
// WITH_STRAIN: the nodal strains as well (planes 3..5 of sn) -- the E_nodes / S_nodes pair of strain_nodes that the
// reference hands to its field plots (optimize.py:63,86,163,183,262,300); keep == nullptr: every element counts
// (BESO never shortens `els`)
template <bool WITH_STRAIN>
__global__ void __launch_bounds__(ET) k_stress_nodes(const __grid_constant__ StressTab T,
                                                     const double2 *__restrict__ u, const uint8_t *__restrict__ keep,
                                                     double *__restrict__ sn /* 3 (6) planes x nn */, int nx, int ny) {
  const int npr = nx + 1;
  const int64_t nn = (int64_t)npr * (ny + 1);
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < nn; n += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(n / npr), c = (int)(n - (int64_t)r * npr);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, e0 = 0.0, e1 = 0.0, e2 = 0.0;
    int cnt = 0;
    // ascending element id: SW (this node is local 2), SE (3), NW (1), NE (0)
    const int er[4] = {r - 1, r - 1, r, r}, ec[4] = {c - 1, c, c - 1, c}, la[4] = {2, 3, 1, 0};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (er[q] < 0 || er[q] >= ny || ec[q] < 0 || ec[q] >= nx) continue;
      if (keep != nullptr && !keep[(size_t)er[q] * nx + ec[q]]) continue;
      const double2 *p = u + (size_t)er[q] * npr + ec[q];
      const double2 q0 = __ldg(p), q1 = __ldg(p + 1), q2 = __ldg(p + npr + 1), q3 = __ldg(p + npr);
      const double w[8] = {q0.x, q0.y, q1.x, q1.y, q2.x, q2.y, q3.x, q3.y};
      double eps[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        double t = 0.0;
#pragma unroll
        for (int j = 0; j < 8; ++j) t = fma(T.B[la[q]][i][j], w[j], t);
        eps[i] = t;
      }
      s0 = add(add(s0, mul(T.f1, eps[0])), mul(T.f2, eps[1]));
      s1 = add(add(s1, mul(T.f2, eps[0])), mul(T.f1, eps[1]));
      s2 = add(s2, mul(T.shear, eps[2]));
      if (WITH_STRAIN) { e0 = add(e0, eps[0]); e1 = add(e1, eps[1]); e2 = add(e2, eps[2]); }
      ++cnt;
    }
    sn[n] = s0 / cnt;
    sn[nn + n] = s1 / cnt;
    sn[2 * nn + n] = s2 / cnt;
    if (WITH_STRAIN) {
      sn[3 * nn + n] = e0 / cnt;
      sn[4 * nn + n] = e1 / cnt;
      sn[5 * nn + n] = e2 / cnt;
    }
  }
}

// strain_els (solver.py:281-316) + von Mises (optimize.py:89); absent elements get 0; reduces max
__global__ void __launch_bounds__(ET) k_vonmises(const double *__restrict__ sn, const uint8_t *__restrict__ keep,
                                                 double *__restrict__ s, int nx, int ny, double *partials,
                                                 unsigned int *ticket, double *out) {
  const int64_t ne = (int64_t)nx * ny;
  const int npr = nx + 1;
  const int64_t nn = (int64_t)npr * (ny + 1);
  double v[1] = {-1.0 / 0.0};
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x) {
    double val = 0.0;
    if (keep[e]) {
      const int r = (int)(e / nx), c = (int)(e - (int64_t)r * nx);
      const size_t n0 = (size_t)r * npr + c;
      double S[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double *p = sn + (size_t)k * nn + n0;
        S[k] = add(add(add(p[0], p[1]), p[npr + 1]), p[npr]) / 4.0;   // nodes in connectivity order
      }
      val = sqrt(add(add(add(mul(S[0], S[0]), -mul(S[0], S[1])), mul(S[1], S[1])), mul(3.0, mul(S[2], S[2]))));
      v[0] = fmax(v[0], val);
    }
    s[e] = val;
  }
  grid_reduce<1, true>(v, partials, ticket, out);
}

// ---- floating parts: present elements that no chain of node-sharing present elements ties to a support ----------------
// (what makes the reference's direct solve singular once ESO has cut a piece loose: optimize.py:84 / :181 + the guard
// :72 / :172).  Flood fill over the element grid: seeds = present elements touching a support node; a tile of 32 x 32
// elements is iterated to its local fixed point in shared memory, the host repeats the pass until no tile changed.
constexpr int IT = 32;
__global__ void __launch_bounds__(ET) k_island_seed(const uint8_t *__restrict__ keep, const uint8_t *__restrict__ support,
                                                    uint8_t *__restrict__ reach, int nx, int ny) {
  const int64_t ne = (int64_t)nx * ny;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(e / nx), c = (int)(e - (int64_t)r * nx);
    reach[e] = (keep[e] && touches(support, r, c, nx + 1)) ? 1 : 0;
  }
}
__global__ void __launch_bounds__(IT * 8) k_island_sweep(const uint8_t *__restrict__ keep, uint8_t *__restrict__ reach, int nx,
                                                         int ny, int *changed) {
  __shared__ uint8_t k[IT + 2][IT + 2], m[IT + 2][IT + 2];
  __shared__ int s_any;
  const int x0 = blockIdx.x * IT, y0 = blockIdx.y * IT;
  for (int t = threadIdx.y * IT + threadIdx.x; t < (IT + 2) * (IT + 2); t += IT * 8) {
    const int ty = t / (IT + 2), tx = t - ty * (IT + 2);
    const int gx = x0 + tx - 1, gy = y0 + ty - 1;
    const bool in = gx >= 0 && gx < nx && gy >= 0 && gy < ny;
    k[ty][tx] = in ? keep[(size_t)gy * nx + gx] : 0;
    m[ty][tx] = in ? reach[(size_t)gy * nx + gx] : 0;
  }
  bool mine = false;
  for (int sweep = 0; sweep < 2 * IT; ++sweep) {
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) s_any = 0;
    __syncthreads();
    bool got = false;
    for (int ly = threadIdx.y; ly < IT; ly += 8) {
      const int ty = ly + 1, tx = threadIdx.x + 1;
      if (k[ty][tx] && !m[ty][tx]) {
        const int nb = m[ty - 1][tx - 1] | m[ty - 1][tx] | m[ty - 1][tx + 1] | m[ty][tx - 1] | m[ty][tx + 1] | m[ty + 1][tx - 1] |
                       m[ty + 1][tx] | m[ty + 1][tx + 1];
        if (nb) { m[ty][tx] = 1; got = true; }           // (a neighbour set in this very sweep only speeds the fill up)
      }
    }
    if (got) { s_any = 1; mine = true; }
    __syncthreads();
    if (!s_any) break;
  }
  for (int ly = threadIdx.y; ly < IT; ly += 8) {
    const int gx = x0 + threadIdx.x, gy = y0 + ly;
    if (gx < nx && gy < ny && m[ly + 1][threadIdx.x + 1]) reach[(size_t)gy * nx + gx] = 1;
  }
  if (mine) *changed = 1;
}
__global__ void __launch_bounds__(ET) k_island_count(const uint8_t *__restrict__ keep, const uint8_t *__restrict__ reach, int64_t ne,
                                                     double *partials, unsigned int *ticket, double *out) {
  double v[1] = {0.0};
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x)
    v[0] += (keep[e] && !reach[e]) ? 1.0 : 0.0;
  grid_reduce<1>(v, partials, ticket, out);
}

int build_prot(topo_handle *h, const int64_t *protect_nodes, int n_protect, uint8_t **prot_out) {
  // protection bitmap lives in the first nn bytes of h->nodal's tail buffer (h->tmp reinterpret)
  uint8_t *prot = reinterpret_cast<uint8_t *>(h->tmp2);
  CUDA_TRY(cudaMemsetAsync(prot, 0, (size_t)h->nn, h->stream));
  if (n_protect > 0) {
    for (int i = 0; i < n_protect; ++i)
      if (protect_nodes[i] < 0 || protect_nodes[i] >= h->nn) {
        topo_set_error("protected node id %lld out of range", (long long)protect_nodes[i]);
        return TOPO_EINVAL;
      }
    int64_t *d_ids = reinterpret_cast<int64_t *>(h->tmp);
    if ((size_t)n_protect * sizeof(int64_t) > (size_t)h->nn * sizeof(double2)) return TOPO_EINVAL;
    CUDA_TRY(cudaMemcpyAsync(d_ids, protect_nodes, sizeof(int64_t) * n_protect, cudaMemcpyHostToDevice, h->stream));
    KL(h, "k_mark_nodes", k_mark_nodes<<<div_up(n_protect, 256), 256, 0, h->stream>>>(prot, d_ids, n_protect));
  }
  *prot_out = prot;
  return TOPO_OK;
}

}  // namespace

extern "C" int topo_energy_sensitivity(topo_t *h, int use_keep) {
  if (!h) return TOPO_EINVAL;
  (void)use_keep;
  CUDA_TRY(cudaSetDevice(h->device));
  KL(h, "k_energy", k_energy<<<el_blocks(h->ne), ET, 0, h->stream>>>(h->ke, h->u, h->keep, h->sens, h->nx, h->ny, h->partials, h->ticket,
                                                   h->red_out));
  KL(h, "k_div_scalar", k_div_scalar<<<el_blocks(h->ne), ET, 0, h->stream>>>(h->sens, h->red_out, h->ne));
  CUDA_TRY(cudaGetLastError());
  return TOPO_OK;
}

extern "C" int topo_beso_filter(topo_t *h, const double w4[4], const double w2x[2], const double w2y[2],
                                const double we[4], int average_with_prev) {
  if (!h || !w4 || !w2x || !w2y || !we) return TOPO_EINVAL;
  CUDA_TRY(cudaSetDevice(h->device));
  BesoW W;
  for (int i = 0; i < 4; ++i) { W.w4[i] = w4[i]; W.we[i] = we[i]; }
  for (int i = 0; i < 2; ++i) { W.w2x[i] = w2x[i]; W.w2y[i] = w2y[i]; }
  W.wesum = ((we[0] + we[1]) + we[2]) + we[3];          // numpy sum of 4 values: left to right
  cudaStream_t st = h->stream;
  // h->sens holds a/amax already (topo_energy_sensitivity); nodes read it with divisor 1
  double one = 1.0;
  CUDA_TRY(cudaMemcpyAsync(h->red_out + 4, &one, sizeof(double), cudaMemcpyHostToDevice, st));
  KL(h, "k_beso_nodes", k_beso_nodes<<<el_blocks(h->nn), ET, 0, st>>>(W, h->sens, h->red_out + 4, h->nodal, h->nx, h->ny));
  KL(h, "k_beso_els", k_beso_els<<<el_blocks(h->ne), ET, 0, st>>>(W, h->nodal, h->sens, h->nx, h->ny, h->partials, h->ticket, h->red_out));
  KL(h, "k_beso_hist", k_beso_hist<<<el_blocks(h->ne), ET, 0, st>>>(h->sens, h->sens_prev, h->red_out, average_with_prev ? 1 : 0, h->ne,
                                               h->partials, h->ticket, h->red_out + 1));
  KL(h, "k_div_scalar", k_div_scalar<<<el_blocks(h->ne), ET, 0, st>>>(h->sens, h->red_out + 1, h->ne));
  CUDA_TRY(cudaGetLastError());
  return TOPO_OK;
}

extern "C" int topo_beso_nodes(topo_t *h, const double w4[4], const double w2x[2], const double w2y[2]) {
  if (!h || !w4 || !w2x || !w2y) return TOPO_EINVAL;
  CUDA_TRY(cudaSetDevice(h->device));
  BesoW W = {};
  for (int i = 0; i < 4; ++i) W.w4[i] = w4[i];
  for (int i = 0; i < 2; ++i) { W.w2x[i] = w2x[i]; W.w2y[i] = w2y[i]; }
  double one = 1.0;
  CUDA_TRY(cudaMemcpyAsync(h->red_out + 4, &one, sizeof(double), cudaMemcpyHostToDevice, h->stream));
  KL(h, "k_beso_nodes", k_beso_nodes<<<el_blocks(h->nn), ET, 0, h->stream>>>(W, h->sens, h->red_out + 4, h->nodal, h->nx, h->ny));
  CUDA_TRY(cudaGetLastError());
  return TOPO_OK;
}

extern "C" int topo_beso_elements(topo_t *h, const double we[4]) {
  if (!h || !we) return TOPO_EINVAL;
  CUDA_TRY(cudaSetDevice(h->device));
  BesoW W = {};
  for (int i = 0; i < 4; ++i) W.we[i] = we[i];
  W.wesum = ((we[0] + we[1]) + we[2]) + we[3];
  KL(h, "k_beso_els", k_beso_els<<<el_blocks(h->ne), ET, 0, h->stream>>>(W, h->nodal, h->sens, h->nx, h->ny, h->partials, h->ticket,
                                                     h->red_out));
  KL(h, "k_div_scalar", k_div_scalar<<<el_blocks(h->ne), ET, 0, h->stream>>>(h->sens, h->red_out, h->ne));
  CUDA_TRY(cudaGetLastError());
  return TOPO_OK;
}

extern "C" int topo_beso_elements_radius(topo_t *h, double r_min, double dx, double dy) {
  if (!h || !(r_min > 0.0) || !(dx > 0.0) || !(dy > 0.0)) return TOPO_EINVAL;
  CUDA_TRY(cudaSetDevice(h->device));
  BesoTaps T;
  T.lo = -BR_MAX + 1;
  T.hi = BR_MAX;
  bool edge_hit = false;
  int count = 0;
  for (int dj = T.lo; dj <= T.hi; ++dj)
    for (int di = T.lo; di <= T.hi; ++di) {
      const double x = (di - 0.5) * dx, y = (dj - 0.5) * dy;
      const double d = std::sqrt(x * x + y * y);
      const bool in = d < r_min;
      T.r[dj - T.lo][di - T.lo] = in ? d : -1.0;
      count += in ? 1 : 0;
      if (in && (dj == T.lo || dj == T.hi || di == T.lo || di == T.hi)) edge_hit = true;
    }
  if (edge_hit || count < 2) {
    topo_set_error("topo_beso_elements_radius: radius %.3g must cover at least 2 nodes and at most %d node columns/rows", r_min,
                   2 * BR_MAX - 2);
    return TOPO_EINVAL;
  }
  KL(h, "k_beso_els_radius", k_beso_els_radius<<<el_blocks(h->ne), ET, 0, h->stream>>>(T, h->nodal, h->sens, h->nx, h->ny, h->partials,
                                                                                       h->ticket, h->red_out));
  KL(h, "k_div_scalar", k_div_scalar<<<el_blocks(h->ne), ET, 0, h->stream>>>(h->sens, h->red_out, h->ne));
  CUDA_TRY(cudaGetLastError());
  return TOPO_OK;
}

extern "C" int topo_rank_select(topo_t *h, int64_t k, double *alpha) {
  if (!h || !alpha || k < 0 || k >= h->ne) {
    topo_set_error("topo_rank_select: rank %lld outside [0, %lld)", (long long)k, h ? (long long)h->ne : 0ll);
    return TOPO_EINVAL;
  }
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  KL(h, "k_sel_init", k_sel_init<<<1, 256, 0, st>>>(h->sel, (unsigned long long)k));
  for (int pass = 0; pass < 8; ++pass) {
    KL(h, "k_sel_hist", k_sel_hist<<<el_blocks(h->ne), ET, 0, st>>>(h->sens, h->ne, h->sel, h->ticket + 2, h->red_out + 2));
  }
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(alpha, h->red_out + 2, sizeof(double), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return TOPO_OK;
}

static int after_cons_change(topo_handle *h) {
  h->diag_valid = false;
  h->gmg_valid = false;
  TOPO_TRY(topo_apply_loads(h));
  TOPO_TRY(topo_mask_vector(h, h->u));
  return TOPO_OK;
}

extern "C" int topo_beso_apply(topo_t *h, double alpha, const int64_t *protect_nodes, int n_protect, int64_t *n_kept,
                               int64_t *n_protected) {
  if (!h || (n_protect > 0 && !protect_nodes)) return TOPO_EINVAL;
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  uint8_t *prot = nullptr;
  TOPO_TRY(build_prot(h, protect_nodes, n_protect, &prot));
  CUDA_TRY(cudaMemcpyAsync(h->red_out + 2, &alpha, sizeof(double), cudaMemcpyHostToDevice, st));
  KL(h, "k_beso_mask", k_beso_mask<<<el_blocks(h->ne), ET, 0, st>>>(h->sens, h->red_out + 2, prot, h->keep, h->nx, h->ny, h->partials,
                                               h->ticket, h->red_out + 4));
  KL(h, "k_del_node", k_del_node<<<el_blocks(h->nn), ET, 0, st>>>(h->keep, prot, h->fixed, h->nx, h->ny, 1));
  CUDA_TRY(cudaGetLastError());
  double cnt[2];
  CUDA_TRY(cudaMemcpyAsync(cnt, h->red_out + 4, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (n_kept) *n_kept = (int64_t)cnt[0];
  if (n_protected) *n_protected = (int64_t)cnt[1];
  return after_cons_change(h);
}

extern "C" int topo_eso_apply(topo_t *h, double rr, const int64_t *protect_nodes, int n_protect, int64_t *n_kept) {
  if (!h || (n_protect > 0 && !protect_nodes)) return TOPO_EINVAL;
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  uint8_t *prot = nullptr;
  TOPO_TRY(build_prot(h, protect_nodes, n_protect, &prot));
  KL(h, "k_eso_mask", k_eso_mask<<<el_blocks(h->ne), ET, 0, st>>>(h->sens, rr, prot, h->keep, h->scale, h->nx, h->ny, h->partials,
                                              h->ticket, h->red_out + 4));
  KL(h, "k_del_node", k_del_node<<<el_blocks(h->nn), ET, 0, st>>>(h->keep, prot, h->fixed, h->nx, h->ny, 0));
  CUDA_TRY(cudaGetLastError());
  double cnt = 0.0;
  CUDA_TRY(cudaMemcpyAsync(&cnt, h->red_out + 4, sizeof(double), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (n_kept) *n_kept = (int64_t)cnt;
  return after_cons_change(h);
}

extern "C" int topo_floating_elements(topo_t *h, const int64_t *support_nodes, int n_support, int64_t *n_floating) {
  if (!h || !n_floating || (n_support > 0 && !support_nodes)) return TOPO_EINVAL;
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  uint8_t *support = nullptr;
  TOPO_TRY(build_prot(h, support_nodes, n_support, &support));        // node bitmap in tmp2
  uint8_t *reach = reinterpret_cast<uint8_t *>(h->rho_new);           // scratch: nobody reads rho_new between calls
  int *changed = reinterpret_cast<int *>(h->red_out + 8);
  KL(h, "k_island_seed", k_island_seed<<<el_blocks(h->ne), ET, 0, st>>>(h->keep, support, reach, h->nx, h->ny));
  dim3 grid(div_up(h->nx, IT), div_up(h->ny, IT)), block(IT, 8);
  const int max_pass = 4 * (div_up(h->nx, IT) + div_up(h->ny, IT)) + 8;   // a path can wind, but every pass advances >= one tile
  for (int pass = 0; pass < max_pass * 8; ++pass) {
    CUDA_TRY(cudaMemsetAsync(changed, 0, sizeof(int), st));
    KL(h, "k_island_sweep", k_island_sweep<<<grid, block, 0, st>>>(h->keep, reach, h->nx, h->ny, changed));
    int flag = 0;
    CUDA_TRY(cudaMemcpyAsync(&flag, changed, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (!flag) break;
  }
  KL(h, "k_island_count", k_island_count<<<el_blocks(h->ne), ET, 0, st>>>(h->keep, reach, h->ne, h->partials, h->ticket, h->red_out));
  CUDA_TRY(cudaGetLastError());
  double cnt = 0.0;
  CUDA_TRY(cudaMemcpyAsync(&cnt, h->red_out, sizeof(double), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  *n_floating = (int64_t)cnt;
  return TOPO_OK;
}

extern "C" int topo_beso_save_history(topo_t *h) {
  if (!h) return TOPO_EINVAL;
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaMemcpyAsync(h->sens_prev, h->sens, sizeof(double) * (size_t)h->ne, cudaMemcpyDeviceToDevice, h->stream));
  return TOPO_OK;
}

static StressTab make_stress_tab(double E, double nu, double dx, double dy) {
  StressTab T;
  const double ri[4] = {-1.0, 1.0, 1.0, -1.0}, si[4] = {-1.0, -1.0, 1.0, 1.0};
  for (int k = 0; k < 4; ++k) {
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 8; ++j) T.B[k][i][j] = 0.0;
    for (int a = 0; a < 4; ++a) {
      const double dNx = 0.25 * ri[a] * (1 + si[k] * si[a]) * (2.0 / dx);
      const double dNy = 0.25 * si[a] * (1 + ri[k] * ri[a]) * (2.0 / dy);
      T.B[k][0][2 * a] = dNx;
      T.B[k][1][2 * a + 1] = dNy;
      T.B[k][2][2 * a] = dNy;
      T.B[k][2][2 * a + 1] = dNx;
    }
  }
  T.shear = E / (2 * (1 + nu));
  T.f1 = E / (1 - nu * nu);
  T.f2 = nu * E / (1 - nu * nu);
  return T;
}

extern "C" int topo_vonmises_sensitivity(topo_t *h, double E, double nu, double dx, double dy) {
  if (!h || !(dx > 0.0) || !(dy > 0.0)) return TOPO_EINVAL;
  CUDA_TRY(cudaSetDevice(h->device));
  const StressTab T = make_stress_tab(E, nu, dx, dy);
  cudaStream_t st = h->stream;
  double *sn = reinterpret_cast<double *>(h->z);         // z and p are contiguous scratch outside a solve
  KL(h, "k_stress_nodes", k_stress_nodes<false><<<el_blocks(h->nn), ET, 0, st>>>(T, h->u, h->keep, sn, h->nx, h->ny));
  KL(h, "k_vonmises", k_vonmises<<<el_blocks(h->ne), ET, 0, st>>>(sn, h->keep, h->sens, h->nx, h->ny, h->partials, h->ticket, h->red_out));
  KL(h, "k_div_scalar", k_div_scalar<<<el_blocks(h->ne), ET, 0, st>>>(h->sens, h->red_out, h->ne));
  CUDA_TRY(cudaGetLastError());
  return TOPO_OK;
}

extern "C" int topo_strain_nodes(topo_t *h, double E, double nu, double dx, double dy, int present_only,
                                 double *strain_host, double *stress_host) {
  if (!h || !strain_host || !stress_host || !(dx > 0.0) || !(dy > 0.0)) return TOPO_EINVAL;
  CUDA_TRY(cudaSetDevice(h->device));
  const StressTab T = make_stress_tab(E, nu, dx, dy);
  cudaStream_t st = h->stream;
  double *sn = reinterpret_cast<double *>(h->z);         // z, p and Ap are contiguous scratch outside a solve: 6 planes
  KL(h, "k_stress_nodes", k_stress_nodes<true><<<el_blocks(h->nn), ET, 0, st>>>(T, h->u, present_only ? h->keep : nullptr, sn,
                                                                                  h->nx, h->ny));
  CUDA_TRY(cudaGetLastError());
  const size_t plane3 = sizeof(double) * 3 * (size_t)h->nn;
  CUDA_TRY(cudaMemcpyAsync(stress_host, sn, plane3, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(strain_host, sn + 3 * (size_t)h->nn, plane3, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return TOPO_OK;
}
