// Stages 3+4 of the SIMP iteration: SIMP interpolation, element sensitivity fused with the objective
// reduction, radius filter as a shared-memory-tiled stencil, optimality-criteria update with a
// device-side bisection.  Reference: optimize.py:430,443-452; solver.py:412-477.
#include <algorithm>
#include <cmath>

#include "topo_internal.cuh"

namespace {

constexpr int ET = 256;
inline int el_blocks(int64_t n) { return (int)std::min<int64_t>(div_up(n, ET), 148 * 8); }

// numpy minimum/maximum propagate NaN (CUDA fmin/fmax do not)
__device__ __forceinline__ double np_min(double a, double b) { return (a != a) ? a : ((b != b) ? b : (a < b ? a : b)); }
__device__ __forceinline__ double np_max(double a, double b) { return (a != a) ? a : ((b != b) ? b : (a > b ? a : b)); }

__device__ __forceinline__ double powp(double x, double p) {
  if (p == 2.0) return x * x;            // numpy's fast path for ** 2 (rho**(penal-1), optimize.py:444)
  if (p == 1.0) return x;
  return pow(x, p);
}

// scale_e = emin + rho^penal (emax - emin)                               (optimize.py:430)
__global__ void k_simp_scale(const double *__restrict__ rho, double *__restrict__ scale, int64_t ne, double penal,
                             double emin, double emax) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x)
    scale[e] = emin + pow(rho[e], penal) * (emax - emin);
}

// s_e = u_e^T ke u_e ; dc_e = (-penal rho^(penal-1) (emax-emin)) s_e ; objective += scale_e s_e   (optimize.py:443-444)
// Evaluated as sum_k (l_k . d)^2 with ke = sum_k l_k l_k^T and d = u_e minus the displacement of local node 0
// (ke annihilates translations): the reference's u^T ke u loses all digits where u_e is almost a rigid motion
// (free end of a long cantilever: |u|^2 |ke| eps exceeds the true energy on fine meshes) and then turns
// NEGATIVE at random, which np.sqrt(-dc/lmid) converts into NaN densities (solver.py:441).  Same value to
// rounding where the reference is accurate; non-negative everywhere.
// A thread marches up a column of elements: the two upper nodes of an element are the lower
// nodes of the next one, so a row costs two 16-byte loads instead of four, there is no 64-bit index division, and only
// the `rank` non-zero rows of the factor are evaluated (5 for the plane-stress quad4: three rigid-body modes).
__global__ void __launch_bounds__(ET) k_simp_sens_march(const __grid_constant__ KeParam lf, int rank, const double2 *__restrict__ u,
                                                        const double *__restrict__ rho, const double *__restrict__ scale,
                                                        double *__restrict__ dc, int nx, int ny, int rows_per_chunk, double penal,
                                                        double emin, double emax, double *partials, unsigned int *ticket,
                                                        double *out) {
  const int npr = nx + 1;
  const int c = blockIdx.x * ET + threadIdx.x;
  const int r0 = blockIdx.y * rows_per_chunk, r1 = min(r0 + rows_per_chunk, ny);
  double v[1] = {0.0};
  if (c < nx && r0 < r1) {
    const double2 *p = u + (size_t)r0 * npr + c;
    double2 q0 = __ldg(p), q1 = __ldg(p + 1);
    size_t e = (size_t)r0 * nx + c;
    for (int r = r0; r < r1; ++r, p += npr, e += nx) {
      const double2 q3 = __ldg(p + npr), q2 = __ldg(p + npr + 1);
      const double d[6] = {q1.x - q0.x, q1.y - q0.y, q2.x - q0.x, q2.y - q0.y, q3.x - q0.x, q3.y - q0.y};
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (k < rank) {
          double t = 0.0;
#pragma unroll
          for (int i = 0; i < 6; ++i) t = fma(lf.k[8 * k + 2 + i], d[i], t);
          s = fma(t, t, s);
        }
      }
      const double rh = rho[e];
      dc[e] = (-penal * powp(rh, penal - 1.0) * (emax - emin)) * s;
      v[0] = fma(scale[e], s, v[0]);
      q0 = q3;
      q1 = q2;
    }
  }
  grid_reduce<1>(v, partials, ticket, out);
}

// ---- filter ----------------------------------------------------------------------------------------
constexpr int FH_MAX = 8;               // stencil half width (elements)
struct FilterTaps {
  int hx, hy;
  double w[(2 * FH_MAX + 1) * (2 * FH_MAX + 1)];   // row-major [dj+hy][di+hx], <=0 entries are skipped
};

constexpr int FT = 32;                  // tile edge (elements)
// out_i = sum_j w_ij rho_j dc_j / (sum_j w_ij * max(1e-3, rho_i)); taps in (dj, di) row-major order
// Row-slab mode: ghost_lo / ghost_hi hold rho*dc of the hy element rows below / above the slab (nullptr at the
// ends of the global mesh, where the stencil is truncated like the reference's).
__global__ void __launch_bounds__(256) k_filter(const __grid_constant__ FilterTaps taps, const double *__restrict__ rho,
                                                const double *__restrict__ dc, double *__restrict__ out, int nx,
                                                int ny, const double *__restrict__ ghost_lo,
                                                const double *__restrict__ ghost_hi) {
  extern __shared__ double tile[];      // (FT+2hy) x (FT+2hx)
  const int hx = taps.hx, hy = taps.hy;
  const int tw = FT + 2 * hx, th = FT + 2 * hy;
  const int x0 = blockIdx.x * FT, y0 = blockIdx.y * FT;
  for (int t = threadIdx.y * 32 + threadIdx.x; t < tw * th; t += 256) {
    const int ty = t / tw, tx = t - ty * tw;
    const int gx = x0 + tx - hx, gy = y0 + ty - hy;
    double q = 0.0;
    if (gx >= 0 && gx < nx) {
      if (gy >= 0 && gy < ny) {
        const size_t e = (size_t)gy * nx + gx;
        q = rho[e] * dc[e];
      } else if (gy < 0 && ghost_lo != nullptr) {
        q = ghost_lo[(size_t)(gy + hy) * nx + gx];          // rows -hy .. -1
      } else if (gy >= ny && ghost_hi != nullptr) {
        q = ghost_hi[(size_t)(gy - ny) * nx + gx];          // rows ny .. ny+hy-1
      }
    }
    tile[t] = q;
  }
  __syncthreads();
  const int gx = x0 + threadIdx.x;
  if (gx >= nx) return;
  const int wrow = 2 * hx + 1;
#pragma unroll 1
  for (int k = 0; k < FT / 8; ++k) {
    const int ly = threadIdx.y + 8 * k;
    const int gy = y0 + ly;
    if (gy >= ny) break;
    double num = 0.0, den = 0.0;
    for (int dj = -hy; dj <= hy; ++dj) {
      const bool rowok = (gy + dj >= 0 || ghost_lo != nullptr) && (gy + dj < ny || ghost_hi != nullptr);
      const double *trow = tile + (ly + hy + dj) * tw + threadIdx.x + hx;
      for (int di = -hx; di <= hx; ++di) {
        const double w = taps.w[(dj + hy) * wrow + di + hx];
        if (w > 0.0) {
          num = fma(w, trow[di], num);
          if (rowok && gx + di >= 0 && gx + di < nx) den += w;
        }
      }
    }
    const size_t e = (size_t)gy * nx + gx;
    out[e] = num / (den * np_max(0.001, rho[e]));
  }
}

// Same filter for the common square stencils (hx == hy == H known at compile time: H = 3 for the reference's r_min = 4
// element widths, 2 for r = 2.5, 1 for r = 1.5).  The generic kernel spends ~500 instructions per element on the tap
// loop (weight fetched through a dynamic constant index, a branch and the domain test for the denominator per tap:
// 0.45 TB/s = 7 % of the HBM roofline in ncu); here a thread owns FB vertically adjacent elements, reads each tile
// column segment once for all of them ((2H+1)(FB+2H) LDS.64 instead of FB (2H+1)^2), keeps the (H+1)^2 distinct weights
// in registers, and the denominator is the constant sum of all weights except in the frame of H elements along the
// mesh boundary, where it is summed tap by tap as before.
template <int H, int FB>
__global__ void __launch_bounds__(256) k_filter_sq(const __grid_constant__ FilterTaps taps, const double *__restrict__ rho,
                                                   const double *__restrict__ dc, double *__restrict__ out, int nx, int ny,
                                                   const double *__restrict__ ghost_lo, const double *__restrict__ ghost_hi) {
  constexpr int W = 2 * H + 1, TW = FT + 2 * H, TH = 8 * FB + 2 * H;      // block (32, 8): FT columns x 8*FB rows
  __shared__ double tile[TH * TW];
  const int x0 = blockIdx.x * FT, y0 = blockIdx.y * (8 * FB);
  for (int t = threadIdx.y * 32 + threadIdx.x; t < TW * TH; t += 256) {
    const int ty = t / TW, tx = t - ty * TW;
    const int gx = x0 + tx - H, gy = y0 + ty - H;
    double q = 0.0;
    if (gx >= 0 && gx < nx) {
      if (gy >= 0 && gy < ny) {
        const size_t e = (size_t)gy * nx + gx;
        q = rho[e] * dc[e];
      } else if (gy < 0 && ghost_lo != nullptr) {
        q = ghost_lo[(size_t)(gy + H) * nx + gx];
      } else if (gy >= ny && ghost_hi != nullptr) {
        q = ghost_hi[(size_t)(gy - ny) * nx + gx];
      }
    }
    tile[t] = q;
  }
  // weights by (|dj|, |di|); taps outside the radius carry 0 (the generic kernel skips them: same sums)
  double w[H + 1][H + 1];
  double wsum = 0.0;
#pragma unroll
  for (int a = 0; a <= H; ++a)
#pragma unroll
    for (int b = 0; b <= H; ++b) {
      const double v = taps.w[(a + H) * W + b + H];
      w[a][b] = v > 0.0 ? v : 0.0;
      wsum += w[a][b] * ((a == 0 ? 1.0 : 2.0) * (b == 0 ? 1.0 : 2.0));
    }
  __syncthreads();
  const int gx = x0 + threadIdx.x;
  const int ly0 = threadIdx.y * FB, gy0 = y0 + ly0;
  if (gx >= nx || gy0 >= ny) return;
  double num[FB];
#pragma unroll
  for (int k = 0; k < FB; ++k) num[k] = 0.0;
#pragma unroll
  for (int di = -H; di <= H; ++di) {
    double col[FB + 2 * H];
#pragma unroll
    for (int i = 0; i < FB + 2 * H; ++i) col[i] = tile[(ly0 + i) * TW + threadIdx.x + H + di];
#pragma unroll
    for (int k = 0; k < FB; ++k)
#pragma unroll
      for (int dj = -H; dj <= H; ++dj) num[k] = fma(w[dj < 0 ? -dj : dj][di < 0 ? -di : di], col[k + H + dj], num[k]);
  }
#pragma unroll
  for (int k = 0; k < FB; ++k) {
    const int gy = gy0 + k;
    if (gy >= ny) break;
    double den = wsum;
    const bool lo_open = ghost_lo != nullptr, hi_open = ghost_hi != nullptr;
    if (gx < H || gx >= nx - H || (gy < H && !lo_open) || (gy >= ny - H && !hi_open)) {     // frame: truncated stencil
      den = 0.0;
      for (int dj = -H; dj <= H; ++dj) {
        const bool rowok = (gy + dj >= 0 || lo_open) && (gy + dj < ny || hi_open);
        for (int di = -H; di <= H; ++di) {
          const double v = taps.w[(dj + H) * W + di + H];
          if (v > 0.0 && rowok && gx + di >= 0 && gx + di < nx) den += v;
        }
      }
    }
    const size_t e = (size_t)gy * nx + gx;
    out[e] = num[k] / (den * np_max(0.001, rho[e]));
  }
}

// ---- optimality criteria ---------------------------------------------------------------------------
__device__ __forceinline__ double oc_candidate(double rho, double dc, double lmid, double move) {
  // np.maximum(0, np.maximum(rho-move, np.minimum(1, np.minimum(rho+move, rho*np.sqrt(-dc/lmid)))))
  const double b = rho * sqrt(-dc / lmid);
  return np_max(0.0, np_max(rho - move, np_min(1.0, np_min(rho + move, b))));
}

__global__ void k_oc_init(OcState *oc, double g) {
  oc->l1 = 0.0;
  oc->l2 = 1e9;
  oc->g = g;
  oc->gt = g;
  oc->lmid = 0.0;
  oc->steps = 0;
  oc->change = 0.0;
  oc->need_exact = 0;
  oc->ka = -1.0;
  oc->kb = 1e300;
  oc->done = ((oc->l2 - oc->l1) / (oc->l1 + oc->l2) > 1e-3) ? 0 : 1;
}

// Bisection steps whose outcome is already known from the measured bracket (ka, kb): the reference's loop
// (solver.py:441-447) is replayed with its own arithmetic, only the evaluation of gt is skipped.
__device__ __forceinline__ void oc_fastforward(OcState *oc) {
  if (oc->done || oc->need_exact) return;
  double l1 = oc->l1, l2 = oc->l2, lmid_last = oc->lmid;
  const double ka = oc->ka, kb = oc->kb;
  int steps = 0;
  bool done = false;
  for (int it = 0; it < 4096; ++it) {
    if (!((l2 - l1) / (l1 + l2) > 1e-3)) { done = true; break; }
    const double lmid = 0.5 * (l2 + l1);
    if (lmid >= kb) l2 = lmid;            // gt(lmid) <= gt(kb) < 0
    else if (lmid <= ka) l1 = lmid;       // gt(lmid) >= gt(ka) > 0
    else break;
    lmid_last = lmid;
    ++steps;
  }
  oc->l1 = l1;
  oc->l2 = l2;
  oc->lmid = lmid_last;
  oc->steps += steps;
  oc->done = done ? 1 : 0;
}

// A_e = rho sqrt(-dc): rho*sqrt(-dc/lmid) == A_e / sqrt(lmid) up to a few ulp
__global__ void __launch_bounds__(ET) k_oc_prep(const double *__restrict__ rho, const double *__restrict__ dc,
                                                double *__restrict__ amp, int64_t ne) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x)
    amp[e] = rho[e] * sqrt(-dc[e]);
}

// EXACT bisection step with the reference's formula: gt = g + sum(rho_new(lmid) - rho);
// gt > 0 ? l1 = lmid : l2 = lmid.  only_when_needed: run only if the fast pass asked for it.
__device__ __forceinline__ void oc_decide_step(OcState *oc, double sum) {
  const double lmid = 0.5 * (oc->l2 + oc->l1);
  const double gt = oc->g + sum;
  if (gt > 0.0) oc->l1 = lmid; else oc->l2 = lmid;
  oc->gt = gt;
  oc->lmid = lmid;
  oc->steps += 1;
  oc->need_exact = 0;
  oc->done = ((oc->l2 - oc->l1) / (oc->l1 + oc->l2) > 1e-3) ? 0 : 1;
  oc_fastforward(oc);
}
// defer != 0 (row-slab mode): leave the per-rank sum in out[]; the decision is taken by k_oc_decide after the
// host's all-reduce
__global__ void __launch_bounds__(ET) k_oc_step(const double *__restrict__ rho, const double *__restrict__ dc,
                                                int64_t ne, double move, int only_when_needed, double *partials,
                                                unsigned int *ticket, double *out, OcState *oc, int defer) {
  if (oc->done) return;
  if (only_when_needed && !oc->need_exact) return;
  const double lmid = 0.5 * (oc->l2 + oc->l1);
  double v[1] = {0.0};
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x) {
    const double rh = rho[e];
    v[0] += oc_candidate(rh, dc[e], lmid, move) - rh;
  }
  if (grid_reduce<1>(v, partials, ticket, out) && threadIdx.x == 0 && !defer) oc_decide_step(oc, v[0]);
}

// FAST pass: the 15 multipliers of the next 4 bisection levels (heap order, generated with exactly the
// reference's 0.5*(l2+l1) arithmetic) are evaluated together with the cheap form A_e/sqrt(lmid).  The
// last block then walks the decision tree.  Each clamped term differs from the reference's by a few
// ulp (<= 1e-15 absolute), so a sum further than `margin` from zero has the reference's sign; a sum
// inside the margin stops the walk and requests one exact step (k_oc_step) for that decision.
constexpr int OC_LEVELS = 4, OC_CAND = (1 << OC_LEVELS) - 1;
// walks the decision tree of one fast pass from the 15 candidate sums (+ the common part in v[OC_CAND])
__device__ __forceinline__ void oc_decide_multi(OcState *oc, const double *v, double margin) {
  double l1 = oc->l1, l2 = oc->l2;
  int j = 0, steps = 0;
  bool done = false, need = false;
  double gt_last = oc->gt, lmid_last = oc->lmid;
  for (int d = 0; d < OC_LEVELS; ++d) {
    if (!((l2 - l1) / (l1 + l2) > 1e-3)) { done = true; break; }
    double vj = 0.0;
#pragma unroll
    for (int q = 0; q < OC_CAND; ++q) vj = (q == j) ? v[q] : vj;
    const double gt = oc->g + vj + v[OC_CAND];
    const bool is_nan = !(gt == gt);
    if (!is_nan && !(fabs(gt) > margin)) { need = true; break; }    // too close to call: the exact step decides
    const double lmid = 0.5 * (l2 + l1);                            // == the candidate evaluated at heap node j
    if (gt > 0.0) { l1 = lmid; j = 2 * j + 2; } else { l2 = lmid; j = 2 * j + 1; }   // NaN -> else, like the reference
    gt_last = gt;
    lmid_last = lmid;
    ++steps;
  }
  if (!done && !need && !((l2 - l1) / (l1 + l2) > 1e-3)) done = true;
  oc->l1 = l1;
  oc->l2 = l2;
  oc->gt = gt_last;
  oc->lmid = lmid_last;
  oc->steps += steps;
  oc->need_exact = need ? 1 : 0;
  oc->done = done ? 1 : 0;
  oc_fastforward(oc);
}
// probe pass: sums at the 15 multipliers base*2^(j-7) only tighten the bracket (no bisection step is taken)
__device__ __forceinline__ void oc_decide_probe(OcState *oc, const double *v, double margin, double base) {
  double ka = oc->ka, kb = oc->kb;
  for (int j = 0; j < OC_CAND; ++j) {
    const double lam = ldexp(base, j - 7);
    const double gt = oc->g + v[j] + v[OC_CAND];
    if (gt > margin) ka = fmax(ka, lam);
    else if (gt < -margin) kb = fmin(kb, lam);
  }
  oc->ka = ka;
  oc->kb = kb;
  oc_fastforward(oc);
}
// grid of the multi-candidate pass: a few CTAs per SM (16 partial sums per CTA are folded by the last block)
static int oc_multi_blocks(const topo_handle *h) { return 148 * h->knobs.oc_ctas_per_sm; }
__global__ void __launch_bounds__(ET) k_oc_multi(const double *__restrict__ rho, const double *__restrict__ amp,
                                                 int64_t ne, double move, double margin, double *partials,
                                                 unsigned int *ticket, double *out, OcState *oc, int defer,
                                                 double probe_base) {
  if (oc->done || oc->need_exact) return;
  // candidate multipliers in heap order; each of the first 15 threads does ONE 1/sqrt and shares it
  __shared__ double s_tinv[OC_CAND];
  if (probe_base > 0.0) {
    if (threadIdx.x < OC_CAND) s_tinv[threadIdx.x] = 1.0 / sqrt(ldexp(probe_base, (int)threadIdx.x - 7));
  } else if (threadIdx.x < OC_CAND) {
    double lo[OC_CAND], hi[OC_CAND], lmj = 0.0;
    lo[0] = oc->l1;
    hi[0] = oc->l2;
#pragma unroll
    for (int j = 0; j < OC_CAND; ++j) {
      const double m = 0.5 * (hi[j] + lo[j]);
      if (j == (int)threadIdx.x) lmj = m;
      if (2 * j + 2 < OC_CAND) {
        lo[2 * j + 1] = lo[j]; hi[2 * j + 1] = m;        // gt <= 0 : l2 = lmid
        lo[2 * j + 2] = m; hi[2 * j + 2] = hi[j];        // gt  > 0 : l1 = lmid
      }
    }
    s_tinv[threadIdx.x] = 1.0 / sqrt(lmj);
  }
  __syncthreads();
  double tinv[OC_CAND];
#pragma unroll
  for (int j = 0; j < OC_CAND; ++j) tinv[j] = s_tinv[j];
  double tmin = tinv[0], tmax = tinv[0];
#pragma unroll
  for (int j = 1; j < OC_CAND; ++j) { tmin = fmin(tmin, tinv[j]); tmax = fmax(tmax, tinv[j]); }
  double v[OC_CAND + 1];                 // [OC_CAND]: contributions common to all candidates (+ NaN contamination)
#pragma unroll
  for (int j = 0; j <= OC_CAND; ++j) v[j] = 0.0;
  auto element = [&](double rh, double am) {
    const double cl = fmax(0.0, rh - move), ch = fmin(1.0, rh + move);
    // elements clamped the same way by ALL 15 multipliers of this pass contribute a common constant
    // (exactly: a*t is monotone in t, so the clamp result is identical for every candidate)
    const double bmin = am * tmin, bmax = am * tmax;
    if (bmax <= cl) {
      v[OC_CAND] += cl - rh;
    } else if (bmin >= ch) {
      v[OC_CAND] += ch - rh;
    } else {
      v[OC_CAND] += (am - am) + (rh - rh);   // numpy's minimum/maximum propagate NaN; fmin/fmax do not
#pragma unroll
      for (int j = 0; j < OC_CAND; ++j) v[j] += fmax(cl, fmin(ch, am * tinv[j])) - rh;
    }
  };
  // 16-byte loads, four elements in flight per thread: at 2 CTAs per SM the plain one-element loop kept 8 KB per SM in
  // flight and ran at 0.9 TB/s (ncu, profiles/r2_stage_kernels.md); the assignment of elements to threads is fixed, so
  // the sums stay deterministic
  const int64_t npair = ne >> 1, stride = (int64_t)gridDim.x * blockDim.x;
  const double2 *rho2 = reinterpret_cast<const double2 *>(rho), *amp2 = reinterpret_cast<const double2 *>(amp);
  int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; q + stride < npair; q += 2 * stride) {
    const double2 r0 = __ldg(rho2 + q), a0 = __ldg(amp2 + q), r1 = __ldg(rho2 + q + stride), a1 = __ldg(amp2 + q + stride);
    element(r0.x, a0.x); element(r0.y, a0.y); element(r1.x, a1.x); element(r1.y, a1.y);
  }
  if (q < npair) {
    const double2 r0 = __ldg(rho2 + q), a0 = __ldg(amp2 + q);
    element(r0.x, a0.x); element(r0.y, a0.y);
  }
  if ((ne & 1) && blockIdx.x == 0 && threadIdx.x == 0) element(rho[ne - 1], amp[ne - 1]);
  if (grid_reduce<OC_CAND + 1>(v, partials, ticket, out) && threadIdx.x == 0 && !defer) {
    if (probe_base > 0.0) oc_decide_probe(oc, v, margin, probe_base);
    else oc_decide_multi(oc, v, margin);
  }
}

// exact gt at the last evaluated multiplier (what the reference returns as the new constraint value)
__global__ void __launch_bounds__(ET) k_oc_final_gt(const double *__restrict__ rho, const double *__restrict__ dc,
                                                    int64_t ne, double move, double *partials, unsigned int *ticket,
                                                    double *out, OcState *oc, int defer) {
  if (oc->steps == 0) return;
  const double lmid = oc->lmid;
  double v[1] = {0.0};
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x) {
    const double rh = rho[e];
    v[0] += oc_candidate(rh, dc[e], lmid, move) - rh;
  }
  if (grid_reduce<1>(v, partials, ticket, out) && threadIdx.x == 0 && !defer) oc->gt = oc->g + v[0];
}

// rho_new at the LAST evaluated multiplier (what the reference returns) + change = max|rho_new - rho|
__global__ void __launch_bounds__(ET) k_oc_apply(const double *__restrict__ rho, const double *__restrict__ dc,
                                                 double *__restrict__ rho_new, int64_t ne, double move,
                                                 double *partials, unsigned int *ticket, double *out, OcState *oc,
                                                 int defer) {
  const double lmid = oc->lmid;
  const int steps = oc->steps;
  double v[1] = {0.0};
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x) {
    const double rh = rho[e];
    // zero bisection steps: the reference returns its zero-initialised rho_new (solver.py:439)
    const double rn = steps > 0 ? oc_candidate(rh, dc[e], lmid, move) : 0.0;
    rho_new[e] = rn;
    const double d = fabs(rn - rh);
    v[0] = (d > v[0] || d != d) ? d : v[0];
  }
  if (grid_reduce<1, true>(v, partials, ticket, out) && threadIdx.x == 0 && !defer) oc->change = v[0];
}

// single-GPU path: the two kernels above in one pass (rho_new, the sum for gt and the max for change share the candidate)
__global__ void __launch_bounds__(ET) k_oc_finish(const double *__restrict__ rho, const double *__restrict__ dc,
                                                  double *__restrict__ rho_new, int64_t ne, double move, double *partials,
                                                  unsigned int *ticket, double *out, OcState *oc) {
  const double lmid = oc->lmid;
  const int steps = oc->steps;
  double sum = 0.0, mx = 0.0;                  // sum(rho_new - rho) for gt, max|rho_new - rho| for change
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x) {
    const double rh = rho[e];
    const double cand = steps > 0 ? oc_candidate(rh, dc[e], lmid, move) : 0.0;
    rho_new[e] = cand;                         // zero bisection steps: the reference returns its zero-initialised rho_new
    sum += cand - rh;
    const double d = fabs(cand - rh);
    mx = (d > mx || d != d) ? d : mx;
  }
  // two reductions with different operators: the sum first, then the max through the same deterministic machinery
  double s[1] = {sum};
  const bool last_sum = grid_reduce<1>(s, partials, ticket, out);
  if (last_sum && threadIdx.x == 0 && steps > 0) oc->gt = oc->g + s[0];
  double m[1] = {mx};
  const bool last_max = grid_reduce<1, true>(m, partials + TOPO_MAX_PARTIAL_BLOCKS, ticket + 1, out + 1);
  if (last_max && threadIdx.x == 0) oc->change = m[0];
}

// row-slab mode: decisions from the all-reduced sums in red[]
__global__ void k_oc_decide(OcState *oc, const double *red, double margin, int kind, double probe_base) {
  if (kind == 0 || kind == 4) {        // fast pass / probe pass (16 sums)
    if (oc->done || oc->need_exact) return;
    double v[OC_CAND + 1];
    for (int j = 0; j <= OC_CAND; ++j) v[j] = red[j];
    if (kind == 4) oc_decide_probe(oc, v, margin, probe_base);
    else oc_decide_multi(oc, v, margin);
  } else if (kind == 1) {              // exact step
    if (oc->done || !oc->need_exact) return;
    oc_decide_step(oc, red[0]);
  } else if (kind == 2) {              // final gt
    if (oc->steps > 0) oc->gt = oc->g + red[0];
  } else {                             // change (max-reduced)
    oc->change = red[0];
  }
}

}  // namespace

extern "C" int topo_simp_scale(topo_t *h, double penal, double emin, double emax) {
  if (!h) return TOPO_EINVAL;
  CUDA_TRY(cudaSetDevice(h->device));
  KL(h, "k_simp_scale", k_simp_scale<<<el_blocks(h->ne), ET, 0, h->stream>>>(h->rho, h->scale, h->ne, penal, emin, emax));
  CUDA_TRY(cudaGetLastError());
  h->diag_valid = false;
  h->gmg_valid = false;
  return TOPO_OK;
}

extern "C" int topo_simp_sensitivity(topo_t *h, double penal, double emin, double emax, double *objective) {
  if (!h) return TOPO_EINVAL;
  CUDA_TRY(cudaSetDevice(h->device));
  {
    int rank = 0;                                // rows of the pivoted Cholesky factor (zero rows beyond the rank)
    for (int k = 0; k < 8; ++k) {
      bool nz = false;
      for (int i = 0; i < 8; ++i) nz = nz || h->ke_factor.k[8 * k + i] != 0.0;
      if (nz) rank = k + 1;
    }
    const int bx = div_up(h->nx, ET);
    const int rows = std::max(8, div_up(h->ny, std::max(1, 4096 / bx)));
    dim3 grid(bx, div_up(h->ny, rows));
    KL(h, "k_simp_sens", k_simp_sens_march<<<grid, ET, 0, h->stream>>>(h->ke_factor, rank, h->u, h->rho, h->scale, h->dc, h->nx, h->ny, rows,
                                                                        penal, emin, emax, h->partials, h->ticket, h->red_out));
  }
  CUDA_TRY(cudaGetLastError());
  if (objective) {
    CUDA_TRY(cudaMemcpyAsync(objective, h->red_out, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
  }
  return TOPO_OK;
}

extern "C" int topo_filter_sensitivity(topo_t *h, double rmin, double dx, double dy) {
  if (!h || !(rmin > 0.0) || !(dx > 0.0) || !(dy > 0.0)) return TOPO_EINVAL;
  CUDA_TRY(cudaSetDevice(h->device));
  FilterTaps taps;
  taps.hx = taps.hy = 0;                        // largest offsets that still carry a positive weight
  while (taps.hx <= FH_MAX && rmin - (taps.hx + 1) * dx > 0.0) taps.hx++;
  while (taps.hy <= FH_MAX && rmin - (taps.hy + 1) * dy > 0.0) taps.hy++;
  if (taps.hx > FH_MAX || taps.hy > FH_MAX) {
    topo_set_error("filter radius %.3g spans more than %d elements", rmin, FH_MAX);
    return TOPO_EINVAL;
  }
  const int wrow = 2 * taps.hx + 1;
  for (int dj = -taps.hy; dj <= taps.hy; ++dj)
    for (int di = -taps.hx; di <= taps.hx; ++di)
      taps.w[(dj + taps.hy) * wrow + di + taps.hx] = rmin - std::sqrt((di * dx) * (di * dx) + (dj * dy) * (dj * dy));
  dim3 grid(div_up(h->nx, FT), div_up(h->ny, FT)), block(32, 8);
  const size_t smem = sizeof(double) * (FT + 2 * taps.hx) * (FT + 2 * taps.hy);
  const double *glo = nullptr, *ghi = nullptr;
  if (h->dist.on) {
    if (taps.hy != h->dist.filt_hy) {
      topo_set_error("row-slab filter: ghost rows were packed for hy=%d, the filter needs %d", h->dist.filt_hy, taps.hy);
      return TOPO_ESTATE;
    }
    glo = h->dist.has_lo ? h->dist.filt_recv_lo : nullptr;
    ghi = h->dist.has_hi ? h->dist.filt_recv_hi : nullptr;
  }
  constexpr int FB = 4;
  dim3 gsq(div_up(h->nx, FT), div_up(h->ny, 8 * FB));
  if (taps.hx == taps.hy && taps.hx == 3)
    KL(h, "k_filter", k_filter_sq<3, FB><<<gsq, block, 0, h->stream>>>(taps, h->rho, h->dc, h->dc2, h->nx, h->ny, glo, ghi));
  else if (taps.hx == taps.hy && taps.hx == 2)
    KL(h, "k_filter", k_filter_sq<2, FB><<<gsq, block, 0, h->stream>>>(taps, h->rho, h->dc, h->dc2, h->nx, h->ny, glo, ghi));
  else if (taps.hx == taps.hy && taps.hx == 1)
    KL(h, "k_filter", k_filter_sq<1, FB><<<gsq, block, 0, h->stream>>>(taps, h->rho, h->dc, h->dc2, h->nx, h->ny, glo, ghi));
  else
    KL(h, "k_filter", k_filter<<<grid, block, smem, h->stream>>>(taps, h->rho, h->dc, h->dc2, h->nx, h->ny, glo, ghi));
  CUDA_TRY(cudaGetLastError());
  std::swap(h->dc, h->dc2);
  return TOPO_OK;
}

// row-slab mode: rho*dc of the first / last hy element rows -> send buffers (host: exchange TOPO_X_FILTER, then
// topo_filter_sensitivity)
namespace {
__global__ void k_filter_pack(const double *__restrict__ rho, const double *__restrict__ dc, double *__restrict__ lo,
                              double *__restrict__ hi, int nx, int ny, int hy) {
  const int64_t n = (int64_t)hy * nx;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    lo[t] = rho[t] * dc[t];
    const size_t e = (size_t)(ny - hy) * nx + t;
    hi[t] = rho[e] * dc[e];
  }
}
}  // namespace
extern "C" int topo_dist_filter_pack(topo_t *h, double rmin, double dy) {
  if (!h || !h->dist.on || !(rmin > 0.0) || !(dy > 0.0)) return TOPO_EINVAL;
  CUDA_TRY(cudaSetDevice(h->device));
  int hy = 0;
  while (hy <= FH_MAX && rmin - (hy + 1) * dy > 0.0) hy++;
  if (hy > FH_MAX || hy > h->ny) {
    topo_set_error("row-slab filter: radius %.3g spans %d element rows (max %d, slab height %d)", rmin, hy, FH_MAX, h->ny);
    return TOPO_EINVAL;
  }
  h->dist.filt_hy = hy;
  if (hy > 0)
    KL(h, "k_filter_pack", k_filter_pack<<<el_blocks((int64_t)hy * h->nx), ET, 0, h->stream>>>(h->rho, h->dc, h->dist.filt_send_lo,
                                                                                             h->dist.filt_send_hi, h->nx, h->ny, hy));
  CUDA_TRY(cudaGetLastError());
  return TOPO_OK;
}

extern "C" int topo_oc_update(topo_t *h, double move, double *g, double *change, int *nsteps) {
  if (!h || !g) return TOPO_EINVAL;
  if (h->dist.on) {
    topo_set_error("row-slab handle: the OC bisection is driven through topo_dist_oc");
    return TOPO_ESTATE;
  }
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  const int nb = el_blocks(h->ne);
  CUDA_TRY(cudaEventRecord(h->ev0, st));
  KL(h, "k_oc_init", k_oc_init<<<1, 1, 0, st>>>(h->oc, *g));
  double *amp = h->sens;                   // scratch: SIMP does not use the BESO sensitivity field
  KL(h, "k_oc_prep", k_oc_prep<<<nb, ET, 0, st>>>(h->rho, h->dc, amp, h->ne));
  // rounding margin of the fast pass: (#elements) x (a few ulp of a value <= 1) with head-room
  const double margin = 64.0 * 2.220446049250313e-16 * (double)h->ne + 1e-12;
  OcState *hs = h->oc_host;
  hs->done = 0;
  int rounds = 0;
  const int hard_cap = 700;                // the bracket [0,1e9] underflows after ~1100 halvings = 275 rounds
  // with a multiplier from the previous design iteration the probe pass brackets the root within a factor 2 and
  // ~3 fast passes finish the bisection (instead of ~23): poll after 4 rounds, else after 12
  const bool probe = h->oc_lam_prev > 0.0 && !h->knobs.oc_no_probe;
  if (probe)
    KL(h, "k_oc_multi", k_oc_multi<<<std::min(nb, oc_multi_blocks(h)), ET, 0, st>>>(h->rho, amp, h->ne, move, margin, h->partials, h->ticket,
                                                                      h->red_out, h->oc, 0, h->oc_lam_prev));
  while (rounds < hard_cap) {
    const int batch = probe ? 4 : 12;
    for (int i = 0; i < batch; ++i) {
      KL(h, "k_oc_multi", k_oc_multi<<<std::min(nb, oc_multi_blocks(h)), ET, 0, st>>>(h->rho, amp, h->ne, move, margin, h->partials, h->ticket,
                                                        h->red_out, h->oc, 0, 0.0));
      KL(h, "k_oc_step", k_oc_step<<<nb, ET, 0, st>>>(h->rho, h->dc, h->ne, move, 1, h->partials, h->ticket,
                                                      h->red_out, h->oc, 0));
    }
    rounds += batch;
    CUDA_TRY(cudaMemcpyAsync(hs, h->oc, sizeof(OcState), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (hs->done) break;
  }
  KL(h, "k_oc_finish", k_oc_finish<<<nb, ET, 0, st>>>(h->rho, h->dc, h->rho_new, h->ne, move, h->partials, h->ticket,
                                                      h->red_out, h->oc));
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(hs, h->oc, sizeof(OcState), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaEventRecord(h->ev1, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaEventElapsedTime(&h->timing.oc_ms, h->ev0, h->ev1));
  std::swap(h->rho, h->rho_new);          // rho_new now holds the previous design
  h->oc_lam_prev = (hs->steps > 0 && hs->lmid > 0.0 && hs->lmid < 1e300) ? hs->lmid : 0.0;
  *g = hs->gt;
  if (change) *change = hs->change;
  if (nsteps) *nsteps = hs->steps;
  return TOPO_OK;
}

// ---- row-slab mode: the OC bisection in pieces around the host's all-reduces --------------------------------------
// The sums of every pass are left per rank in red_out (TOPO_S_RED); after the all-reduce the decision kernel
// runs on every rank with identical inputs, so all ranks walk the same bisection.
//   op 0  begin(g = value)                       -> nothing to reduce
//   op 1  fast pass                               -> all-reduce SUM 16, then op 2
//   op 2  decide fast pass
//   op 3  exact step (only if requested)          -> all-reduce SUM 1, then op 4
//   op 4  decide exact step
//   op 5  final gt                                -> all-reduce SUM 1, then op 6
//   op 6  decide final gt
//   op 7  apply (rho_new, change partial)         -> all-reduce MAX 1, then op 8
//   op 8  set change, swap rho <-> rho_new
//   op 9  probe pass around value = the previous multiplier -> all-reduce SUM 16, then op 10
//   op 10 tighten the bracket from the probe sums (value as in op 9)
// ne_global sizes the rounding margin of the fast pass like the single-GPU path.
extern "C" int topo_dist_oc(topo_t *h, int op, double move, double value, long long ne_global) {
  if (!h || !h->dist.on) return TOPO_EINVAL;
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  const int nb = el_blocks(h->ne);
  double *amp = h->sens;
  const double margin = 64.0 * 2.220446049250313e-16 * (double)ne_global + 1e-12;
  switch (op) {
    case 0:
      KL(h, "k_oc_init", k_oc_init<<<1, 1, 0, st>>>(h->oc, value));
      KL(h, "k_oc_prep", k_oc_prep<<<nb, ET, 0, st>>>(h->rho, h->dc, amp, h->ne));
      break;
    case 1:
      KL(h, "k_oc_multi", k_oc_multi<<<std::min(nb, oc_multi_blocks(h)), ET, 0, st>>>(h->rho, amp, h->ne, move, margin, h->partials, h->ticket,
                                                                        h->red_out, h->oc, 1, 0.0));
      break;
    case 2: KL(h, "k_oc_decide", k_oc_decide<<<1, 1, 0, st>>>(h->oc, h->red_out, margin, 0, 0.0)); break;
    case 9:                                   // probe pass around `value` (the previous multiplier) -> all-reduce 16, op 10
      KL(h, "k_oc_multi", k_oc_multi<<<std::min(nb, oc_multi_blocks(h)), ET, 0, st>>>(h->rho, amp, h->ne, move, margin, h->partials, h->ticket,
                                                                        h->red_out, h->oc, 1, value));
      break;
    case 10: KL(h, "k_oc_decide", k_oc_decide<<<1, 1, 0, st>>>(h->oc, h->red_out, margin, 4, value)); break;
    case 3:
      KL(h, "k_oc_step", k_oc_step<<<nb, ET, 0, st>>>(h->rho, h->dc, h->ne, move, 1, h->partials, h->ticket, h->red_out, h->oc, 1));
      break;
    case 4: KL(h, "k_oc_decide", k_oc_decide<<<1, 1, 0, st>>>(h->oc, h->red_out, margin, 1, 0.0)); break;
    case 5:
      CUDA_TRY(cudaMemsetAsync(h->red_out, 0, sizeof(double), st));      // zero steps: the kernel returns early
      KL(h, "k_oc_final_gt", k_oc_final_gt<<<nb, ET, 0, st>>>(h->rho, h->dc, h->ne, move, h->partials, h->ticket, h->red_out, h->oc, 1));
      break;
    case 6: KL(h, "k_oc_decide", k_oc_decide<<<1, 1, 0, st>>>(h->oc, h->red_out, margin, 2, 0.0)); break;
    case 7:
      KL(h, "k_oc_apply", k_oc_apply<<<nb, ET, 0, st>>>(h->rho, h->dc, h->rho_new, h->ne, move, h->partials, h->ticket, h->red_out,
                                                        h->oc, 1));
      break;
    case 8:
      KL(h, "k_oc_decide", k_oc_decide<<<1, 1, 0, st>>>(h->oc, h->red_out, margin, 3, 0.0));
      std::swap(h->rho, h->rho_new);
      break;
    default:
      return TOPO_EINVAL;
  }
  CUDA_TRY(cudaGetLastError());
  return TOPO_OK;
}

extern "C" int topo_dist_oc_state(topo_t *h, int *done, int *need_exact, int *steps, double *gt, double *change,
                                  double *lmid) {
  if (!h || !h->dist.on) return TOPO_EINVAL;
  CUDA_TRY(cudaSetDevice(h->device));
  OcState *hs = h->oc_host;
  CUDA_TRY(cudaMemcpyAsync(hs, h->oc, sizeof(OcState), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (done) *done = hs->done;
  if (need_exact) *need_exact = hs->need_exact;
  if (steps) *steps = hs->steps;
  if (gt) *gt = hs->gt;
  if (change) *change = hs->change;
  if (lmid) *lmid = hs->lmid;
  return TOPO_OK;
}
