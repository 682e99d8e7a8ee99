// The multigrid "tail": every level from the first with <= tail_max_nodes nodes down to the dense coarsest solve in ONE
// launch of a thread-block cluster, with NOTHING but registers, shared memory and distributed shared memory between
// the right-hand side of the first tail level and its result.  No counterpart in the reference (SuperLU,
// optimize.py:437).
//
// Why this shape (measured, profiles/r2_tail.md): the generic tail (k_tail, gmg_cycle.cu) hands data between CTAs
// through global memory, so every stage pays store -> MEMBAR.ALL.GPU -> barrier.cluster -> CCTL.IVALL (L1 flushed) ->
// L2 loads of operator, iterate AND of every spilled register / stack slot: ~2 us of fixed latency per stage whatever
// the level size (a bare cluster barrier is 0.28 us, a neighbour hand-over with st.async + mbarrier 0.35 us --
// scripts/micro/cluster_sync.cu), 77 us per V-cycle for levels of <= 8385 nodes.  Here
//   * a level is cut into blocks of consecutive nodes, one block per CTA, ONE NODE PER THREAD; the node's nine 2x2
//     stencil blocks (fp32, as stored), w*D^-1, b and the fp64 iterate stay in registers for a whole smoothing phase;
//   * the iterate travels as an fp32 copy in shared memory; a thread whose node lies within one node row of its
//     block's end also sends its value into the neighbouring CTA's halo with st.async, which credits that CTA's
//     mbarrier -- a sweep is 9 LDS.64 + 36 FFMA + the fp64 update + one neighbour hand-over, no fence, no
//     cluster-wide rendezvous, no global memory;
//   * residuals, iterates and right-hand sides of all tail levels live in shared memory; restriction and prolongation
//     read them from the owner CTA (ld.shared::cluster); a cluster barrier separates the phases only;
//   * levels of <= T2 nodes are done by CTA 0 alone with __syncthreads, the coarsest by its dense inverse;
//   * the kernel has no stack frame and no spills (ptxas -v), so the L1 flush of a cluster barrier costs nothing.
// Products A x are formed in fp32 -- the stencils are stored in fp32 anyway -- iterate, right-hand side and residual stay
// fp64 (same arithmetic as the register-resident small levels of round 1; identical PCG iteration counts).
#include <algorithm>

#include <cooperative_groups.h>

#include "topo_internal.cuh"
#include "gmg_common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int T2 = 640;                 // threads per CTA = nodes per block at most (96 registers per thread)
constexpr int T2_XS = 3 * T2;           // float2 per iterate buffer: block + a halo of <= one block on each side
constexpr int T2_ARENA = 2048;          // double2 per CTA for the iterates (and as many for the right-hand sides) of all levels
constexpr int T2_MAXLV = 12;
constexpr int T2_DENSE = 96;            // largest dense coarsest system
constexpr size_t T2_SMEM = 16 + sizeof(float2) * 2 * T2_XS + sizeof(double2) * 2 * T2 + sizeof(double2) * 2 * T2_ARENA +
                           sizeof(double) * 2 * T2_DENSE;

struct Tail2Level {
  const float4 *st;          // 9 planes x nn
  const double2 *dinv;
  double2 *x;                // global result (written for the first tail level only)
  const double2 *b;          // global right-hand side (read for the first tail level in row-slab mode)
  int nx, ny;
  int B;                     // nodes per CTA (cluster levels: >= npr + 1; small levels: all nodes)
  unsigned magic;            // ceil(2^32 / B): owner(m) = umulhi(m, magic), checked exhaustively on the host
  int off;                   // offset of this level's block in the per-CTA arenas
  int nu;                    // smoothing sweeps before = after
};
struct Tail2Args {
  Tail2Level lv[T2_MAXLV];
  int n;                     // levels (the last one is the dense coarsest)
  int ncl;                   // leading levels worked on by the whole cluster
  const void *r_parent;      // residual of the level above (float2 when that is the fine grid); nullptr: lv[0].b is in place
  int parent_f32, pnx, pny;
  const double *coarse_inv;
  int coarse_n;
  double omega;
  const int *done;
  int *err;
  unsigned long long *dbg;
};

__device__ __forceinline__ unsigned long long t2_timer() {
  unsigned long long v;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
  return v;
}
#define T2_STAMP() do { if (a.dbg != nullptr && rank == 0 && t == 0) { a.dbg[1 + a.dbg[0]] = t2_timer(); a.dbg[0] += 1; } } while (0)

__device__ __forceinline__ unsigned t2_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned t2_mapa(unsigned local_addr, unsigned rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ double2 t2_ld_remote2(unsigned cluster_addr) {
  double2 v;
  asm volatile("ld.shared::cluster.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(cluster_addr) : "memory");
  return v;
}
__device__ __forceinline__ void t2_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

// geometry of one level as seen by this thread
struct T2Geo {
  int npr, nn, H, B;
  int count;        // nodes of this CTA
  int count_hi;     // nodes of the next CTA
  int n, loc;       // global node, index in the halo-extended iterate buffer
  bool active;
};
__device__ __forceinline__ T2Geo t2_geo(const Tail2Level &L, int rank, int t) {
  T2Geo g;
  g.npr = L.nx + 1;
  g.nn = g.npr * (L.ny + 1);
  g.H = g.npr + 1;
  g.B = L.B;
  g.count = min(max(g.nn - rank * g.B, 0), g.B);
  g.count_hi = min(max(g.nn - (rank + 1) * g.B, 0), g.B);
  g.n = rank * g.B + t;
  g.loc = t + g.H;
  g.active = t < g.count;
  return g;
}

// (A x)(node) from the fp32 copy of the iterate: all nine loads first, three partial sums per component
__device__ __forceinline__ void t2_apply(const float4 (&st)[9], const float2 *__restrict__ xs, int loc, int npr, float &ax,
                                         float &ay) {
  const float2 *p = xs + loc;
  const float2 v0 = p[-npr - 1], v1 = p[-npr], v2 = p[-npr + 1], v3 = p[-1], v4 = p[0], v5 = p[1], v6 = p[npr - 1],
               v7 = p[npr], v8 = p[npr + 1];
  float a0 = st[0].x * v0.x, b0 = st[0].z * v0.x;
  float a1 = st[1].x * v1.x, b1 = st[1].z * v1.x;
  float a2 = st[2].x * v2.x, b2 = st[2].z * v2.x;
  a0 = fmaf(st[0].y, v0.y, a0); b0 = fmaf(st[0].w, v0.y, b0);
  a1 = fmaf(st[1].y, v1.y, a1); b1 = fmaf(st[1].w, v1.y, b1);
  a2 = fmaf(st[2].y, v2.y, a2); b2 = fmaf(st[2].w, v2.y, b2);
  a0 = fmaf(st[3].x, v3.x, a0); b0 = fmaf(st[3].z, v3.x, b0);
  a1 = fmaf(st[4].x, v4.x, a1); b1 = fmaf(st[4].z, v4.x, b1);
  a2 = fmaf(st[5].x, v5.x, a2); b2 = fmaf(st[5].z, v5.x, b2);
  a0 = fmaf(st[3].y, v3.y, a0); b0 = fmaf(st[3].w, v3.y, b0);
  a1 = fmaf(st[4].y, v4.y, a1); b1 = fmaf(st[4].w, v4.y, b1);
  a2 = fmaf(st[5].y, v5.y, a2); b2 = fmaf(st[5].w, v5.y, b2);
  a0 = fmaf(st[6].x, v6.x, a0); b0 = fmaf(st[6].z, v6.x, b0);
  a1 = fmaf(st[7].x, v7.x, a1); b1 = fmaf(st[7].z, v7.x, b1);
  a2 = fmaf(st[8].x, v8.x, a2); b2 = fmaf(st[8].z, v8.x, b2);
  a0 = fmaf(st[6].y, v6.y, a0); b0 = fmaf(st[6].w, v6.y, b0);
  a1 = fmaf(st[7].y, v7.y, a1); b1 = fmaf(st[7].w, v7.y, b1);
  a2 = fmaf(st[8].y, v8.y, a2); b2 = fmaf(st[8].w, v8.y, b2);
  ax = (a0 + a1) + a2;
  ay = (b0 + b1) + b2;
}

struct T2Ctx {
  unsigned long long *bar;       // two mbarriers, one per iterate buffer
  float2 *xs0, *xs1;             // iterate copies incl. halos
  unsigned uses0, uses1;         // completed phases of each mbarrier
  int *err;
};

// own value into buffer `buf` (local + the neighbours' halos).  Cluster levels: st.async carries data and completion
// count in one DSMEM transaction; the receiver armed its mbarrier with the bytes it expects from both sides.
template <bool CL>
__device__ __forceinline__ void t2_publish(const T2Ctx &C, int buf, const T2Geo &g, int rank, int t, const double2 &x) {
  float2 *xs = buf ? C.xs1 : C.xs0;
  if (CL && t == 0) {
    const unsigned expect = 8u * (unsigned)(((rank > 0 && g.count > 0) ? g.H : 0) + min(g.H, g.count_hi));
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(t2_u32(C.bar + buf)), "r"(expect) : "memory");
  }
  if (!g.active) return;
  const float2 v = make_float2((float)x.x, (float)x.y);
  xs[g.loc] = v;
  if (CL) {
    const unsigned long long bits = ((unsigned long long)__float_as_uint(v.y) << 32) | (unsigned long long)__float_as_uint(v.x);
    if (rank > 0 && t < g.H)                                   // my node is in the halo above the previous block
      asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];\n" ::
                   "r"(t2_mapa(t2_u32(xs + g.B + g.H + t), rank - 1)), "l"(bits), "r"(t2_mapa(t2_u32(C.bar + buf), rank - 1)) : "memory");
    if (g.count_hi > 0 && t >= g.B - g.H)                      // ... in the halo below the next block
      asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];\n" ::
                   "r"(t2_mapa(t2_u32(xs + t - g.B + g.H), rank + 1)), "l"(bits), "r"(t2_mapa(t2_u32(C.bar + buf), rank + 1)) : "memory");
  }
}
// buffer `buf` is complete (own values and both halos) and every thread of the CTA may read it
template <bool CL>
__device__ __forceinline__ void t2_handover(T2Ctx &C, int buf) {
  if (CL) {
    const unsigned addr = t2_u32(C.bar + buf);
    const unsigned parity = (buf ? C.uses1 : C.uses0) & 1u;
    bool ok = false;
    for (unsigned spin = 0; spin < (1u << 20) && !ok; ++spin) {   // bounded: a protocol error must not hang the device
      unsigned r;
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                   : "=r"(r) : "r"(addr), "r"(parity) : "memory");
      ok = r != 0u;
    }
    if (!ok && threadIdx.x == 0) atomicExch(C.err, 1);
    if (buf) C.uses1 += 1; else C.uses0 += 1;
  }
  __syncthreads();
}

// node state of a smoothing phase
struct T2Node {
  float4 st[9];
  double2 wd;        // omega * D^-1 (0 at constrained DOFs)
  double2 b, x;
};
__device__ __forceinline__ void t2_load(const Tail2Level &L, const T2Geo &g, double omega, T2Node &N) {
#pragma unroll
  for (int k = 0; k < 9; ++k)                                  // planes of missing neighbours hold zeros (k_cells_to_stencil)
    N.st[k] = g.active ? __ldg(L.st + (size_t)k * g.nn + g.n) : make_float4(0.f, 0.f, 0.f, 0.f);
  const double2 d = g.active ? __ldg(L.dinv + g.n) : make_double2(0.0, 0.0);
  N.wd = make_double2(omega * d.x, omega * d.y);
}
template <bool CL>
__device__ __forceinline__ void t2_sweeps(T2Ctx &C, const T2Geo &g, int rank, int t, T2Node &N, int &cur, int nsweeps) {
  for (int s = 0; s < nsweeps; ++s) {
    if (g.active) {
      float ax, ay;
      t2_apply(N.st, cur ? C.xs1 : C.xs0, g.loc, g.npr, ax, ay);
      N.x.x = fma(N.wd.x, N.b.x - (double)ax, N.x.x);
      N.x.y = fma(N.wd.y, N.b.y - (double)ay, N.x.y);
    }
    t2_publish<CL>(C, cur ^ 1, g, rank, t, N.x);
    t2_handover<CL>(C, cur ^ 1);
    cur ^= 1;
  }
}
// pre-smoothing from a zero guess (nu sweeps, the first one pointwise), residual -> rs, iterate / right-hand side -> arenas
template <bool CL>
__device__ __forceinline__ void t2_down(const Tail2Args &a, int l, T2Ctx &C, int rank, int t, double2 b, double2 *rs,
                                        double2 *ax_, double2 *ab_) {
  const Tail2Level &L = a.lv[l];
  const T2Geo g = t2_geo(L, rank, t);
  if (g.count == 0) return;                                     // a CTA without nodes on this level
  T2Node N;
  t2_load(L, g, a.omega, N);
  N.b = b;
  N.x = make_double2(N.wd.x * b.x, N.wd.y * b.y);
  int cur = 0;
  t2_publish<CL>(C, 0, g, rank, t, N.x);
  t2_handover<CL>(C, 0);
  t2_sweeps<CL>(C, g, rank, t, N, cur, L.nu - 1);
  if (g.active) {
    float ax, ay;
    t2_apply(N.st, cur ? C.xs1 : C.xs0, g.loc, g.npr, ax, ay);
    rs[t] = make_double2(N.wd.x != 0.0 ? N.b.x - (double)ax : 0.0, N.wd.y != 0.0 ? N.b.y - (double)ay : 0.0);
    ax_[L.off + t] = N.x;
    ab_[L.off + t] = N.b;
  }
}
// owner CTA and index inside its block of node m on a level
__device__ __forceinline__ void t2_owner(const Tail2Level &L, int m, int &owner, int &local) {
  owner = (int)__umulhi((unsigned)m, L.magic);
  local = m - owner * L.B;
}
// (P^T r)(I, J): fine residual in the owners' shared memory (REMOTE: through the cluster address space)
template <bool REMOTE>
__device__ __forceinline__ double2 t2_restrict(const Tail2Level &F, const double2 *rs, int I, int J) {
  const int nprf = F.nx + 1;
  int X, Y;
  double wxl, wxr, wyl, wyr;
  restrict_taps(I, F.nx, X, wxl, wxr);
  restrict_taps(J, F.ny, Y, wyl, wyr);
  const double wx[3] = {wxl, 1.0, wxr}, wy[3] = {wyl, 1.0, wyr};
  const unsigned base = t2_u32(rs);
  double sx = 0.0, sy = 0.0;
#pragma unroll
  for (int dj = -1; dj <= 1; ++dj) {
#pragma unroll
    for (int di = -1; di <= 1; ++di) {
      const double w = wy[dj + 1] * wx[di + 1];
      if (w == 0.0) continue;
      const int m = (Y + dj) * nprf + (X + di);
      double2 v;
      if (REMOTE) {
        int owner, local;
        t2_owner(F, m, owner, local);
        v = t2_ld_remote2(t2_mapa(base + 16u * (unsigned)local, (unsigned)owner));
      } else {
        v = rs[m];
      }
      sx = fma(w, v.x, sx);
      sy = fma(w, v.y, sy);
    }
  }
  return make_double2(sx, sy);
}
// (P xc)(i, j): coarse iterate in the owners' arenas
template <bool REMOTE>
__device__ __forceinline__ double2 t2_prolong(const Tail2Level &Cl, const double2 *arena_x, int i, int j, int nxf, int nyf) {
  int I0, I1, J0, J1;
  double wx0, wx1, wy0, wy1;
  prolong_taps(i, nxf, I0, I1, wx0, wx1);
  prolong_taps(j, nyf, J0, J1, wy0, wy1);
  const int nprc = Cl.nx + 1;
  const unsigned base = t2_u32(arena_x + Cl.off);
  auto at = [&](int J, int I) -> double2 {
    const int m = J * nprc + I;
    if (REMOTE) {
      int owner, local;
      t2_owner(Cl, m, owner, local);
      return t2_ld_remote2(t2_mapa(base + 16u * (unsigned)local, (unsigned)owner));
    }
    return arena_x[Cl.off + m];
  };
  const double2 v00 = at(J0, I0);
  double px = wy0 * wx0 * v00.x, py = wy0 * wx0 * v00.y;
  if (wx1 != 0.0) {
    const double2 v = at(J0, I1);
    px = fma(wy0 * wx1, v.x, px); py = fma(wy0 * wx1, v.y, py);
  }
  if (wy1 != 0.0) {
    const double2 v = at(J1, I0);
    px = fma(wy1 * wx0, v.x, px); py = fma(wy1 * wx0, v.y, py);
    if (wx1 != 0.0) {
      const double2 w = at(J1, I1);
      px = fma(wy1 * wx1, w.x, px); py = fma(wy1 * wx1, w.y, py);
    }
  }
  return make_double2(px, py);
}
// coarse correction + nu post-smoothing sweeps; the result goes to the arena (and to global memory for level 0)
template <bool CL>
__device__ __forceinline__ void t2_up(const Tail2Args &a, int l, T2Ctx &C, int rank, int t, double2 *ax_, const double2 *ab_) {
  const Tail2Level &L = a.lv[l];
  const T2Geo g = t2_geo(L, rank, t);
  if (g.count == 0) return;
  T2Node N;
  t2_load(L, g, a.omega, N);
  N.b = make_double2(0.0, 0.0);
  N.x = make_double2(0.0, 0.0);
  if (g.active) {
    N.b = ab_[L.off + t];
    N.x = ax_[L.off + t];
    const int r = g.n / g.npr, c = g.n - r * g.npr;
    const double2 p = CL ? t2_prolong<true>(a.lv[l + 1], ax_, c, r, L.nx, L.ny) : t2_prolong<false>(a.lv[l + 1], ax_, c, r, L.nx, L.ny);
    if (N.wd.x != 0.0) N.x.x += p.x;
    if (N.wd.y != 0.0) N.x.y += p.y;
  }
  int cur = 0;
  t2_publish<CL>(C, 0, g, rank, t, N.x);
  t2_handover<CL>(C, 0);
  t2_sweeps<CL>(C, g, rank, t, N, cur, L.nu);
  if (g.active) {
    ax_[L.off + t] = N.x;
    if (l == 0) L.x[g.n] = N.x;
  }
}

__global__ void __launch_bounds__(T2, 1) k_tail2(const __grid_constant__ Tail2Args a) {
  if (a.done && *a.done) return;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int t = threadIdx.x;
  extern __shared__ __align__(16) unsigned char t2_raw[];
  T2Ctx C;
  C.bar = reinterpret_cast<unsigned long long *>(t2_raw);
  C.xs0 = reinterpret_cast<float2 *>(t2_raw + 16);
  C.xs1 = C.xs0 + T2_XS;
  C.uses0 = C.uses1 = 0u;
  C.err = a.err;
  double2 *rs0 = reinterpret_cast<double2 *>(C.xs1 + T2_XS);       // residuals, double-buffered by level parity
  double2 *rs1 = rs0 + T2;
  double2 *ar_x = rs1 + T2;                                         // iterates of all levels (this CTA's nodes)
  double2 *ar_b = ar_x + T2_ARENA;                                  // right-hand sides
  double *dense_b = reinterpret_cast<double *>(ar_b + T2_ARENA);    // coarsest right-hand side
  // halo entries outside the domain are read with zero coefficients and must hold finite numbers
  for (int i = t; i < 2 * T2_XS; i += T2) C.xs0[i] = make_float2(0.f, 0.f);
  if (t == 0) {                                                     // one pending arrival each: the CTA's own expect_tx
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(t2_u32(C.bar)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(t2_u32(C.bar + 1)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    if (a.dbg != nullptr && rank == 0) a.dbg[0] = 0;
  }
  // right-hand side of this thread's node on the first tail level (global loads: issued before the first barrier)
  const int n_levels = a.n, ncl = a.ncl;
  double2 b = make_double2(0.0, 0.0);
  {
    const Tail2Level &L0 = a.lv[0];
    const T2Geo g = t2_geo(L0, rank, t);
    if (g.active && (ncl > 0 || rank == 0)) {
      const int r = g.n / g.npr, c = g.n - r * g.npr;
      if (a.r_parent == nullptr) b = L0.b[g.n];
      else if (a.parent_f32) b = restrict_node(reinterpret_cast<const float2 *>(a.r_parent), c, r, a.pnx, a.pny);
      else b = restrict_node(reinterpret_cast<const double2 *>(a.r_parent), c, r, a.pnx, a.pny);
    }
  }
  t2_cluster_sync();            // nobody pushes into a block that is still being cleared / whose barriers do not exist yet
  T2_STAMP();
  // ---- cluster levels, downwards ----
  for (int l = 0; l < ncl; ++l) {
    double2 *rs = (l & 1) ? rs1 : rs0;
    t2_down<true>(a, l, C, rank, t, b, rs, ar_x, ar_b);
    t2_cluster_sync();
    T2_STAMP();
    b = make_double2(0.0, 0.0);
    if (l + 1 < ncl || rank == 0) {                                  // owners of level l+1 (a small level lives on CTA 0)
      const Tail2Level &Cn = a.lv[l + 1];
      const T2Geo gc = t2_geo(Cn, rank, t);
      if (gc.active) {
        const int J = gc.n / gc.npr, I = gc.n - J * gc.npr;
        b = t2_restrict<true>(a.lv[l], rs, I, J);
      }
    }
  }
  // ---- small levels and the dense coarsest solve: CTA 0 alone ----
  if (rank == 0) {
    for (int l = ncl; l + 1 < n_levels; ++l) {
      double2 *rs = (l & 1) ? rs1 : rs0;
      t2_down<false>(a, l, C, 0, t, b, rs, ar_x, ar_b);
      __syncthreads();
      T2_STAMP();
      const Tail2Level &Cn = a.lv[l + 1];
      const T2Geo gc = t2_geo(Cn, 0, t);
      b = make_double2(0.0, 0.0);
      if (gc.active) {
        const int J = gc.n / gc.npr, I = gc.n - J * gc.npr;
        b = t2_restrict<false>(a.lv[l], rs, I, J);
      }
    }
    {  // coarsest: x = mask(inv b); four lanes per row
      const Tail2Level &K = a.lv[n_levels - 1];
      const int nnK = (K.nx + 1) * (K.ny + 1), N = a.coarse_n;
      if (t < nnK) { dense_b[2 * t] = b.x; dense_b[2 * t + 1] = b.y; }
      __syncthreads();
      const int row = t >> 2, part = t & 3;
      double acc = 0.0;
      if (row < N) {
        const double *ir = a.coarse_inv + (size_t)row * N;
        for (int j = part; j < N; j += 4) acc = fma(__ldg(ir + j), dense_b[j], acc);
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      if (row < N && part == 0) {
        const double d = reinterpret_cast<const double *>(K.dinv)[row];
        const double v = d != 0.0 ? acc : 0.0;
        reinterpret_cast<double *>(ar_x + K.off)[row] = v;
        if (n_levels == 1) reinterpret_cast<double *>(K.x)[row] = v;
      }
      __syncthreads();
      T2_STAMP();
    }
    for (int l = n_levels - 2; l >= ncl; --l) {
      t2_up<false>(a, l, C, 0, t, ar_x, ar_b);
      __syncthreads();
      T2_STAMP();
    }
  }
  // ---- cluster levels, upwards ----
  for (int l = ncl - 1; l >= 0; --l) {
    t2_cluster_sync();          // the iterate of level l+1 is complete in its owners' arenas
    T2_STAMP();
    t2_up<true>(a, l, C, rank, t, ar_x, ar_b);
  }
  T2_STAMP();
  if (ncl > 0) t2_cluster_sync();   // no CTA exits while a peer may still read its shared memory
}

// ---- host side: the plan ------------------------------------------------------------------------------------------
struct Tail2Plan {
  bool ok;
  int cs;                  // cluster size
  Tail2Args args;
};

int t2_cluster_size(const topo_handle *h) {
  static int cs_dev[TOPO_MAX_DEVICES] = {};
  int &cs = cs_dev[topo_dev_slot(h)];
  if (cs == 0) {
    cs = -1;
    if (cudaFuncSetAttribute(k_tail2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T2_SMEM) == cudaSuccess &&
        cudaFuncSetAttribute(k_tail2, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
      for (int c : {16, 8, 4, 2, 1}) {
        if (c > h->knobs.tail_cluster_max) continue;
        cudaLaunchConfig_t q = {};
        q.gridDim = dim3(c); q.blockDim = dim3(T2); q.dynamicSmemBytes = T2_SMEM;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = c; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        q.attrs = at; q.numAttrs = 1;
        int ncl = 0;
        if (cudaOccupancyMaxActiveClusters(&ncl, k_tail2, &q) == cudaSuccess && ncl >= 1) { cs = c; break; }
      }
    }
    cudaGetLastError();
  }
  return cs > 0 ? cs : 0;
}

// Does the tail starting at level `tail` fit the shared-memory kernel?  (One node per thread on every level, a halo of
// at most one block, all iterates in the arena, owner(m) exact through the multiply-high.)
bool t2_plan(const topo_handle *h, int tail, Tail2Plan &P) {
  P.ok = false;
  const int L = h->n_levels;
  const int n = L - tail;
  if (h->knobs.tail_kind != 2 || n < 1 || n > T2_MAXLV || h->coarse_n > T2_DENSE || 4 * h->coarse_n > 4 * T2) return false;
  const int cs = t2_cluster_size(h);
  if (cs < 1) return false;
  Tail2Args &A = P.args;
  A = Tail2Args{};
  A.n = n;
  int ncl = 0, off = 0;
  for (int k = 0; k < n; ++k) {
    const GmgLevel &G = h->levels[tail + k];
    Tail2Level &T = A.lv[k];
    T.st = reinterpret_cast<const float4 *>(G.st);
    T.dinv = G.dinv; T.x = G.x; T.b = G.b; T.nx = G.nx; T.ny = G.ny;
    const int npr = G.nx + 1, nn = (int)G.nn, H = npr + 1;
    const bool last = (k == n - 1);
    if (nn > T2 && !last) {                                  // cluster level
      if (k != ncl) return false;                            // cluster levels must lead
      ++ncl;
      T.B = std::max((nn + cs - 1) / cs, H);
      if (T.B > T2 || (nn + T.B - 1) / T.B > cs) return false;
    } else {
      if (nn > T2) return false;
      T.B = nn;
    }
    if (T.B + 2 * H > T2_XS) return false;
    T.magic = (unsigned)((0x100000000ull + (unsigned long long)T.B - 1) / (unsigned long long)T.B);
    for (int m = 0; m < nn; ++m)
      if ((int)(((unsigned long long)(unsigned)m * T.magic) >> 32) != m / T.B) return false;
    T.off = off;
    off += T.B;
    T.nu = last ? 0 : ((nn > T2) ? h->nu_tail_big : h->nu_tail_small);
    if (!last && T.nu < 1) return false;
  }
  if (off > T2_ARENA) return false;
  if (ncl > 0 && cs < 2) return false;
  A.ncl = ncl;
  P.cs = ncl > 0 ? cs : 1;
  P.ok = true;
  return true;
}

}  // namespace

// the plan is a function of the hierarchy only: computed once per handle and tail start, dropped by topo_gmg_free
static const Tail2Plan *t2_cached_plan(const topo_handle *hc, int tail) {
  topo_handle *h = const_cast<topo_handle *>(hc);
  if (h->tail2_cache == nullptr || h->tail2_cache_tail != tail) {
    if (h->tail2_cache == nullptr) h->tail2_cache = new Tail2Plan();
    t2_plan(h, tail, *static_cast<Tail2Plan *>(h->tail2_cache));
    h->tail2_cache_tail = tail;
  }
  return static_cast<const Tail2Plan *>(h->tail2_cache);
}

void topo_tail2_forget(topo_handle *h) {
  delete static_cast<Tail2Plan *>(h->tail2_cache);
  h->tail2_cache = nullptr;
  h->tail2_cache_tail = -1;
}

// true: the tail starting at `tail` runs in k_tail2 (its result is left in levels[tail].x)
bool topo_tail2_applicable(const topo_handle *h, int tail) { return t2_cached_plan(h, tail)->ok; }

int topo_tail2_launch(topo_handle *h, int tail, const double2 *r_parent, int pnx, int pny, int parent_f32) {
  const Tail2Plan *P = t2_cached_plan(h, tail);
  if (!P->ok) {
    topo_set_error("multigrid tail: the shared-memory plan does not fit this hierarchy");
    return TOPO_ESTATE;
  }
  Tail2Args A = P->args;
  A.r_parent = r_parent; A.parent_f32 = parent_f32; A.pnx = pnx; A.pny = pny;
  A.coarse_inv = h->coarse_inv;
  A.coarse_n = h->coarse_n;
  A.omega = h->omega;
  A.done = &h->sc->done;
  A.err = &h->sc->tail_error;
  A.dbg = h->tail_dbg;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(P->cs);
  cfg.blockDim = dim3(T2);
  cfg.dynamicSmemBytes = T2_SMEM;
  cfg.stream = h->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = P->cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  KL(h, "k_tail2", CUDA_TRY(cudaLaunchKernelEx(&cfg, k_tail2, A)));
  return TOPO_OK;
}
