// Microbenchmark behind the next step for the multigrid tail ("what a cluster stage costs"):
// k_tail and k_tail_ds pay ~2.5 us per stage of a cluster level, and moving the inter-CTA DATA into distributed
// shared memory did not change that -- so the cost is the cluster-wide barrier.  This program times, on one cluster of
// 16 CTAs x 768 threads (the tail's launch shape), one "stage hand-over" done three ways:
//   (A) cooperative_groups cluster.sync()            (barrier.cluster arrive.release + wait.acquire: MEMBAR.ALL.GPU)
//   (B) barrier.cluster.arrive.relaxed + wait        (no memory ordering: lower bound of the barrier hardware itself)
//   (C) neighbour-to-neighbour hand-over: the first NPR threads of a CTA send 8 bytes each into the halo rows of the
//       two neighbouring CTAs with st.async (data + mbarrier complete_tx in one DSMEM transaction); a CTA waits on its
//       own mbarrier for the bytes of both neighbours, then __syncthreads.  No fence, no cluster-wide rendezvous --
//       the pattern a row-block smoother needs (a block only reads its two neighbours' halo rows).
// (C) checks the payload (every value carries the iteration number and the sender's rank).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cluster_sync cluster_sync.cu        Run: ./cluster_sync [iterations]
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e__)); exit(1); } } while (0)

constexpr int THREADS = 768, CLUSTER = 16, NPR = 129;     // 129 nodes per row: level 4 of the 2048x1024 hierarchy

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long v;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
  return v;
}

// (A) ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS) k_cg_sync(int iters, unsigned long long *out) {
  cg::cluster_group cl = cg::this_cluster();
  cl.sync();
  const unsigned long long t0 = gtimer();
  for (int i = 0; i < iters; ++i) cl.sync();
  const unsigned long long t1 = gtimer();
  if (cl.block_rank() == 0 && threadIdx.x == 0) out[0] = t1 - t0;
}

// (B) ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS) k_relaxed(int iters, unsigned long long *out) {
  cg::cluster_group cl = cg::this_cluster();
  cl.sync();
  const unsigned long long t0 = gtimer();
  for (int i = 0; i < iters; ++i) {
    asm volatile("barrier.cluster.arrive.relaxed.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.aligned;\n" ::: "memory");
  }
  const unsigned long long t1 = gtimer();
  if (cl.block_rank() == 0 && threadIdx.x == 0) out[0] = t1 - t0;
}

// (C) ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
// shared::cluster address of `local_addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ unsigned mapa(unsigned local_addr, unsigned rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded: a protocol error must end the kernel (flag + early exit), never hang or fault the device
__device__ __forceinline__ bool mbar_wait(unsigned long long *bar, unsigned parity) {
  const unsigned addr = smem_u32(bar);
  for (unsigned spin = 0; spin < (1u << 22); ++spin) {
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}
// 8 bytes into a peer's shared memory; the peer's mbarrier is credited with them when they have landed
__device__ __forceinline__ void st_async_b64(unsigned remote_addr, unsigned long long v, unsigned remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];\n" ::"r"(remote_addr), "l"(v),
               "r"(remote_bar) : "memory");
}

__global__ void __launch_bounds__(THREADS) k_neighbour(int iters, unsigned long long *out, int *errors) {
  cg::cluster_group cl = cg::this_cluster();
  const int rank = (int)cl.block_rank(), CS = (int)cl.num_blocks(), t = threadIdx.x;
  __shared__ __align__(8) unsigned long long bar[2];                  // one per iteration parity
  __shared__ __align__(8) unsigned long long halo[2][2][NPR];         // [parity][0 = from below, 1 = from above][column]
  const bool has_lo = rank > 0, has_hi = rank + 1 < CS;
  const unsigned expect = 8u * NPR * ((has_lo ? 1u : 0u) + (has_hi ? 1u : 0u));
  if (t == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  cl.sync();                                                          // every CTA's barriers exist before anyone sends
  int bad = 0;
  const unsigned long long t0 = gtimer();
  for (int it = 0; it < iters; ++it) {
    const int p = it & 1;
    const unsigned phase = (unsigned)(it >> 1) & 1u;
    if (t == 0) mbar_expect_tx(&bar[p], expect);                      // arms this phase (the only pending arrival)
    if (t < NPR) {
      const unsigned long long v = ((unsigned long long)(unsigned)it << 32) | (unsigned)(rank * 1024 + t);
      if (has_lo) st_async_b64(mapa(smem_u32(&halo[p][1][t]), rank - 1), v, mapa(smem_u32(&bar[p]), rank - 1));   // I am "above" my lower neighbour
      if (has_hi) st_async_b64(mapa(smem_u32(&halo[p][0][t]), rank + 1), v, mapa(smem_u32(&bar[p]), rank + 1));   // and "below" the upper one
    }
    if (!mbar_wait(&bar[p], phase)) { if (t == 0) atomicAdd(errors, 1 << 20); break; }   // gave up: report, leave
    if (t < NPR) {                                                    // payload of both neighbours for THIS iteration
      if (has_lo && halo[p][0][t] != (((unsigned long long)(unsigned)it << 32) | (unsigned)((rank - 1) * 1024 + t))) ++bad;
      if (has_hi && halo[p][1][t] != (((unsigned long long)(unsigned)it << 32) | (unsigned)((rank + 1) * 1024 + t))) ++bad;
    }
    __syncthreads();                                                  // the smoother's local hand-over between sweeps
  }
  const unsigned long long t1 = gtimer();
  if (bad) atomicAdd(errors, bad);
  if (rank == 0 && t == 0) out[0] = t1 - t0;
  cl.sync();                                                          // no CTA exits while a peer may still write into it
}

template <class K, class... A>
static void launch(K kern, A... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CLUSTER);
  cfg.blockDim = dim3(THREADS);
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CLUSTER; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  CK(cudaLaunchKernelEx(&cfg, kern, args...));
  CK(cudaDeviceSynchronize());
}

int main(int argc, char **argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 2000;
  unsigned long long *out;
  int *errors;
  CK(cudaMallocManaged(&out, sizeof(*out)));
  CK(cudaMallocManaged(&errors, sizeof(*errors)));
  *errors = 0;
  for (int rep = 0; rep < 2; ++rep) {          // second pass = warm
    launch(k_cg_sync, iters, out);
    const double a = (double)*out / iters;
    launch(k_relaxed, iters, out);
    const double b = (double)*out / iters;
    launch(k_neighbour, iters, out, errors);
    const double c = (double)*out / iters;
    if (rep == 1) {
      printf("cluster of %d CTAs x %d threads, %d iterations\n", CLUSTER, THREADS, iters);
      printf("(A) cg cluster.sync()                                  : %8.1f ns per hand-over\n", a);
      printf("(B) barrier.cluster arrive.relaxed + wait              : %8.1f ns\n", b);
      printf("(C) st.async halo rows + mbarrier wait + __syncthreads : %8.1f ns   (payload errors: %d)\n", c, *errors);
    }
  }
  return 0;
}
