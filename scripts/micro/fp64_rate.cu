// Microbenchmark: sustained fp64 / fp32 FMA throughput and L2/HBM copy rate on the current GPU.
#include <cstdio>
#include <cuda_runtime.h>
template <typename T, int CH>
__global__ void k_fma(T *out, T a, T b, int iters) {
  T acc[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) acc[i] = (T)(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) acc[i] = acc[i] * a + b;
  }
  T s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_copy(const double2 *__restrict__ a, double2 *__restrict__ b, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}
template <typename F> float timeit(F f, int rep) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); for (int i = 0; i < rep; ++i) f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / rep;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("%s SMs=%d clock=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  void *buf; cudaMalloc(&buf, 1 << 26);
  const int blocks = p.multiProcessorCount * 8, thr = 256, iters = 4096;
  for (int rep = 0; rep < 2; ++rep) {
    float ms = timeit([&] { k_fma<double, 8><<<blocks, thr>>>((double *)buf, 1.0000001, 1e-9, iters); }, 5);
    double fl = 2.0 * blocks * thr * 8.0 * iters;
    printf("fp64 FMA: %.3f ms -> %.2f TFLOP/s (%.1f FMA/clk/SM at %.0f MHz)\n", ms, fl / ms / 1e9,
           fl / 2 / (ms * 1e-3) / p.multiProcessorCount / (p.clockRate * 1e3), p.clockRate / 1e3);
    ms = timeit([&] { k_fma<float, 8><<<blocks, thr>>>((float *)buf, 1.0000001f, 1e-9f, iters); }, 5);
    printf("fp32 FMA: %.3f ms -> %.2f TFLOP/s\n", ms, fl / ms / 1e9);
  }
  size_t n = (size_t)1 << 26;  // 1 GiB of double2 each way
  double2 *a, *b; cudaMalloc(&a, n * 16); cudaMalloc(&b, n * 16); cudaMemset(a, 0, n * 16);
  float ms = timeit([&] { k_copy<<<p.multiProcessorCount * 16, 512>>>(a, b, n); }, 5);
  printf("copy 1GiB+1GiB: %.3f ms -> %.0f GB/s\n", ms, 2.0 * n * 16 / ms / 1e6);
  size_t m = (size_t)1 << 21;  // 32 MiB each way: L2 resident
  ms = timeit([&] { k_copy<<<p.multiProcessorCount * 16, 512>>>(a, b, m); }, 20);
  printf("copy 32MiB+32MiB (L2): %.4f ms -> %.0f GB/s\n", ms, 2.0 * m * 16 / ms / 1e6);
  return 0;
}
