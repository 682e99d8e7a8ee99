#include <cstdio>
#include <cuda_runtime.h>
int main() {
  int a = 0, b = 0, l2 = 0; size_t lim = 0;
  cudaDeviceGetAttribute(&a, cudaDevAttrMaxPersistingL2CacheSize, 0);
  cudaDeviceGetAttribute(&b, cudaDevAttrMaxAccessPolicyWindowSize, 0);
  cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, 0);
  cudaDeviceGetLimit(&lim, cudaLimitPersistingL2CacheSize);
  printf("L2 %d B, maxPersisting %d B, maxWindow %d B, current limit %zu\n", l2, a, b, lim);
  cudaError_t e = cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)a);
  cudaDeviceGetLimit(&lim, cudaLimitPersistingL2CacheSize);
  printf("set -> %s, limit now %zu\n", cudaGetErrorString(e), lim);
  return 0;
}
