// Where does the start-up of a CUDA process on this box go?  Times cuInit (driver), the first runtime call, the primary
// context of device 0, a module load (first kernel-attribute query) and a 1 GB allocation.  Developer tool.
#include <cuda.h>
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
__global__ void k() {}
int main() {
  typedef std::chrono::steady_clock C;
  auto ms = [](C::time_point a, C::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
  C::time_point t0 = C::now();
  cuInit(0);
  C::time_point t1 = C::now();
  int n = 0;
  cudaGetDeviceCount(&n);
  C::time_point t2 = C::now();
  cudaSetDevice(0);
  cudaFree(0);
  C::time_point t3 = C::now();
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, k);
  C::time_point t4 = C::now();
  void* p = nullptr;
  cudaMalloc(&p, 1ull << 30);
  C::time_point t5 = C::now();
  void* h = nullptr;
  cudaHostAlloc(&h, 64ull << 20, cudaHostAllocPortable);
  C::time_point t6 = C::now();
  std::printf("devices %d: cuInit %.0f ms, cudaGetDeviceCount %.0f ms, primary context (cudaSetDevice + cudaFree(0)) %.0f ms, module load %.0f ms, "
              "cudaMalloc 1 GB %.0f ms, cudaHostAlloc 64 MB %.0f ms\n", n, ms(t0, t1), ms(t1, t2), ms(t2, t3), ms(t3, t4), ms(t4, t5), ms(t5, t6));
  return 0;
}
