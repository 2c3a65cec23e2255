// host_util.cuh -- growing device / page-locked host buffers shared by the translation units of libmvgcuda, and the narrow
// internal view of a context that the geometric-filter unit (geometric_api.cu, compiled with -fmad=false) works through.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>

struct mvgcuda_ctx;

namespace mvgcuda {

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;  // elements
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, n * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

template <typename T>
struct PinnedBuf {
  T* p = nullptr;
  size_t cap = 0;
  // grow, preserving the first `keep` elements
  cudaError_t reserve(size_t n, size_t keep = 0) {
    if (n <= cap) return cudaSuccess;
    size_t ncap = std::max(n, cap + cap / 2);
    T* np = nullptr;
    cudaError_t e = cudaMallocHost(&np, ncap * sizeof(T));
    if (e != cudaSuccess) return e;
    if (keep) memcpy(np, p, keep * sizeof(T));
    if (p) cudaFreeHost(p);
    p = np;
    cap = ncap;
    return cudaSuccess;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

// What geometric_api.cu may see of a context (defined in mvgcuda_api.cu).
struct CtxView {
  int device;
  int sm_count;
  cudaStream_t stream;
  const float2* feats;   // (x, y) of every arena row; nullptr until features were set
  const int* row0;       // first arena row of every image (host)
  const int* rows;       // rows of every image (host)
  int n_images;
  void** geo;            // slot for the geometric unit's own state ...
  void (**geo_free)(void*);  // ... and its destructor, called by mvgcuda_destroy
};
CtxView ctx_view(mvgcuda_ctx* ctx);
void ctx_set_error(mvgcuda_ctx* ctx, const char* msg);
int ctx_wait_uploads(mvgcuda_ctx* ctx);  // order the context's stream after every streamed image

}  // namespace mvgcuda
