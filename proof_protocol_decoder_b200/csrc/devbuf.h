// devbuf.h — growable device buffers and the CUDA error check shared by the host-side sources.
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "host_arena.h"  // Fail

#define CUDA_OK(expr)                                                                          \
  do {                                                                                         \
    cudaError_t e_ = (expr);                                                                   \
    if (e_ != cudaSuccess) {                                                                   \
      throw ppd::Fail{PPD_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)};       \
    }                                                                                          \
  } while (0)

namespace ppd {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  // contents are NOT preserved when the buffer grows
  void reserve(size_t n) {
    if (n <= cap) return;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = n + n / 4 + 256;
    CUDA_OK(cudaMalloc(&p, want));
    cap = want;
  }
  // grows keeping the first `keep` bytes (device-to-device copy on `st`, then the old block is freed)
  void reserve_keep(size_t n, size_t keep, cudaStream_t st) {
    if (n <= cap) return;
    size_t want = n + n / 2 + 256;
    void* q = nullptr;
    CUDA_OK(cudaMalloc(&q, want));
    if (p && keep) {
      cudaError_t e = cudaMemcpyAsync(q, p, keep, cudaMemcpyDeviceToDevice, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) {
        cudaFree(q);
        throw Fail{PPD_ERR_CUDA, std::string("reserve_keep: ") + cudaGetErrorString(e)};
      }
    }
    if (p) cudaFree(p);
    p = q;
    cap = want;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() const {
    return reinterpret_cast<T*>(p);
  }
};

}  // namespace ppd
