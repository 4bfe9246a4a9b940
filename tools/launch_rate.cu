// Development probe: how many kernel launches per second one process gets onto one GPU, by the number of host threads
// (one stream each).  The block pipeline issues about 105 launches and 30 memsets / copies per block, mostly kernels of
// 5-60 us; this measures the ceiling that puts on blocks/s whatever the kernels do.
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o tools/launch_rate tools/launch_rate.cu -lpthread && tools/launch_rate
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <thread>
#include <vector>

__global__ void tiny_kernel(unsigned* p) {
  if (p && threadIdx.x == 9999) *p = 1;
}
// about `us` microseconds of dependent work on one warp
__global__ void busy_kernel(unsigned* p, int iters) {
  unsigned x = threadIdx.x;
  for (int i = 0; i < iters; i++) x = x * 1664525u + 1013904223u;
  if (p && x == 0xdeadbeef) *p = x;
}

static double run(int threads, int per_thread, int iters, int blocks) {
  std::vector<cudaStream_t> st(threads);
  for (auto& s : st) cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  std::atomic<int> ready{0};
  std::atomic<bool> go{false};
  std::vector<std::thread> th;
  for (int t = 0; t < threads; t++)
    th.emplace_back([&, t] {
      cudaSetDevice(0);
      ready++;
      while (!go.load()) std::this_thread::yield();
      for (int k = 0; k < per_thread; k++) {
        if (iters)
          busy_kernel<<<blocks, 128, 0, st[t]>>>(nullptr, iters);
        else
          tiny_kernel<<<1, 32, 0, st[t]>>>(nullptr);
      }
      cudaStreamSynchronize(st[t]);
    });
  while (ready.load() < threads) std::this_thread::yield();
  auto t0 = std::chrono::steady_clock::now();
  go = true;
  for (auto& x : th) x.join();
  double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  for (auto& x : st) cudaStreamDestroy(x);
  return s;
}

int main() {
  cudaSetDevice(0);
  cudaFree(0);
  run(4, 200, 0, 1);
  const int per = 4000;
  printf("%8s %10s %14s %14s\n", "threads", "kernel", "launches/s", "us per launch");
  for (int iters : {0, 2000}) {       // empty kernel; ~20 us kernel of 64 thread blocks
    for (int threads : {1, 2, 4, 8, 16, 30, 64}) {
      double s = run(threads, per, iters, 64);
      printf("%8d %10s %14.0f %14.2f\n", threads, iters ? "~20us x64" : "empty", threads * per / s, 1e6 * s / (threads * per));
      fflush(stdout);
    }
  }
  return 0;
}
