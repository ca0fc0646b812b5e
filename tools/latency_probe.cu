// latency_probe.cu -- development probe: host-visible round trip of a tiny kernel on B200 under different completion protocols.
#include <cuda_runtime.h>
#include <cstdio>
#include <chrono>
#include <atomic>
using clk = std::chrono::steady_clock;

__global__ void k_empty() {}
__global__ void k_write_dev(double* out, int n) { if (threadIdx.x < n) out[blockIdx.x * n + threadIdx.x] = threadIdx.x; }
__global__ void k_flag(double* out_host, int n, unsigned* done, volatile unsigned* flag, unsigned seq, int mode)
{
  if (threadIdx.x < n) out_host[blockIdx.x * n + threadIdx.x] = threadIdx.x + seq;
  __syncthreads();
  if (threadIdx.x == 0) {
    if (mode == 0) __threadfence_system(); else __threadfence();
    unsigned prev = atomicAdd(done, 1u);
    if (prev == gridDim.x - 1) {
      *done = 0;
      if (mode == 0) __threadfence_system(); else __threadfence();
      *flag = seq;
    }
  }
}

template <class F> double timeit(F f, int reps) {
  for (int i = 0; i < 50; ++i) f();
  auto t0 = clk::now();
  for (int i = 0; i < reps; ++i) f();
  return std::chrono::duration<double, std::micro>(clk::now() - t0).count() / reps;
}

int main()
{
  cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  double *dev, *host, *hostmap; unsigned *done, *flag;
  cudaMalloc(&dev, 8 * 64 * 8); cudaMalloc(&done, 4); cudaMemset(done, 0, 4);
  cudaHostAlloc(&host, 8 * 64 * 8, cudaHostAllocDefault);
  cudaHostAlloc(&hostmap, 8 * 64 * 8, cudaHostAllocMapped);
  cudaHostAlloc(&flag, 64, cudaHostAllocMapped); *flag = 0;
  const int reps = 3000;
  for (int grid : {1, 5, 15}) {
    printf("grid %d\n", grid);
    printf("  empty kernel + cudaStreamSynchronize            %.1f us\n", timeit([&] { k_empty<<<grid, 256, 0, st>>>(); cudaStreamSynchronize(st); }, reps));
    printf("  write dev + cudaMemcpyAsync D2H + sync          %.1f us\n", timeit([&] { k_write_dev<<<grid, 256, 0, st>>>(dev, 40); cudaMemcpyAsync(host, dev, grid * 40 * 8, cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st); }, reps));
    printf("  write mapped host + cudaStreamSynchronize       %.1f us\n", timeit([&] { k_write_dev<<<grid, 256, 0, st>>>(hostmap, 40); cudaStreamSynchronize(st); }, reps));
    unsigned seq = 0;
    printf("  mapped host + flag, fence.sys, host spin        %.1f us\n", timeit([&] { ++seq; k_flag<<<grid, 256, 0, st>>>(hostmap, 40, done, flag, seq, 0); while (*(volatile unsigned*)flag != seq) {} }, reps));
    cudaStreamSynchronize(st);
    printf("  mapped host + flag, fence.gpu, host spin        %.1f us\n", timeit([&] { ++seq; k_flag<<<grid, 256, 0, st>>>(hostmap, 40, done, flag, seq, 1); while (*(volatile unsigned*)flag != seq) {} }, reps));
    cudaStreamSynchronize(st);
    printf("  launch only (async, amortised over 3000)        %.1f us\n", timeit([&] { k_empty<<<grid, 256, 0, st>>>(); }, reps));
    cudaStreamSynchronize(st);
  }
  return 0;
}
