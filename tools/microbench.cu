// microbench.cu -- development probe: FP64 FMA issue rate / latency on B200 with normal vs denormal
// multiplicands, with the LOP3-per-DFMA mix the scan kernel uses, as a function of ILP and occupancy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench.bin tools/microbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

// ILP independent chains per thread, LOP3 + DFMA per step (denormal multiplicand)
template <int ILP>
__global__ void k_ilp(const uint32_t* __restrict__ in, double* out, int iters)
{
  double acc[ILP], r[ILP];
  for (int i = 0; i < ILP; ++i) { acc[i] = 0.0; r[i] = 1.0 + i + threadIdx.x; }
  uint32_t w = in[threadIdx.x & 31];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int p = 0; p < 16; ++p) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        const double x = __hiloint2double(0, (int)(w & (3u << (2 * p))));
        asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(acc[i]) : "d"(x), "d"(r[i]));
      }
    }
    w = w * 1664525u + 1013904223u;
  }
  double s = 0;
  for (int i = 0; i < ILP; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
void run(int warps_per_sm)
{
  uint32_t* in; double* out;
  cudaMalloc(&in, 128); cudaMalloc(&out, 148 * 64 * 32 * sizeof(double));
  cudaMemset(in, 0x5a, 128);
  const int iters = 4000 / ILP * 2, blocks = 148, threads = 32 * warps_per_sm;
  k_ilp<ILP><<<blocks, threads>>>(in, out, 10);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k_ilp<ILP><<<blocks, threads>>>(in, out, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double fma_per_warp = (double)iters * 16 * ILP;
  const double cyc = ms * 1e-3 * 1.9e9;
  printf("warps/SM %2d ILP %d: %.3f ms  %.2f T DFMA/s  cycles per DFMA per warp %.2f  per SMSP issue interval %.2f\n",
         warps_per_sm, ILP, ms, fma_per_warp * warps_per_sm * 148 * 32 / ms / 1e9, cyc / fma_per_warp,
         cyc / (fma_per_warp * warps_per_sm / 4.0));
  cudaFree(in); cudaFree(out);
}

int main()
{
  run<1>(1); run<2>(1); run<4>(1); run<8>(1);
  run<1>(4); run<2>(4); run<4>(4); run<8>(4);
  run<1>(16); run<2>(16); run<4>(16); run<8>(16);
  run<2>(15); run<2>(20); run<8>(20); run<2>(32); run<8>(32);
  return 0;
}
