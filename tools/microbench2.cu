// microbench2.cu -- development probe for the scan inner loop structure on B200.
//  A: register-resident residuals, fully unrolled (what ptxas turns into 2 interleaved chains, no operand reuse)
//  B: residuals in shared memory, p-loop kept rolled: 8 independent chains per iteration sharing one residual
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__global__ void __launch_bounds__(160, 3) kA(const uint32_t* __restrict__ in, const double* __restrict__ rin, double* out, int iters)
{
  double rp[32];
  for (int p = 0; p < 32; ++p) rp[p] = rin[threadIdx.x * 32 + p];
  double tot = 0;
  uint32_t seed = in[threadIdx.x & 31];
  for (int it = 0; it < iters; ++it) {
    uint2 w[8]; double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { seed = seed * 1664525u + 1013904223u; w[i].x = seed; w[i].y = seed ^ 0x9e3779b9u; acc[i] = 0; }
#pragma unroll
    for (int p = 0; p < 16; ++p)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fma(__hiloint2double(0, (int)(w[i].x & (3u << (2 * p)))), rp[p], acc[i]);
#pragma unroll
    for (int p = 0; p < 16; ++p)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fma(__hiloint2double(0, (int)(w[i].y & (3u << (2 * p)))), rp[16 + p], acc[i]);
    for (int i = 0; i < 8; ++i) tot += acc[i];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = tot;
}

template <int UNROLL>
__global__ void __launch_bounds__(160) kB(const uint32_t* __restrict__ in, const double* __restrict__ rin, double* out, int iters)
{
  extern __shared__ double sr[];  // [32][T]
  const int T = blockDim.x, t = threadIdx.x;
  for (int q = 0; q < 32; ++q) sr[q * T + t] = rin[t * 32 + q];
  __syncthreads();
  double tot = 0;
  uint32_t seed = in[t & 31];
  for (int it = 0; it < iters; ++it) {
    uint2 w[8]; double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { seed = seed * 1664525u + 1013904223u; w[i].x = seed; w[i].y = seed ^ 0x9e3779b9u; acc[i] = 0; }
    const double* rq = sr + t;
    uint32_t mask = 3u;
#pragma unroll UNROLL
    for (int q = 0; q < 16; ++q) {
      const double r = *rq; rq += T;
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fma(__hiloint2double(0, (int)(w[i].x & mask)), r, acc[i]);
      mask <<= 2;
    }
    mask = 3u;
#pragma unroll UNROLL
    for (int q = 0; q < 16; ++q) {
      const double r = *rq; rq += T;
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fma(__hiloint2double(0, (int)(w[i].y & mask)), r, acc[i]);
      mask <<= 2;
    }
    for (int i = 0; i < 8; ++i) tot += acc[i];
  }
  out[blockIdx.x * blockDim.x + t] = tot;
}

int main()
{
  uint32_t* in; double* out; double* rin;
  cudaMalloc(&in, 128); cudaMalloc(&out, 148 * 16 * 160 * sizeof(double)); cudaMalloc(&rin, 160 * 32 * 8);
  cudaMemset(in, 0x5a, 128); cudaMemset(rin, 0x3f, 160 * 32 * 8);
  const int iters = 400;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  for (int ctas = 1; ctas <= 6; ++ctas) {
    const int blocks = 148 * ctas;
    const double fma = (double)blocks * 160 * iters * 256;
    if (ctas <= 3) {
      kA<<<blocks, 160>>>(in, rin, out, 2); cudaDeviceSynchronize();
      cudaEventRecord(e0); kA<<<blocks, 160>>>(in, rin, out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
      printf("A regs-unrolled   CTAs/SM %d (warps/SM %2d): %.3f ms %.2f T DFMA/s\n", ctas, ctas * 5, ms, fma / ms / 1e9);
    }
    cudaFuncSetAttribute(kB<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 256);
    cudaFuncSetAttribute(kB<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 256);
    cudaFuncSetAttribute(kB<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 256);
    if (ctas > 5) continue;  // 5 x 40 KB
    kB<1><<<blocks, 160, 160 * 256>>>(in, rin, out, 2); cudaDeviceSynchronize();
    cudaEventRecord(e0); kB<1><<<blocks, 160, 160 * 256>>>(in, rin, out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("B smem-rolled u1  CTAs/SM %d (warps/SM %2d): %.3f ms %.2f T DFMA/s\n", ctas, ctas * 5, ms, fma / ms / 1e9);
    cudaEventRecord(e0); kB<2><<<blocks, 160, 160 * 256>>>(in, rin, out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("B smem-rolled u2  CTAs/SM %d (warps/SM %2d): %.3f ms %.2f T DFMA/s\n", ctas, ctas * 5, ms, fma / ms / 1e9);
    cudaEventRecord(e0); kB<4><<<blocks, 160, 160 * 256>>>(in, rin, out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("B smem-rolled u4  CTAs/SM %d (warps/SM %2d): %.3f ms %.2f T DFMA/s\n", ctas, ctas * 5, ms, fma / ms / 1e9);
  }
  return 0;
}
