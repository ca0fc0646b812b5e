// microbench3.cu -- development probe: per-instruction throughput on B200 for the integer ops the IMMA scan uses.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

template <int OP>
__global__ void __launch_bounds__(512) k(uint32_t* out, int iters, uint32_t seed)
{
  uint32_t x[8];
  for (int i = 0; i < 8; ++i) x[i] = seed * (threadIdx.x + 1) + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (OP == 0) asm volatile("prmt.b32 %0, %0, 0, 0x1111;" : "+r"(x[i]));
        if (OP == 1) asm volatile("lop3.b32 %0, %0, 0xC0300C03, %1, 0x6a;" : "+r"(x[i]) : "r"(seed));
        if (OP == 2) asm volatile("mad.lo.u32 %0, %0, 129, %1;" : "+r"(x[i]) : "r"(seed));
        if (OP == 3) asm volatile("shf.r.wrap.b32 %0, %0, %0, 7;" : "+r"(x[i]));
        if (OP == 4) asm volatile("bfe.u32 %0, %0, 8, 8;" : "+r"(x[i]));
        if (OP == 5) asm volatile("prmt.b32 %0, %0, %1, 0x4441;" : "+r"(x[i]) : "r"(seed));
        if (OP == 6) asm volatile("and.b32 %0, %0, 0xC0300C03;" : "+r"(x[i]));
        if (OP == 7) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(seed));
      }
    }
  }
  uint32_t s = 0;
  for (int i = 0; i < 8; ++i) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// IMMA throughput: CH independent accumulator chains per warp
template <int CH>
__global__ void __launch_bounds__(512) kmma(int* out, int iters)
{
  int c[CH][4];
  for (int j = 0; j < CH; ++j) for (int q = 0; q < 4; ++q) c[j][q] = 0;
  uint32_t a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int j = 0; j < CH; ++j)
        asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+r"(c[j][0]), "+r"(c[j][1]), "+r"(c[j][2]), "+r"(c[j][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  int s = 0;
  for (int j = 0; j < CH; ++j) for (int q = 0; q < 4; ++q) s += c[j][q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP> void run(const char* name, int warps)
{
  uint32_t* out; cudaMalloc(&out, 148 * 1024 * 4);
  const int iters = 2000;
  k<OP><<<148, 32 * warps>>>(out, 10, 12345); cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<OP><<<148, 32 * warps>>>(out, iters, 12345); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double n = (double)iters * 64 * warps;   // warp-instructions per SM
  printf("%-28s warps/SM %2d: %.2f cycles per warp-instr per SMSP (at 1.9 GHz)\n", name, warps, ms * 1e-3 * 1.9e9 / (n / 4));
  cudaFree(out);
}
template <int CH> void runmma(int warps)
{
  int* out; cudaMalloc(&out, 148 * 1024 * 4);
  const int iters = 1000;
  kmma<CH><<<148, 32 * warps>>>(out, 10); cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); kmma<CH><<<148, 32 * warps>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double n = (double)iters * 8 * CH * warps;
  printf("IMMA.16832 chains/warp %d warps/SM %2d: %.2f cycles per IMMA per SMSP; %.1f TOPS\n", CH, warps, ms * 1e-3 * 1.9e9 / (n / 4),
         n * 148 * 16 * 8 * 32 * 2 / (ms * 1e-3) / 1e12);
  cudaFree(out);
}
int main()
{
  for (int w : {4, 16}) {
    run<0>("PRMT imm (dependent x8 ILP)", w); run<5>("PRMT reg,imm", w); run<1>("LOP3 imm", w); run<6>("AND imm", w);
    run<2>("IMAD", w); run<3>("SHF", w); run<4>("BFE", w); run<7>("IADD", w);
  }
  runmma<1>(4); runmma<2>(4); runmma<4>(4); runmma<1>(16); runmma<4>(16); runmma<1>(20); runmma<2>(20);
  return 0;
}
