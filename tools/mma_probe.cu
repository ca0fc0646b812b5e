// mma_probe.cu -- development probe: checks the register-fragment layout assumed for
// mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 (A 16x32 s8 row-major, B 32x8 s8 col-major, C 16x8 s32).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

__global__ void k(const int8_t* A, const int8_t* B, int* C)
{
  const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
  auto packA = [&](int row, int col0) {
    uint32_t v = 0;
    for (int j = 0; j < 4; ++j) v |= (uint32_t)(uint8_t)A[row * 32 + col0 + j] << (8 * j);
    return v;
  };
  auto packB = [&](int k0, int n) {
    uint32_t v = 0;
    for (int j = 0; j < 4; ++j) v |= (uint32_t)(uint8_t)B[(k0 + j) * 8 + n] << (8 * j);
    return v;
  };
  uint32_t a0 = packA(g, 4 * t), a1 = packA(g + 8, 4 * t), a2 = packA(g, 16 + 4 * t), a3 = packA(g + 8, 16 + 4 * t);
  uint32_t b0 = packB(4 * t, g), b1 = packB(16 + 4 * t, g);
  int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(c0), "+r"(c1), "+r"(c2), "+r"(c3)
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  C[g * 8 + 2 * t] = c0; C[g * 8 + 2 * t + 1] = c1; C[(g + 8) * 8 + 2 * t] = c2; C[(g + 8) * 8 + 2 * t + 1] = c3;
}

int main()
{
  int8_t hA[16 * 32], hB[32 * 8]; int hC[128], ref[128];
  srand(1);
  for (auto& v : hA) v = rand() % 3;
  for (auto& v : hB) v = (rand() % 256) - 128;
  for (int i = 0; i < 16; ++i) for (int n = 0; n < 8; ++n) { int s = 0; for (int k = 0; k < 32; ++k) s += hA[i * 32 + k] * hB[k * 8 + n]; ref[i * 8 + n] = s; }
  int8_t *dA, *dB; int* dC;
  cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dC, sizeof hC);
  cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
  k<<<1, 32>>>(dA, dB, dC);
  cudaMemcpy(hC, dC, sizeof hC, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int i = 0; i < 128; ++i) bad += hC[i] != ref[i];
  printf("mma m16n8k32 s8 layout check: %s (%d mismatches) err=%s\n", bad ? "WRONG" : "OK", bad, cudaGetErrorString(cudaGetLastError()));
  return bad != 0;
}
