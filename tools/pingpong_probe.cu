// Round trip host -> persistent kernel -> host through mapped pinned memory (development probe):
//   host writes seq to a mailbox word, the kernel polls it (ld.relaxed.sys or ld.volatile), echoes it to a second
//   mapped word, the host polls that.  Prints the average round trip for each polling flavour and for a kernel
//   that reads a 1552-byte mailbox with 97 threads per trip (what k_colstats_server does).
#include <cuda_runtime.h>
#include <stdio.h>
#include <time.h>
#include <atomic>
static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

template <int MODE>
__global__ void k_echo(const unsigned int* mail, unsigned int* echo, int rounds)
{
  __shared__ unsigned int sh;
  unsigned int last = 0;
  for (int r = 0; r < rounds; ++r) {
    if (MODE == 2) {   // 97 threads read 16 bytes each per trip
      for (;;) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (threadIdx.x < 97) asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(reinterpret_cast<const uint4*>(mail) + threadIdx.x) : "memory");
        if (threadIdx.x == 0) sh = v.x;
        __syncthreads();
        const unsigned int s = sh;
        __syncthreads();
        if (s != last) { last = s; break; }
      }
    } else if (threadIdx.x == 0) {
      unsigned int s;
      do {
        if (MODE == 0) asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(s) : "l"(mail) : "memory");
        else s = *reinterpret_cast<const volatile unsigned int*>(mail);
      } while (s == last);
      last = s;
    }
    if (threadIdx.x == 0) { *reinterpret_cast<volatile unsigned int*>(echo) = last; }
    if (MODE != 2) __syncthreads();
  }
}

int main()
{
  unsigned int *mail, *echo;
  cudaHostAlloc(&mail, 4096, cudaHostAllocMapped);
  cudaHostAlloc(&echo, 4096, cudaHostAllocMapped);
  const int rounds = 20000;
  for (int mode = 0; mode < 3; ++mode) {
    mail[0] = 0; echo[0] = 0;
    std::atomic_thread_fence(std::memory_order_seq_cst);
    if (mode == 0) k_echo<0><<<1, 256>>>(mail, echo, rounds);
    if (mode == 1) k_echo<1><<<1, 256>>>(mail, echo, rounds);
    if (mode == 2) k_echo<2><<<1, 256>>>(mail, echo, rounds);
    const double t0 = now();
    for (int r = 1; r <= rounds; ++r) {
      *reinterpret_cast<volatile unsigned int*>(mail) = (unsigned int)r;
      std::atomic_thread_fence(std::memory_order_seq_cst);
      while (*reinterpret_cast<volatile unsigned int*>(echo) != (unsigned int)r) {
        if (now() - t0 > 20.0) { printf("mode %d: timed out at round %d\n", mode, r); return 1; }
      }
    }
    const double t1 = now();
    cudaDeviceSynchronize();
    printf("mode %d (%s): %.2f us per round trip\n", mode, mode == 0 ? "ld.relaxed.sys, 1 thread" : mode == 1 ? "volatile, 1 thread" : "97 x 16 B per trip", 1e6 * (t1 - t0) / rounds);
  }
  return 0;
}
