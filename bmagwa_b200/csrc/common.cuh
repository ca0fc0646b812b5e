// common.cuh -- shared declarations of the device side of libbmagwa_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdexcept>
#include <string>
#include <vector>
#include <atomic>

namespace bmg {

// ---------------------------------------------------------------------------------------
// error handling: C++ exceptions inside, int status + thread-local text at the C ABI
// (the reference throws std::runtime_error and lets it terminate)
// ---------------------------------------------------------------------------------------
struct Error : std::runtime_error {
  explicit Error(const std::string& m) : std::runtime_error(m) {}
};

void set_last_error(const std::string& m);

#define BMG_CUDA(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      throw ::bmg::Error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + \
                         ":" + std::to_string(__LINE__) + ")");                                \
  } while (0)

#define BMG_REQUIRE(cond, msg)                 \
  do {                                         \
    if (!(cond)) throw ::bmg::Error(msg);      \
  } while (0)

extern std::atomic<uint64_t> g_launches, g_h2d_bytes, g_d2h_bytes;
inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

// every host<->device copy of the library goes through these, so traffic can be reported (bench.py e2e)
inline void copy_h2d(void* dst, const void* src, size_t bytes, cudaStream_t st)
{
  g_h2d_bytes.fetch_add(bytes, std::memory_order_relaxed);
  BMG_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
}
inline void copy_d2h(void* dst, const void* src, size_t bytes, cudaStream_t st)
{
  g_d2h_bytes.fetch_add(bytes, std::memory_order_relaxed);
  BMG_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
}
inline void copy_h2d_sync(void* dst, const void* src, size_t bytes)
{
  g_h2d_bytes.fetch_add(bytes, std::memory_order_relaxed);
  BMG_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
}
inline void copy_d2h_sync(void* dst, const void* src, size_t bytes)
{
  g_d2h_bytes.fetch_add(bytes, std::memory_order_relaxed);
  BMG_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
}

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  void alloc(size_t count)
  {
    release();
    n = count;
    if (count) BMG_CUDA(cudaMalloc((void**)&p, count * sizeof(T)));
  }
  void zero(cudaStream_t st = 0)
  {
    if (n) BMG_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), st));
  }
  void release()
  {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  ~DevBuf() { release(); }
};

template <class T>
struct PinnedBuf {
  T* p = nullptr;
  size_t n = 0;
  PinnedBuf() {}
  PinnedBuf(const PinnedBuf&) = delete;
  PinnedBuf& operator=(const PinnedBuf&) = delete;
  void alloc(size_t count)
  {
    release();
    n = count;
    if (count) BMG_CUDA(cudaHostAlloc((void**)&p, count * sizeof(T), cudaHostAllocMapped));
  }
  void release()
  {
    if (p) cudaFreeHost(p);
    p = nullptr;
    n = 0;
  }
  ~PinnedBuf() { release(); }
};

// ---------------------------------------------------------------------------------------
// Packed genotype layout of the device store (DESIGN.md "Data layout in HBM")
//
//   SNP-major; one column = words_per_snp 32-bit words (16 individuals per word, individual
//   16w+p in bits [2p+1:2p], i.e. the same bit order as a little-endian read of the .bed bytes);
//   words_per_snp = ceil(n/16) rounded up to a multiple of 4 (16-byte aligned columns, so a
//   column or a tile of columns is a legal cp.async.bulk source).
//   The 2-bit field holds the genotype VALUE: 00 = 0, 01 = 1, 10 = 2.  A missing cell is stored
//   as 00 (the reference's initial imputation, data_model.hpp:80-84) and recorded in a sparse
//   CSR index (miss_off / miss_idx); 11 never occurs.  Padding bits beyond individual n-1 are 0.
// ---------------------------------------------------------------------------------------
inline int64_t words_for(int64_t n) { return (n + 15) / 16; }
inline int64_t stride_words_for(int64_t n) { return (words_for(n) + 3) / 4 * 4; }

// in-order position <-> heap index of the reference's implicit proposal tree
void build_inorder_permutation(int64_t m, std::vector<int32_t>& order);

struct PeerShard {
  int64_t lo, hi;
  const uint32_t* codes;  // device pointer valid on this device (P2P / IPC mapped)
  bool ipc_opened;
};

struct Store {
  int device = 0;
  int64_t n = 0, m_g = 0, lo = 0, hi = 0;   // shard [lo, hi) of m_g
  int64_t m = 0;                            // hi - lo
  int64_t W = 0, Wp = 0;                    // words per column, padded stride
  int m_e = 0;
  int sm_count = 148;
  bool recode = false;
  DevBuf<uint32_t> codes;                   // m * Wp (+ slack)
  DevBuf<int32_t> n1, n2, nmiss;            // per local SNP
  DevBuf<uint8_t> swapped;
  DevBuf<double> mom;                       // 2 per local SNP: (s, v)
  DevBuf<double> snp_mean, snp_var;         // per local SNP mean and unbiased variance (Data::compute_g_var_and_mean terms)
  DevBuf<int64_t> miss_off;                 // m + 1
  DevBuf<int32_t> miss_idx;                 // total missing cells
  std::vector<int64_t> h_miss_off;          // host mirror
  int64_t n_missing = 0;
  DevBuf<double> y, e;                      // n, n * m_e (col-major, includes the ones column)
  std::vector<double> h_y;
  DevBuf<int32_t> inorder;                  // in-order position -> local SNP (heap index)
  std::vector<int32_t> h_inorder;
  std::vector<PeerShard> peers;
  double summaries[6] = {0, 0, 0, 0, 0, 0};

  const uint32_t* column_ptr(int64_t snp_global) const;  // local or peer column (device pointer)
  bool is_local(int64_t snp) const { return snp >= lo && snp < hi; }
};

}  // namespace bmg
