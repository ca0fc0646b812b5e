// colstats.cu -- per-proposal column statistics (src/model.hpp:453-470) and the probit latent
// update.
//
// For each candidate SNP c of a move, in ONE launch:  x_c'y, x_c'E_j, x_c'x_l for the model's
// SNPs l, and x_c'x_d among the candidates.  The reference unpacks the column to n doubles and
// runs ddot + dgemv over the n x k double matrix it keeps per model (and copies on every
// accept/reject, model.hpp:115-162); here the model's columns stay packed (the store IS the
// gamma-column cache) and genotype-by-genotype products are exact integer popcount arithmetic.
#include "common.cuh"
#include "store.cuh"
#include "philox.cuh"
#include <stdlib.h>
#include <atomic>

namespace bmg {

constexpr int kSegWords = 2048;  // words of a column handled by one CTA (32768 individuals)

struct ColStatArgs {
  const uint32_t* const* cand_cols;   // m_c packed columns
  const uint32_t* const* model_cols;  // k packed columns
  int m_c, k, m_e;
  int64_t n, W;
  int n_seg;
  const double* y;   // n
  const double* e;   // n x m_e col-major
  double* out;       // [m_c][n_seg][n_tasks], n_tasks = m_e + 1 + k + m_c  (task 0 = y, 1.. = E_j, then loci, then cands)
};

__device__ __forceinline__ int packed_dot(uint32_t a, uint32_t b)
{
  // fields hold values 0,1,2 as 00,01,10:  a*b = 4 a1 b1 + 2 a1 b0 + 2 a0 b1 + a0 b0
  const uint32_t M = 0x55555555u;
  const uint32_t a0 = a & M, a1 = (a >> 1) & M, b0 = b & M, b1 = (b >> 1) & M;
  return 4 * __popc(a1 & b1) + 2 * (__popc(a1 & b0) + __popc(a0 & b1)) + __popc(a0 & b0);
}

__global__ void __launch_bounds__(256) k_column_stats(const ColStatArgs a)
{
  __shared__ uint32_t cw[kSegWords];
  const int c = blockIdx.x, seg = blockIdx.y;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nw = blockDim.x >> 5;
  const int64_t w0 = (int64_t)seg * kSegWords;
  const int nwords = (int)min((int64_t)kSegWords, a.W - w0);
  const uint32_t* col = a.cand_cols[c];
  for (int w = t; w < nwords; w += blockDim.x) cw[w] = col[w0 + w];
  __syncthreads();
  const int n_tasks = a.m_e + 1 + a.k + a.m_c;
  double* out = a.out + ((int64_t)c * a.n_seg + seg) * n_tasks;
  for (int task = warp; task < n_tasks; task += nw) {
    double res;
    if (task <= a.m_e) {
      // x_c . y (task 0) or x_c . E_j (task 1 + j)
      const double* vec = task == 0 ? a.y : a.e + (int64_t)(task - 1) * a.n;
      double acc = 0.0;
      for (int w = lane; w < nwords; w += 32) {
        const uint32_t word = cw[w];
        if (word == 0) continue;
        const int64_t i0 = 16 * (w0 + w);
#pragma unroll
        for (int p = 0; p < 16; ++p) {
          const int64_t i = i0 + p;
          const double g = (double)((word >> (2 * p)) & 3u);
          if (i < a.n) acc = fma(g, vec[i], acc);
        }
      }
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      res = acc;
    } else {
      const int q = task - a.m_e - 1;
      const uint32_t* other = q < a.k ? a.model_cols[q] : a.cand_cols[q - a.k];
      int acc = 0;
      for (int w = lane; w < nwords; w += 32) acc += packed_dot(cw[w], other[w0 + w]);
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      res = (double)acc;
    }
    if (lane == 0) out[task] = res;
  }
}

// ---- latency path: pointers passed BY VALUE (kernel parameters, or the persistent server's mailbox), results written
// straight into mapped pinned host memory as self-validating tagged words the host spins on.  No cudaMemcpy and no
// stream synchronisation on the per-iteration path.
constexpr int kInlinePtrs = 250;   // candidates + model columns of one request (2 kB of kernel parameters / mailbox)
struct ColStatInline {
  const uint32_t* cols[kInlinePtrs];  // m_c candidate columns, then k model columns
  int m_c, k, m_e;
  int64_t n, W;
  int n_seg, seg_words;
  const double* y;
  const double* e;
  ulonglong2* out_host;      // tagged results in mapped host memory: [m_c][n_seg][n_tasks], or [m_c][n_tasks] with `part`
  unsigned int seq;          // launch sequence number = the tag
  // Many slices per column (n_seg > kHostReduceSegs, i.e. n > 8,192): the per-slice results stay in device memory
  // (part, [m_c][n_seg][n_tasks]) and the LAST work item to finish (ticket) adds them in slice order -- the order the host
  // uses otherwise, so the numbers are the same -- and publishes only the sums: the host then reads 16 (1 + m_e + k + m_c)
  // bytes per candidate instead of n_seg times as much (at n = 50,000: 0.4 kB instead of 22 kB of freshly written lines).
  double* part;              // nullptr: the host adds the slices
  unsigned int* ticket;      // [4], slot seq & 3; left at 0 by the last item
};
constexpr int kHostReduceSegs = 8;
constexpr int kReduceTile = 4096;   // doubles of shared memory for the last item's reduction (32 kB)

// A result travels to the host as two 8-byte words, each carrying half of the double and the launch's sequence
// number in its upper half, written by ONE 16-byte store.  Every word validates itself, so the kernel needs no
// completion flag, no CTA counter and no system-scope fence (each of which costs a PCIe round trip): the host polls
// the words until they carry the expected tag.
__device__ __forceinline__ void publish_tagged(ulonglong2* slot, double v, unsigned int seq)
{
  const unsigned long long b = (unsigned long long)__double_as_longlong(v), tag = (unsigned long long)seq << 32;
  const unsigned long long w0 = (b & 0xFFFFFFFFull) | tag, w1 = (b >> 32) | tag;
  // .volatile: written through to host memory at once (a plain store may sit in L2 until the kernel ends, which a
  // persistent kernel never does)
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(slot), "l"(w0), "l"(w1) : "memory");
}

constexpr int kFastSegMax = 256;   // words per CTA on the latency path (64 or 256)
constexpr int kPreCov = 3;         // covariate columns (constant included) fetched before the candidate's words are known

// one (candidate, slice) work item; `a` may live in the kernel parameters or in shared memory
__device__ __forceinline__ void colstat_item(const ColStatInline& a, const int c, const int seg, uint32_t* cw, double (*fpart)[8])
{
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nw = blockDim.x >> 5;
  const int seg_words = a.seg_words;
  const int64_t w0 = (int64_t)seg * seg_words;
  const int nwords = (int)min((int64_t)seg_words, a.W - w0);
  const uint32_t* col = a.cols[c];
  const int n_fp = a.m_e + 1;
  const int n_tasks = n_fp + a.k + a.m_c;
  ulonglong2* out = a.out_host + ((int64_t)c * a.n_seg + seg) * n_tasks;
  double* part = a.part ? a.part + ((int64_t)c * a.n_seg + seg) * n_tasks : nullptr;
  const int64_t i_lo = 16 * w0, i_hi = min(a.n, i_lo + 16 * (int64_t)nwords);
  const bool small = seg_words <= 64;   // 1024 individuals per CTA: 4 per thread, everything this CTA reads issues at once
  // Every load of the CTA is issued BEFORE the candidate's words are waited for: y / covariates of this thread's
  // individuals and the first four tasks' words do not depend on them, so the kernel is one memory round trip deep.
  // (the first kPreCov covariate columns travel in registers; a data set with more pays one more memory round for the rest:
  // registers are what decides whether a scan CTA of another chain fits beside a server CTA, see server_start)
  double yv[4], ev[4][kPreCov];
  uint32_t ow0[4][2];
  if (small) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t i = i_lo + t + 256 * j;
      const bool ok = i < i_hi;
      yv[j] = ok ? __ldcg(a.y + i) : 0.0;   // y may be rewritten (probit) while the persistent server runs: bypass L1
#pragma unroll
      for (int q = 1; q <= kPreCov; ++q) ev[j][q - 1] = (ok && q < n_fp) ? a.e[(int64_t)(q - 1) * a.n + i] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int task = n_fp + warp + u * nw;
      const int q = task - n_fp;
      const uint32_t* other = task < n_tasks ? (q < a.k ? a.cols[a.m_c + q] : a.cols[q - a.k]) + w0 : nullptr;
#pragma unroll
      for (int j = 0; j < 2; ++j) ow0[u][j] = (other != nullptr && lane + 32 * j < nwords) ? __ldcg(other + lane + 32 * j) : 0u;
    }
  }
  // column words bypass L1 (__ldcg): a patched column of the overlay cache is rewritten between requests while the
  // persistent server keeps running, and L1 lines are only dropped at kernel boundaries
  if (t < nwords) cw[t] = __ldcg(col + w0 + t);
  __syncthreads();

  // (1) x_c'y and x_c'E_j: thread <-> individual (coalesced); a thread covers individuals i_lo + t + 256 j.
  {
    double acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = 0.0;
    if (small) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int li = t + 256 * j;
        const uint32_t word = cw[(li >> 4) & 63];
        const bool ok = i_lo + li < i_hi;
        const double g = ok ? (double)((word >> (2 * (li & 15))) & 3u) : 0.0;
        acc[0] = fma(g, yv[j], acc[0]);
#pragma unroll
        for (int q = 1; q <= kPreCov; ++q) acc[q] = fma(g, ev[j][q - 1], acc[q]);
#pragma unroll
        for (int q = kPreCov + 1; q < 8; ++q)
          if (q < n_fp && ok) acc[q] = fma(g, a.e[(int64_t)(q - 1) * a.n + i_lo + li], acc[q]);
      }
    } else {
#pragma unroll 4
      for (int64_t i = i_lo + t; i < i_hi; i += 256) {
        const uint32_t word = cw[(i - i_lo) >> 4];
        const double g = (double)((word >> (2 * ((i - i_lo) & 15))) & 3u);
        acc[0] = fma(g, __ldcg(a.y + i), acc[0]);
#pragma unroll
        for (int q = 1; q < 8; ++q)
          if (q < n_fp) acc[q] = fma(g, a.e[(int64_t)(q - 1) * a.n + i], acc[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      double v = acc[q];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) fpart[warp][q] = v;
    }
  }
  // (2) x_c'x_l for model columns and other candidates: exact integer popcount arithmetic.  A warp takes tasks
  //     warp, warp+8, ...; the words of FOUR tasks are loaded before any is used (one memory round per group).
  if (small) {
    for (int task0 = n_fp + warp; task0 < n_tasks; task0 += 4 * nw) {
      uint32_t ow[4][2];
      if (task0 == n_fp + warp) {   // first group: loaded before the barrier
#pragma unroll
        for (int u = 0; u < 4; ++u) { ow[u][0] = ow0[u][0]; ow[u][1] = ow0[u][1]; }
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int task = task0 + u * nw;
          const int q = task - n_fp;
          const uint32_t* other = task < n_tasks ? (q < a.k ? a.cols[a.m_c + q] : a.cols[q - a.k]) + w0 : nullptr;
#pragma unroll
          for (int j = 0; j < 2; ++j) ow[u][j] = (other != nullptr && lane + 32 * j < nwords) ? __ldcg(other + lane + 32 * j) : 0u;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int task = task0 + u * nw;
        int acc = 0;
#pragma unroll
        for (int j = 0; j < 2; ++j)
          if (lane + 32 * j < nwords) acc += packed_dot(cw[lane + 32 * j], ow[u][j]);
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0 && task < n_tasks) {
          if (part) part[task] = (double)acc;
          else publish_tagged(out + task, (double)acc, a.seq);
        }
      }
    }
  } else {
    for (int task = n_fp + warp; task < n_tasks; task += nw) {
      const int q = task - n_fp;
      const uint32_t* other = (q < a.k ? a.cols[a.m_c + q] : a.cols[q - a.k]) + w0;
      uint32_t ow[kFastSegMax / 32];
#pragma unroll
      for (int j = 0; j < kFastSegMax / 32; ++j) ow[j] = (lane + 32 * j < nwords) ? __ldcg(other + lane + 32 * j) : 0u;
      int acc = 0;
#pragma unroll
      for (int j = 0; j < kFastSegMax / 32; ++j)
        if (lane + 32 * j < nwords) acc += packed_dot(cw[lane + 32 * j], ow[j]);
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) {
        if (part) part[task] = (double)acc;
        else publish_tagged(out + task, (double)acc, a.seq);
      }
    }
  }
  __syncthreads();
  if (t < n_fp) {
    double v = 0.0;
    for (int wv = 0; wv < nw; ++wv) v += fpart[wv][t];
    if (part) part[t] = v;
    else publish_tagged(out + t, v, a.seq);
  }
  if (part == nullptr) return;
  // ---- the last item of the request adds the slices (in slice order) and publishes the sums
  __shared__ int is_last;
  __threadfence();   // this thread's slice results are visible device-wide ...
  __syncthreads();
  if (t == 0) {      // ... before the item is counted
    const unsigned int total = (unsigned int)(a.m_c * a.n_seg);
    unsigned int* tk = a.ticket + (a.seq & 3u);
    const unsigned int got = atomicAdd(tk, 1u);
    is_last = got == total - 1u;
    if (is_last) *tk = 0u;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // All slice results of a tile of outputs are fetched from L2 at once (every thread a few independent loads), then one
  // thread per output adds its n_seg values in slice order from shared memory.
  __shared__ double red[kReduceTile];
  const int n_out = a.m_c * n_tasks;                       // outputs (candidate, task), task fastest
  const int tile_out = max(1, kReduceTile / a.n_seg);      // outputs per tile
  for (int o0 = 0; o0 < n_out; o0 += tile_out) {
    const int cnt = min(tile_out, n_out - o0);
    if (a.n_seg <= kReduceTile) {
      for (int e = t; e < cnt * a.n_seg; e += blockDim.x) {
        const int g = e / cnt, o = o0 + (e - g * cnt);     // consecutive threads read consecutive tasks of one slice
        const int ci = o / n_tasks, task = o - ci * n_tasks;
        red[(e - g * cnt) * a.n_seg + g] = __ldcg(a.part + ((int64_t)ci * a.n_seg + g) * n_tasks + task);
      }
      __syncthreads();
      for (int o = t; o < cnt; o += blockDim.x) {
        const double* r = red + o * a.n_seg;
        double v = 0.0;
        for (int g = 0; g < a.n_seg; ++g) v += r[g];
        publish_tagged(a.out_host + o0 + o, v, a.seq);
      }
      __syncthreads();
    } else if (t == 0) {                                    // more slices than the tile holds (n > 4 M): plain loop
      const int ci = o0 / n_tasks, task = o0 - ci * n_tasks;
      double v = 0.0;
      for (int g = 0; g < a.n_seg; ++g) v += __ldcg(a.part + ((int64_t)ci * a.n_seg + g) * n_tasks + task);
      publish_tagged(a.out_host + o0, v, a.seq);
    }
  }
}

__global__ void __launch_bounds__(256) k_column_stats_inline(const __grid_constant__ ColStatInline a)
{
  __shared__ uint32_t cw[kFastSegMax];
  __shared__ double fpart[8][8];   // [warp][fp task], up to 8 fp tasks (y + 7 covariate columns) on the fast path
  colstat_item(a, blockIdx.x, blockIdx.y, cw, fpart);
}

// ---- persistent server for the latency path ----------------------------------------------------------------
// A kernel launch costs ~4-5 us before the first instruction runs; with one launch per MCMC move that is a third of
// the column-statistics round trip.  The server is ONE long-running kernel: CTA 0 polls a mailbox in mapped pinned
// host memory (the host posts a request by plain stores), forwards the request through device memory to the worker
// CTAs, and every CTA processes its share of the (candidate, slice) items exactly as k_column_stats_inline does,
// publishing the same self-validating tagged results.  The mailbox is made of 16-byte chunks that each carry the
// request's sequence number, so one warp-wide read returns a request that validates itself (no second PCIe round
// trip).  The server is stopped before every scan (it would take registers from the scan's CTAs) and exits by
// itself when no request arrives for kServerIdleCycles.
constexpr int kMailChunks = 1 + kInlinePtrs;            // chunk 0: {seq, m_c, k, n_seg}; chunk 1+i: {ptr lo, ptr hi, seq, 0}
static_assert(kMailChunks <= 256, "one mailbox chunk per thread of the poller CTA");
constexpr unsigned int kServerStop = 0xFFFFFFFFu;
constexpr long long kServerIdleCycles = 200000000ll;    // ~0.1 s: bounds what an unexpected implicit device synchronisation
                                                        // (cudaMalloc / cudaFreeHost somewhere in the process) can cost; the
                                                        // next request restarts the server

struct ServerArgs {
  const uint4* mail;        // host (mapped, pinned)
  uint4* dev_req;           // device: [2][kMailChunks]
  unsigned int* dev_flag;   // device: sequence number of the request in dev_req[seq & 1]
  ColStatInline base;       // everything but cols / m_c / k / n_seg / seq
};

__device__ __forceinline__ uint4 ld_sys_v4(const uint4* p)
{
  uint4 v;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p)
{
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned int* p, unsigned int v)
{
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __maxnreg__(96) k_colstats_server(const __grid_constant__ ServerArgs sa)
{
  __shared__ uint32_t cw[kFastSegMax];
  __shared__ double fpart[8][8];
  __shared__ ColStatInline req;
  __shared__ uint4 head;
  __shared__ unsigned int sh_seq;
  const int t = threadIdx.x;
  unsigned int last = ld_acquire_gpu(sa.dev_flag);   // whatever the previous server instance served last
  long long t_idle = clock64();
  for (int i = t; i < (int)(sizeof(ColStatInline) / sizeof(uint32_t)); i += blockDim.x)
    reinterpret_cast<uint32_t*>(&req)[i] = reinterpret_cast<const uint32_t*>(&sa.base)[i];
  __syncthreads();
  for (;;) {
    if (blockIdx.x == 0) {
      // ---- poller: one read of the mailbox per trip; forward a complete request to the workers
      for (;;) {
        // first trip: head + 63 columns (the usual request); a larger request costs a second trip for the rest
        uint4 ch = make_uint4(0, 0, 0, 0);
        if (t < 64) ch = ld_sys_v4(sa.mail + t);
        if (t == 0) head = ch;
        __syncthreads();
        const uint4 h = head;
        const unsigned int s = h.x;
        bool fresh = s != last && s != 0;
        int ok = 1;
        if (fresh && s != kServerStop) {
          const int needed = min(kMailChunks, 1 + (int)h.y + (int)h.z);
          if (t >= 64 && t < needed) ch = ld_sys_v4(sa.mail + t);
          ok = (t == 0 || t >= needed || ch.z == s) ? 1 : 0;
        }
        ok = __syncthreads_and(ok);
        if (fresh && ok) {
          if (s != kServerStop) {
            const int needed = min(kMailChunks, 1 + (int)h.y + (int)h.z);
            if (t < needed) sa.dev_req[(size_t)(s & 1u) * kMailChunks + t] = ch;
            __threadfence();
          }
          __syncthreads();
          if (t == 0) st_release_gpu(sa.dev_flag, s);
          break;
        }
        if (clock64() - t_idle > kServerIdleCycles) {   // abandoned: shut the whole server down
          if (t == 0) st_release_gpu(sa.dev_flag, kServerStop);
          __syncthreads();
          break;
        }
      }
    }
    // ---- every CTA: wait for the forwarded request
    if (t == 0) {
      unsigned int s;
      while ((s = ld_acquire_gpu(sa.dev_flag)) == last) {
        if (clock64() - t_idle > 2 * kServerIdleCycles) { s = kServerStop; break; }
      }
      sh_seq = s;
    }
    __syncthreads();
    const unsigned int s = sh_seq;
    if (s == kServerStop) return;
    const uint4* rq = sa.dev_req + (size_t)(s & 1u) * kMailChunks;
    const uint4 h = __ldcg(rq);   // forwarded through L2: never from a stale L1 line
    const int m_c = (int)h.y, k = (int)h.z, n_seg = (int)h.w;
    if (t == 0) { req.m_c = m_c; req.k = k; req.n_seg = n_seg; req.seq = s; }
    if (t < m_c + k) {
      const uint4 c4 = __ldcg(rq + 1 + t);
      req.cols[t] = reinterpret_cast<const uint32_t*>((unsigned long long)c4.x | ((unsigned long long)c4.y << 32));
    }
    __syncthreads();
    for (int item = blockIdx.x; item < m_c * n_seg; item += gridDim.x) {
      colstat_item(req, item / n_seg, item % n_seg, cw, fpart);
      __syncthreads();
    }
    last = s;
    t_idle = clock64();
  }
}

__global__ void k_sum_segments(const double* __restrict__ seg_out, int m_c, int n_seg, int n_tasks, double* __restrict__ out)
{
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= m_c * n_tasks) return;
  const int c = gid / n_tasks, task = gid % n_tasks;
  double s = 0.0;
  for (int g = 0; g < n_seg; ++g) s += seg_out[((int64_t)c * n_seg + g) * n_tasks + task];
  out[gid] = s;
}

// ---- host side of the persistent server ----------------------------------------------------------------------
// Chains of this process that serve their moves through a persistent kernel, per device.  Their servers must all be resident
// at once (a server CTA takes its work items by index and spins until its kernel is told to stop), and a scan CTA must still
// find room beside them: together they get one CTA per SM.  The reference runs thread.n_threads chains over one Data
// (src/main.cpp:54-108; its bundled testdata.ini: two); here these are n_threads chains on one GPU.
static std::atomic<int> g_server_chains[64];

void chain_expect_server(Chain* c)
{
  if (c->server_counted || !c->server_enabled) return;
  const int dev = c->store->device;
  if (dev < 0 || dev >= 64) return;
  g_server_chains[dev].fetch_add(1, std::memory_order_relaxed);
  c->server_counted = true;
}
void chain_forget_server(Chain* c)
{
  if (!c->server_counted) return;
  g_server_chains[c->store->device].fetch_sub(1, std::memory_order_relaxed);
  c->server_counted = false;
}

static void server_start(Chain* c, const ColStatInline& base, unsigned int last_served)
{
  Store* s = c->store;
  if (c->server_stream == nullptr) {
    BMG_CUDA(cudaStreamCreateWithFlags(&c->server_stream, cudaStreamNonBlocking));
    c->server_mail.alloc(kMailChunks);
    c->server_req.alloc(2 * kMailChunks);
    c->server_flag.alloc(1);
    memset(c->server_mail.p, 0, kMailChunks * sizeof(uint4));
  }
  chain_expect_server(c);   // a chain driven through bmg_chain_column_stats alone is counted from its first request on
  {
    // one CTA per SM over all the servers of this device: alone, a request of up to sm_count (candidate, slice) items is one wave
    const int dev = s->device;
    static const bool share = getenv("BMG_SERVER_SHARE") == nullptr || atoi(getenv("BMG_SERVER_SHARE")) != 0;   // development: 0 = sm_count CTAs each
    const int sharers = share && dev >= 0 && dev < 64 ? std::max(1, g_server_chains[dev].load(std::memory_order_relaxed)) : 1;
    c->server_ctas = std::max(8, s->sm_count / sharers);
  }
  // nothing is "fresh" until the next post: mailbox head and device flag both carry the last served sequence number
  volatile uint32_t* head = reinterpret_cast<volatile uint32_t*>(c->server_mail.p);
  head[0] = last_served;
  std::atomic_thread_fence(std::memory_order_seq_cst);
  const unsigned int flag0 = last_served;
  BMG_CUDA(cudaMemcpyAsync(c->server_flag.p, &flag0, sizeof(flag0), cudaMemcpyHostToDevice, c->server_stream));
  BMG_CUDA(cudaStreamSynchronize(c->server_stream));
  // The server shares its SMs with kernels launched while it lives (the scan service of a shard group, group.cu).  Two things
  // decide whether a scan CTA (7 warps of 80 registers, 53 kB of shared memory) finds room beside a server CTA: registers
  // -- each of an SM's four sub-partitions holds 16 K, and with 181 registers per server thread a sub-partition had room for
  // one scan warp where the CTA needs two, so every scan of another chain waited until the server was stopped (measured: up
  // to a full Rao-Blackwell period, 150 ms); hence __maxnreg__(96) -- and the L1 / shared-memory split, which cannot change
  // while a CTA is resident, hence the largest carve-out up front.
  static bool carveout_set = false;
  if (!carveout_set) {
    BMG_CUDA(cudaFuncSetAttribute(k_colstats_server, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    carveout_set = true;
  }
  ServerArgs sa;
  sa.mail = reinterpret_cast<const uint4*>(c->server_mail.p);
  sa.dev_req = reinterpret_cast<uint4*>(c->server_req.p);
  sa.dev_flag = c->server_flag.p;
  sa.base = base;
  k_colstats_server<<<c->server_ctas, 256, 0, c->server_stream>>>(sa);
  count_launch();
  const cudaError_t le = cudaGetLastError();
  if (le != cudaSuccess) throw Error(std::string("k_colstats_server launch: ") + cudaGetErrorString(le));
  c->server_running = true;
  c->server_base_y = base.y; c->server_seg_words = base.seg_words; c->server_base_out = (const void*)base.out_host;
}

void chain_server_stop(Chain* c)
{
  if (!c->server_running) return;
  static const bool timing = getenv("BMG_TIMING") != nullptr;
  struct timespec ta, tb;
  if (timing) clock_gettime(CLOCK_MONOTONIC, &ta);
  volatile uint32_t* head = reinterpret_cast<volatile uint32_t*>(c->server_mail.p);
  head[0] = kServerStop;
  std::atomic_thread_fence(std::memory_order_seq_cst);
  cudaStreamSynchronize(c->server_stream);
  c->server_running = false;
  if (timing) {
    clock_gettime(CLOCK_MONOTONIC, &tb);
    static int shown = 0;
    if (shown++ < 4) fprintf(stderr, "[bmg timing] column-statistics server stopped in %.1f us after %lld requests\n",
                             1e6 * ((tb.tv_sec - ta.tv_sec) + 1e-9 * (tb.tv_nsec - ta.tv_nsec)), (long long)c->server_requests);
  }
}

static void server_post(Chain* c, const ColStatInline& a)
{
  // the server may have timed out while the host was busy elsewhere
  if (c->server_running && cudaStreamQuery(c->server_stream) != cudaErrorNotReady) c->server_running = false;
  if (c->server_running && (c->server_base_y != a.y || c->server_seg_words != a.seg_words || c->server_base_out != (const void*)a.out_host))
    chain_server_stop(c);
  if (!c->server_running) server_start(c, a, a.seq - 1);   // a.seq is the request about to be posted
  volatile uint64_t* w = reinterpret_cast<volatile uint64_t*>(c->server_mail.p);
  const uint64_t tag = (uint64_t)a.seq;
  for (int i = 0; i < a.m_c + a.k; ++i) {   // pointer half first, tag half second (x86 keeps the store order)
    w[2 * (1 + i)] = (uint64_t)(uintptr_t)a.cols[i];
    w[2 * (1 + i) + 1] = tag;
  }
  std::atomic_thread_fence(std::memory_order_release);
  w[1] = (uint64_t)(uint32_t)a.k | ((uint64_t)(uint32_t)a.n_seg << 32);
  w[0] = tag | ((uint64_t)(uint32_t)a.m_c << 32);   // head last: {seq, m_c} | {k, n_seg}
  std::atomic_thread_fence(std::memory_order_seq_cst);
}

// the pending request again, as one launch of k_column_stats_inline (see chain_column_stats_wait)
static void server_fallback_inline(Chain* c)
{
  c->server_running = false;
  if (++c->server_failures >= 2) {
    c->server_enabled = false;
    if (getenv("BMG_TIMING"))
      fprintf(stderr, "[bmg timing] column-statistics server exited twice without serving: one launch per move from now on\n");
  }
  BMG_REQUIRE(c->cs_last_req.size() == sizeof(ColStatInline), "column statistics: no request to repeat");
  ColStatInline a;
  memcpy(&a, c->cs_last_req.data(), sizeof(a));
  if (a.ticket) BMG_CUDA(cudaMemsetAsync(a.ticket, 0, 4 * sizeof(unsigned int), c->stream));   // the dead instance may have counted some items
  k_column_stats_inline<<<dim3(a.m_c, a.n_seg), 256, 0, c->stream>>>(a);
  count_launch();
  const cudaError_t le = cudaGetLastError();
  if (le != cudaSuccess) throw Error(std::string("k_column_stats_inline launch: ") + cudaGetErrorString(le));
  ++c->server_fallbacks;
}

// second half of the latency path: spin on the completion flag, then reduce the per-slice partials on the host
void chain_column_stats_wait(Chain* c, double* xy, double* xe, double* xx_model, double* xx_cand)
{
  if (!c->cs_pending) return;
  Store* s = c->store;
  cudaStream_t st = c->stream;
  const int m_c = c->cs_p_mc, k = c->cs_p_k, n_seg = c->cs_p_nseg;
  const int n_tasks = s->m_e + 1 + k + m_c;
  static const bool timing = getenv("BMG_TIMING") != nullptr;
  static long long hist[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  static long long hist_n = 0;
  struct timespec ta;
  if (timing) clock_gettime(CLOCK_MONOTONIC, &ta);
  const unsigned long long tag = (unsigned long long)c->cs_p_seq << 32;
  const volatile unsigned long long* words = reinterpret_cast<const volatile unsigned long long*>(c->cs_map.p);
  unsigned long spins = 0;
  const bool device_reduced = c->cs_p_reduced;
  const int n_host_seg = device_reduced ? 1 : n_seg;
  for (int ci = 0; ci < m_c; ++ci) {
    for (int task = 0; task < n_tasks; ++task) {
      double v = 0.0;
      for (int g = 0; g < n_host_seg; ++g) {
        const size_t slot = ((size_t)ci * n_host_seg + g) * n_tasks + task;
        unsigned long long w0, w1;
        for (;;) {
          w0 = words[2 * slot];
          w1 = words[2 * slot + 1];
          if ((w0 & 0xFFFFFFFF00000000ull) == tag && (w1 & 0xFFFFFFFF00000000ull) == tag) break;
          cudaStream_t qs = c->server_running ? c->server_stream : st;
          if ((++spins & 0xFFFFF) == 0 && cudaStreamQuery(qs) != cudaErrorNotReady) {   // finished (or failed) without publishing?
            BMG_CUDA(cudaStreamSynchronize(qs));
            w0 = words[2 * slot];
            w1 = words[2 * slot + 1];
            if ((w0 & 0xFFFFFFFF00000000ull) == tag && (w1 & 0xFFFFFFFF00000000ull) == tag) break;
            if (!c->server_running) throw Error("k_column_stats_inline finished without publishing its results");
            // The server instance is gone and this request was never served: it was posted into the window between the
            // poller's last mailbox read and the kernel's exit (idle time-out), or a serialising tool (ncu,
            // compute-sanitizer, cuda-gdb) ran the persistent kernel alone until it timed out.  Serve the SAME request
            // (same tag, same result words) with one ordinary launch on the chain's stream and keep polling; after two
            // failures in a row the chain stops using the server altogether.
            server_fallback_inline(c);
          }
        }
        const unsigned long long bits = (w0 & 0xFFFFFFFFull) | (w1 << 32);
        double part;
        memcpy(&part, &bits, sizeof(part));
        v += part;
      }
      if (task == 0) { if (xy) xy[ci] = v; }
      else if (task <= s->m_e) { if (xe) xe[(size_t)ci * s->m_e + task - 1] = v; }
      else if (task <= s->m_e + k) { if (xx_model) xx_model[(size_t)ci * k + task - 1 - s->m_e] = v; }
      else if (xx_cand) xx_cand[(size_t)ci * m_c + task - 1 - s->m_e - k] = v;
    }
  }
  c->cs_pending = false;
  if (c->server_running) c->server_failures = 0;
  if (timing) {
    struct timespec tb;
    clock_gettime(CLOCK_MONOTONIC, &tb);
    const double us = 1e6 * ((tb.tv_sec - ta.tv_sec) + 1e-9 * (tb.tv_nsec - ta.tv_nsec));
    const double edges[7] = {5, 10, 20, 50, 200, 2000, 50000};
    int b = 0;
    while (b < 7 && us >= edges[b]) ++b;
    ++hist[b];
    if ((++hist_n % 3000) == 0)
      fprintf(stderr, "[bmg timing] column-stats wait histogram (us) <5:%lld <10:%lld <20:%lld <50:%lld <200:%lld <2000:%lld <50000:%lld more:%lld\n",
              hist[0], hist[1], hist[2], hist[3], hist[4], hist[5], hist[6], hist[7]);
  }
}

void chain_column_stats(Chain* c, const int64_t* cand, int m_c, const int64_t* loci, int k, double* xy, double* xe,
                        double* xx_model, double* xx_cand)
{
  chain_column_stats_launch(c, cand, m_c, loci, k, xy, xe, xx_model, xx_cand, false);
}

// launch_only: on the latency path return right after the launch (results via chain_column_stats_wait); the
// general path always completes before returning
void chain_column_stats_launch(Chain* c, const int64_t* cand, int m_c, const int64_t* loci, int k, double* xy, double* xe,
                               double* xx_model, double* xx_cand, bool launch_only)
{
  BMG_REQUIRE(!c->cs_pending, "column statistics: a previous launch has not been collected");
  Store* s = c->store;
  BMG_REQUIRE(m_c >= 1 && m_c <= 256, "bmg_chain_column_stats: 1..256 candidates per call");
  BMG_REQUIRE(k >= 0 && k <= 2048, "bmg_chain_column_stats: model size must be <= 2048");
  BMG_CUDA(cudaSetDevice(s->device));
  cudaStream_t st = c->stream;
  const int n_tasks = s->m_e + 1 + k + m_c;
  const int n_seg = (int)((s->W + kSegWords - 1) / kSegWords);
  // columns as the chain sees them: a SNP with missing calls is read from its patched copy (overlay.cu), built here
  // if it is not cached; every kernel below then treats all columns alike
  std::vector<int64_t>& involved = c->cs_involved;
  involved.assign(cand, cand + m_c);
  involved.insert(involved.end(), loci, loci + k);
  std::vector<const uint32_t*>& colp = c->cs_colp;
  colp.resize(m_c + k);
  const int patched = chain_overlay_columns(c, involved.data(), m_c + k, colp.data());
  if (m_c + k <= kInlinePtrs && s->m_e + 1 <= 8 && getenv("BMG_COLSTATS_SLOW") == nullptr) {
    // many small CTAs (one per candidate and 1024- or 4096-individual slice): one round of memory latency each
    static const int seg_forced = getenv("BMG_COLSTATS_SEG") ? atoi(getenv("BMG_COLSTATS_SEG")) : 0;   // development: 64 or 256
    const int seg_words = seg_forced == 64 || seg_forced == 256 ? seg_forced : (s->W <= 4096 ? 64 : kFastSegMax);
    const int n_seg = (int)((s->W + seg_words - 1) / seg_words);
    const size_t need_fast = (size_t)m_c * n_tasks * n_seg;
    if (c->cs_map.n < 2 * need_fast + 8 || c->cs_seq >= 0xFFFFFFF0u) {   // (re)allocate; also before the tag wraps around
      chain_server_stop(c);   // cudaFreeHost / cudaHostAlloc synchronise the device: a running server would stall them
      BMG_CUDA(cudaStreamSynchronize(st));
      if (c->cs_map.n < 2 * need_fast + 8) c->cs_map.alloc(need_fast * 4 + 64);
      memset(c->cs_map.p, 0, c->cs_map.n * sizeof(double));   // tag 0 is never used by a launch
      if (c->cs_seq >= 0xFFFFFFF0u) { chain_server_stop(c); c->cs_seq = 0; }
    }
    ColStatInline a;
    for (int i = 0; i < m_c + k; ++i) a.cols[i] = colp[i];
    a.m_c = m_c; a.k = k; a.m_e = s->m_e; a.n = s->n; a.W = s->W; a.n_seg = n_seg; a.seg_words = seg_words; a.y = c->y.p; a.e = s->e.p;
    a.out_host = reinterpret_cast<ulonglong2*>(c->cs_map.p);
    a.seq = ++c->cs_seq;
    a.part = nullptr; a.ticket = nullptr;
    static const bool host_reduce_forced = getenv("BMG_COLSTATS_HOSTREDUCE") != nullptr;   // development: A/B of the two ways
    if (n_seg > kHostReduceSegs && !host_reduce_forced) {
      if (c->cs_part.n < need_fast || c->cs_ticket.n == 0) {
        chain_server_stop(c);
        BMG_CUDA(cudaStreamSynchronize(st));
        if (c->cs_part.n < need_fast) c->cs_part.alloc(need_fast * 2);
        if (c->cs_ticket.n == 0) { c->cs_ticket.alloc(4); BMG_CUDA(cudaMemset(c->cs_ticket.p, 0, 4 * sizeof(unsigned int))); BMG_CUDA(cudaDeviceSynchronize()); }
      }
      a.part = c->cs_part.p; a.ticket = c->cs_ticket.p;
    }
    c->cs_p_reduced = a.part != nullptr;
    if (c->server_enabled) {
      if (patched > 0) BMG_CUDA(cudaStreamSynchronize(st));   // the server runs on its own stream: columns must be complete
      c->cs_last_req.resize(sizeof(ColStatInline));            // kept so that the wait can repeat it with a plain launch
      memcpy(c->cs_last_req.data(), &a, sizeof(a));
      server_post(c, a);
      ++c->server_requests;
    } else {
      k_column_stats_inline<<<dim3(m_c, n_seg), 256, 0, st>>>(a);
      count_launch();
      const cudaError_t le = cudaGetLastError();
      if (le != cudaSuccess) throw Error(std::string("k_column_stats_inline launch: ") + cudaGetErrorString(le));
    }
    g_d2h_bytes.fetch_add((a.part ? (size_t)m_c * n_tasks : need_fast) * 2 * sizeof(double), std::memory_order_relaxed);
    g_h2d_bytes.fetch_add(sizeof(ColStatInline), std::memory_order_relaxed);
    c->cs_pending = true;
    c->cs_p_mc = m_c; c->cs_p_k = k; c->cs_p_nseg = n_seg; c->cs_p_seq = a.seq;
    if (!launch_only) chain_column_stats_wait(c, xy, xe, xx_model, xx_cand);
    return;
  }
  chain_server_stop(c);   // the general path allocates and synchronises
  const size_t need = (size_t)m_c * n_tasks * (n_seg + 1);
  if (c->cs_out.n < need) { c->cs_out.alloc(need * 2); c->h_cs.alloc(need * 2); }
  const size_t n_ptr = (size_t)(m_c + k);
  if (c->cs_idx.n < n_ptr + 16) c->cs_idx.alloc(n_ptr * 2 + 4096);
  if (c->h_stage_i.n < n_ptr + 16) c->h_stage_i.alloc(n_ptr * 2 + 4096);
  BMG_CUDA(cudaStreamSynchronize(st));
  int64_t* hp = c->h_stage_i.p;
  for (int i = 0; i < m_c + k; ++i) hp[i] = (int64_t)(uintptr_t)colp[i];
  bmg::copy_h2d(c->cs_idx.p, hp, n_ptr * sizeof(int64_t), st);
  ColStatArgs a;
  a.cand_cols = reinterpret_cast<const uint32_t* const*>(c->cs_idx.p);
  a.model_cols = reinterpret_cast<const uint32_t* const*>(c->cs_idx.p + m_c);
  a.m_c = m_c; a.k = k; a.m_e = s->m_e; a.n = s->n; a.W = s->W; a.n_seg = n_seg; a.y = c->y.p; a.e = s->e.p;
  double* seg_out = c->cs_out.p + (size_t)m_c * n_tasks;
  double* fin = c->cs_out.p;
  a.out = n_seg == 1 ? fin : seg_out;
  k_column_stats<<<dim3(m_c, n_seg), 256, 0, st>>>(a);
  count_launch();
  if (n_seg > 1) {
    k_sum_segments<<<(m_c * n_tasks + 127) / 128, 128, 0, st>>>(seg_out, m_c, n_seg, n_tasks, fin);
    count_launch();
  }
  BMG_CUDA(cudaGetLastError());
  bmg::copy_d2h(c->h_cs.p, fin, (size_t)m_c * n_tasks * sizeof(double), st);
  BMG_CUDA(cudaStreamSynchronize(st));
  for (int ci = 0; ci < m_c; ++ci) {
    const double* row = c->h_cs.p + (size_t)ci * n_tasks;
    if (xy) xy[ci] = row[0];
    if (xe) for (int j = 0; j < s->m_e; ++j) xe[(size_t)ci * s->m_e + j] = row[1 + j];
    if (xx_model) for (int l = 0; l < k; ++l) xx_model[(size_t)ci * k + l] = row[1 + s->m_e + l];
    if (xx_cand) for (int d = 0; d < m_c; ++d) xx_cand[(size_t)ci * m_c + d] = row[1 + s->m_e + k + d];
  }
}

// ---------------------------------------------------------------------------------------
// probit latent update (no reference counterpart)
// ---------------------------------------------------------------------------------------
// partial[block][2 + m_e] = { sum z, sum z^2, sum z E_j ... } (m_e <= 8 here; more covariates fall back to 2 sums)
__global__ void __launch_bounds__(256) k_probit(const double* __restrict__ ye, const double* __restrict__ yg,
                                                const uint8_t* __restrict__ is_case, const double* __restrict__ u_in,
                                                uint64_t seed, uint64_t counter, int64_t n, const double* __restrict__ e, int m_e,
                                                double* __restrict__ y, double* __restrict__ partial)
{
  __shared__ double sm[10][8];
  double acc[10];
#pragma unroll
  for (int q = 0; q < 10; ++q) acc[q] = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double mu = yg[i] + ye[i];
    double u;
    if (u_in) u = u_in[i];
    else { Philox g(seed, counter, (uint64_t)i); u = g.u01(); }
    // inverse-CDF draw from N(mu,1) truncated to (0,inf) [case] or (-inf,0] [control]: with s = mu (case) or -mu (control),
    // t = Phi^-1(u Phi(s)) and z = mu -/+ t.  The quantile is taken in whichever tail its argument is small in: for
    // u Phi(s) > 1/2 the complement (1 - u) + u Phi(-s) is formed without cancellation (1 - u is exact for u >= 1/2), so
    // the draw keeps its relative accuracy for |mu| up to 8 and u within 1e-12 of either end (tests: exact quantiles).
    const double sgn = is_case[i] ? 1.0 : -1.0, s = sgn * mu;
    const double p = u * normcdf(s);
    const double t = p <= 0.5 ? normcdfinv(p) : -normcdfinv((1.0 - u) + u * normcdf(-s));
    const double z = mu - sgn * t;
    y[i] = z;
    acc[0] += z;
    acc[1] += z * z;
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (q < m_e) acc[2 + q] = fma(z, e[(int64_t)q * n + i], acc[2 + q]);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 10; ++q) {
    double v = acc[q];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sm[q][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 2 + m_e) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sm[threadIdx.x][w];
    partial[(int64_t)blockIdx.x * (2 + m_e) + threadIdx.x] = s;
  }
}

__global__ void k_reduce_final2(const double* __restrict__ partial, int blocks, int nq, double* __restrict__ out)
{
  const int q = threadIdx.x;
  if (q >= nq) return;
  double s = 0.0;
  for (int b = 0; b < blocks; ++b) s += partial[(int64_t)b * nq + q];
  out[q] = s;
}

// stats: { sum z, sum z^2, E'z (m_e entries, when m_e <= 8) }
void chain_probit_update(Chain* c, const uint8_t* is_case, const double* u01, uint64_t seed, uint64_t counter,
                         double* stats2, double* ez)
{
  Store* s = c->store;
  BMG_REQUIRE(c->residual_valid, "bmg_chain_probit_update: call bmg_chain_residual first (it leaves the fitted values)");
  BMG_CUDA(cudaSetDevice(s->device));
  cudaStream_t st = c->stream;
  if (is_case) {
    if (c->is_case.n == 0) c->is_case.alloc(s->n);
    bmg::copy_h2d(c->is_case.p, is_case, s->n, st);
    c->have_case = true;
  }
  BMG_REQUIRE(c->have_case, "bmg_chain_probit_update: case/control labels were never set");
  DevBuf<double> u_dev;
  if (u01) {
    u_dev.alloc(s->n);
    bmg::copy_h2d(u_dev.p, u01, s->n * sizeof(double), st);
  }
  const int blocks = (int)std::min<int64_t>(1024, (s->n + 255) / 256);
  const int me = s->m_e <= 8 ? s->m_e : 0;
  k_probit<<<blocks, 256, 0, st>>>(c->yhat_e.p, c->yhat_g.p, c->is_case.p, u01 ? u_dev.p : nullptr, seed, counter, s->n, s->e.p, me,
                                   c->y.p, c->red_partial.p);
  k_reduce_final2<<<1, 32, 0, st>>>(c->red_partial.p, blocks, 2 + me, c->red_out.p);
  count_launch(2);
  BMG_CUDA(cudaGetLastError());
  bmg::copy_d2h(c->h_red.p, c->red_out.p, (2 + me) * sizeof(double), st);
  BMG_CUDA(cudaStreamSynchronize(st));
  c->residual_valid = false;  // the phenotype changed: the residual must be rebuilt
  c->imma_q_valid = false;
  if (stats2) { stats2[0] = c->h_red.p[0]; stats2[1] = c->h_red.p[1]; }
  if (ez) {
    BMG_REQUIRE(me == s->m_e, "probit update: E'z is returned for at most 8 covariate columns");
    for (int q = 0; q < me; ++q) ez[q] = c->h_red.p[2 + q];
  }
}

}  // namespace bmg
