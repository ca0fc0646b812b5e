// scan.cu -- fitted values / residual, the all-SNP Rao-Blackwell genotype scan, its epilogue.
//
// Replaces RaoBlackwellizer::p_raoblackwell (src/sampler.cpp:32-261): for every SNP j of the
// shard  dot_j = sum_i x_ij r_i  over the packed 2-bit column, then the O(1) per-SNP algebra
// (centring with the cached moments, tau, sigma2, model-prior change -> p_r[j]).
//
// Roofline: HBM-bound packed GEMV co-limited by the FP64 pipe (one DFMA per genotype).
// The kernel keeps the non-DFMA instruction count per genotype at ~1:
//   * the residual lives in REGISTERS (each thread owns 32 fixed individuals for the whole
//     kernel), so no load instruction is needed per genotype;
//   * a genotype is turned into an fp64 operand with ONE integer instruction and no shift:
//     the 2-bit value field at bits [2p+1:2p] of the packed word, masked in place, IS the low
//     word of a denormal double  c * 2^(2p) * 2^-1074 ; the residual of the individual at
//     position p is pre-scaled by the exact power of two 2^(960-2p), so every product equals
//     c * r * 2^-114 exactly inside the FMA and all 32 genotypes of a thread accumulate into one
//     register;  the final sum is multiplied by 2^114 (exact).
//   * packed tiles are staged global -> shared with bulk-async copies (cp.async.bulk + mbarrier,
//     the TMA engine; SASS UBLKCP) in a 3-stage ring, so no registers are spent on prefetch.
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include "common.cuh"
#include "store.cuh"
#include "group.cuh"
#include "philox.cuh"

namespace bmg {

// exponent of the residual pre-scaling; products carry 2^(RSCALE-1074)
#define BMG_RSCALE 960
#define BMG_UNSCALE 114  // 1074 - 960

constexpr int kTileSnps = 16;   // SNPs per pipeline stage
constexpr int kStages = 3;
constexpr int kBatch = 8;       // SNPs accumulated per warp-transpose
constexpr int kMaxWarps = 8;

// ---------------------------------------------------------------------------------------
// PTX helpers: mbarrier + bulk async copy (Blackwell/Hopper TMA engine, non-tensor form)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ uint2 ldg_stream_u2(const uint32_t* p)
{
  uint2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}

// genotype field -> fp64 operand: (w & (3 << 2p)) read as the low word of a double with a zero
// high word is the denormal  c * 2^(2p) * 2^-1074  (c = 0, 1, 2)
__device__ __forceinline__ double field_as_double(uint32_t w, int p)
{
  return __hiloint2double(0, (int)(w & (3u << (2 * p))));
}

// one genotype: mask the 2-bit field in place, view it as the low word of a double whose high word is 0
// (a denormal), and accumulate.  volatile keeps the interleaved issue order written below, so that
// consecutive DFMAs belong to different accumulators (the FP64 pipe has a long dependent latency).
__device__ __forceinline__ void fma_field(double& acc, uint32_t w, uint32_t mask, double r)
{
  asm volatile(
      "{\n"
      ".reg .b32 t;\n"
      ".reg .f64 x;\n"
      "and.b32 t, %1, %2;\n"
      "mov.b64 x, {t, %3};\n"
      "fma.rn.f64 %0, x, %4, %0;\n"
      "}\n"
      : "+d"(acc)
      : "r"(w), "r"(mask), "r"(0), "d"(r));
}

// 32 genotypes (two packed words) times 32 register-resident pre-scaled residuals
__device__ __forceinline__ void fma_words(double (&acc)[kBatch], const uint2 (&w)[kBatch], const double (&rp)[32])
{
#pragma unroll
  for (int p = 0; p < 16; ++p) {
#pragma unroll
    for (int i = 0; i < kBatch; ++i) fma_field(acc[i], w[i].x, 3u << (2 * p), rp[p]);
  }
#pragma unroll
  for (int p = 0; p < 16; ++p) {
#pragma unroll
    for (int i = 0; i < kBatch; ++i) fma_field(acc[i], w[i].y, 3u << (2 * p), rp[16 + p]);
  }
}

// butterfly transpose-reduce of 8 per-lane sums over the 32 lanes.  Slot i of a lane holds SNP
// (i ^ x) of the batch with x = lane bits 4..2, so partners always exchange the SAME SNP and no
// select instructions are needed.  On return acc[0] of every lane holds the warp total of SNP x.
__device__ __forceinline__ void warp_transpose_reduce(double (&acc)[kBatch])
{
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i + 4], 16);
#pragma unroll
  for (int i = 0; i < 2; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i + 2], 8);
  acc[0] += __shfl_xor_sync(0xffffffffu, acc[1], 4);
  acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], 2);
  acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], 1);
}

struct ScanArgs {
  const uint32_t* codes;   // local shard, stride Wp words
  int64_t Wp;              // column stride in words
  int64_t m;               // local SNPs
  const double* r_scaled;  // 16 * n_chunks * chunk_words doubles, zero beyond n
  int chunk_words;         // words of a column one CTA covers (64 * warps)
  int n_chunks;
  int64_t tiles;           // ceil(m / kTileSnps)
  int slices;              // CTAs per chunk
  double* out;             // [n_chunks][m]
};

// ---------------------------------------------------------------------------------------
// variant 1: bulk-async staged.   dynamic smem: kStages * kTileSnps * row_words words, then
// partial sums [2][kMaxWarps][kTileSnps] doubles, then kStages mbarriers.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * kMaxWarps, 2) k_scan_dots_tma(const ScanArgs a)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nwarps = blockDim.x >> 5;
  const int chunk = blockIdx.x % a.n_chunks, slice = blockIdx.x / a.n_chunks;
  const bool whole = (a.n_chunks == 1);
  const int row_words = whole ? (int)a.Wp : a.chunk_words;
  const int64_t c0 = (int64_t)chunk * a.chunk_words;  // first word of this chunk in a column
  const int row_copy_words = whole ? (int)a.Wp : (int)min((int64_t)a.chunk_words, a.Wp - c0);
  const int stage_words = kTileSnps * row_words;

  uint32_t* stage0 = reinterpret_cast<uint32_t*>(smem_raw);
  double* part = reinterpret_cast<double*>(smem_raw + (size_t)kStages * stage_words * 4);
  uint64_t* full = reinterpret_cast<uint64_t*>(part + 2 * kMaxWarps * kTileSnps);

  const int64_t tile_lo = a.tiles * slice / a.slices, tile_hi = a.tiles * (slice + 1) / a.slices;
  const int64_t my_tiles = tile_hi - tile_lo;

  if (t == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // producer: one thread issues the copies of a tile
  auto issue = [&](int64_t it) {
    const int s = (int)(it % kStages);
    const int64_t snp0 = (tile_lo + it) * kTileSnps;
    const int rows = (int)min((int64_t)kTileSnps, a.m - snp0);
    uint32_t* dst = stage0 + (size_t)s * stage_words;
    if (whole) {
      const uint32_t bytes = (uint32_t)rows * (uint32_t)row_words * 4u;
      mbar_expect_tx(&full[s], bytes);
      bulk_g2s(dst, a.codes + snp0 * a.Wp, bytes, &full[s]);
    } else {
      const uint32_t rb = (uint32_t)row_copy_words * 4u;
      mbar_expect_tx(&full[s], rb * (uint32_t)rows);
      for (int rr = 0; rr < rows; ++rr) bulk_g2s(dst + (size_t)rr * row_words, a.codes + (snp0 + rr) * a.Wp + c0, rb, &full[s]);
    }
  };
  if (t == 0)
    for (int64_t it = 0; it < min((int64_t)kStages, my_tiles); ++it) issue(it);

  // the 32 individuals this thread owns for the whole kernel
  double rp[32];
  {
    const double* src = a.r_scaled + 16 * (c0 + 2 * t);
#pragma unroll
    for (int p = 0; p < 32; ++p) rp[p] = src[p];
  }
  const int x = (lane >> 2) & 7;  // slot permutation of warp_transpose_reduce

  for (int64_t it = 0; it < my_tiles; ++it) {
    const int s = (int)(it % kStages);
    mbar_wait(&full[s], (uint32_t)((it / kStages) & 1));
    const uint32_t* sm = stage0 + (size_t)s * stage_words + 2 * t;
    double* mypart = part + (size_t)(it & 1) * kMaxWarps * kTileSnps + warp * kTileSnps;
#pragma unroll
    for (int b = 0; b < kTileSnps / kBatch; ++b) {
      uint2 w[kBatch];
      double acc[kBatch];
#pragma unroll
      for (int i = 0; i < kBatch; ++i) {
        w[i] = *reinterpret_cast<const uint2*>(sm + (size_t)(b * kBatch + (i ^ x)) * row_words);
        acc[i] = 0.0;
      }
      fma_words(acc, w, rp);
      warp_transpose_reduce(acc);
      if ((lane & 3) == 0) mypart[b * kBatch + x] = acc[0];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy reads of stage s before its async-proxy refill
    __syncthreads();  // partials visible; every thread is done reading stage s
    if (t == 0 && it + kStages < my_tiles) issue(it + kStages);
    if (t < kTileSnps) {
      const int64_t snp = (tile_lo + it) * kTileSnps + t;
      if (snp < a.m) {
        const double* pp = part + (size_t)(it & 1) * kMaxWarps * kTileSnps + t;
        double sum = 0.0;
        for (int wv = 0; wv < nwarps; ++wv) sum += pp[wv * kTileSnps];
        a.out[(int64_t)chunk * a.m + snp] = scalbn(sum, BMG_UNSCALE);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// variant 0: direct vectorised global loads with a register double buffer (no smem staging)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * kMaxWarps, 2) k_scan_dots_ldg(const ScanArgs a)
{
  __shared__ double part[2 * kMaxWarps * kTileSnps];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nwarps = blockDim.x >> 5;
  const int chunk = blockIdx.x % a.n_chunks, slice = blockIdx.x / a.n_chunks;
  const int64_t c0 = (int64_t)chunk * a.chunk_words;
  const int64_t tile_lo = a.tiles * slice / a.slices, tile_hi = a.tiles * (slice + 1) / a.slices;
  const int64_t batches = (tile_hi - tile_lo) * (kTileSnps / kBatch);
  const int64_t snp_base = tile_lo * kTileSnps;
  const int64_t my_w = c0 + 2 * t;
  // threads whose words lie beyond the padded column read word 0 instead (their residuals are 0)
  const int64_t w_off = (my_w + 1 < a.Wp) ? my_w : 0;

  double rp[32];
  {
    const double* src = a.r_scaled + 16 * my_w;
#pragma unroll
    for (int p = 0; p < 32; ++p) rp[p] = src[p];
  }
  const int x = (lane >> 2) & 7;

  auto load_batch = [&](int64_t bi, uint2 (&w)[kBatch]) {
#pragma unroll
    for (int i = 0; i < kBatch; ++i) {
      int64_t snp = snp_base + bi * kBatch + (i ^ x);
      snp = snp < a.m ? snp : a.m - 1;
      w[i] = ldg_stream_u2(a.codes + snp * a.Wp + w_off);
    }
  };

  uint2 wn[kBatch];
  if (batches > 0) load_batch(0, wn);
  for (int64_t bi = 0; bi < batches; ++bi) {
    uint2 w[kBatch];
    double acc[kBatch];
#pragma unroll
    for (int i = 0; i < kBatch; ++i) {
      w[i] = wn[i];
      acc[i] = 0.0;
    }
    if (bi + 1 < batches) load_batch(bi + 1, wn);
    fma_words(acc, w, rp);
    warp_transpose_reduce(acc);
    double* mypart = part + (size_t)(bi & 1) * kMaxWarps * kTileSnps + warp * kTileSnps;
    if ((lane & 3) == 0) mypart[x] = acc[0];
    __syncthreads();
    if (t < kBatch) {
      const int64_t snp = snp_base + bi * kBatch + t;
      if (snp < a.m) {
        const double* pp = part + (size_t)(bi & 1) * kMaxWarps * kTileSnps + t;
        double sum = 0.0;
        for (int wv = 0; wv < nwarps; ++wv) sum += pp[wv * kTileSnps];
        a.out[(int64_t)chunk * a.m + snp] = scalbn(sum, BMG_UNSCALE);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// fitted values and residual
// ---------------------------------------------------------------------------------------
struct YhatArgs {
  const uint32_t* const* cols;  // k device pointers to packed columns (local or peer)
  const double* beta_g;         // k
  const double* beta_e;         // m_e
  const double* e;              // n x m_e col-major
  int k, m_e;
  int64_t n, W;
  double* yhat_e;
  double* yhat_g;
};

// one thread per packed word (16 individuals)
__global__ void k_yhat(const YhatArgs a)
{
  const int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (w >= a.W) return;
  double g[16];
#pragma unroll
  for (int p = 0; p < 16; ++p) g[p] = 0.0;
  for (int l = 0; l < a.k; ++l) {
    const uint32_t word = a.cols[l][w];
    const double b = a.beta_g[l];
#pragma unroll
    for (int p = 0; p < 16; ++p) g[p] = fma(b, (double)((word >> (2 * p)) & 3u), g[p]);
  }
  const int64_t i0 = 16 * w;
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    const int64_t i = i0 + p;
    if (i < a.n) {
      double ev = 0.0;
      for (int c = 0; c < a.m_e; ++c) ev = fma(a.beta_e[c], a.e[(int64_t)c * a.n + i], ev);
      a.yhat_e[i] = ev;
      a.yhat_g[i] = g[p];
    }
  }
}

constexpr int kRed = 9;

__device__ __forceinline__ double warp_sum(double v)
{
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// r = y - (yhat_g + yhat_e); r_scaled; 9 block-partial reductions (fixed order => deterministic)
__global__ void __launch_bounds__(256) k_residual(const double* __restrict__ y, const double* __restrict__ ye,
                                                  const double* __restrict__ yg, int64_t n, int64_t n_pad,
                                                  double* __restrict__ r, double* __restrict__ r_scaled,
                                                  double* __restrict__ partial)
{
  __shared__ double sm[kRed][8];
  double acc[kRed];
#pragma unroll
  for (int q = 0; q < kRed; ++q) acc[q] = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_pad; i += (int64_t)gridDim.x * blockDim.x) {
    if (i < n) {
      const double e = ye[i], g = yg[i], yh = g + e, rv = y[i] - yh;
      r[i] = rv;
      r_scaled[i] = scalbn(rv, BMG_RSCALE - 2 * (int)(i & 15));
      acc[0] += rv;
      acc[1] += e;  acc[2] += e * e;
      acc[3] += g;  acc[4] += g * g;
      acc[5] += yh; acc[6] += yh * yh;
      acc[7] += g * g;
      acc[8] += (e - y[i]) * g;
    } else {
      r_scaled[i] = 0.0;
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < kRed; ++q) {
    const double v = warp_sum(acc[q]);
    if (lane == 0) sm[q][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < kRed) {
    double s = 0.0;
    for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) s += sm[threadIdx.x][wv];
    partial[(int64_t)blockIdx.x * kRed + threadIdx.x] = s;
  }
}

__global__ void k_reduce_final(const double* __restrict__ partial, int blocks, int nq, double* __restrict__ out)
{
  const int q = threadIdx.x;
  if (q >= nq) return;
  double s = 0.0;
  for (int b = 0; b < blocks; ++b) s += partial[(int64_t)b * nq + q];
  out[q] = s;
}

// ---------------------------------------------------------------------------------------
// missing-cell corrections of the scan: one warp per SNP that has missing cells
// out[3j..] = { sum val * r[idx], sum val, sum val^2 }
// ---------------------------------------------------------------------------------------
__global__ void k_miss_corr(const int64_t* __restrict__ off, const int32_t* __restrict__ idx,
                            const int8_t* __restrict__ val, const double* __restrict__ r, int64_t m,
                            double* __restrict__ out)
{
  const int64_t j = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (j >= m) return;
  const int64_t lo = off[j], hi = off[j + 1];
  if (hi == lo) return;
  double d = 0.0, s1 = 0.0, s2 = 0.0;
  for (int64_t q = lo + lane; q < hi; q += 32) {
    const double v = (double)val[q];
    d += v * r[idx[q]];
    s1 += v;
    s2 += v * v;
  }
  d = warp_sum(d); s1 = warp_sum(s1); s2 = warp_sum(s2);
  if (lane == 0) { out[3 * j] = d; out[3 * j + 1] = s1; out[3 * j + 2] = s2; }
}

// ---------------------------------------------------------------------------------------
// per-SNP algebra of the scan (src/sampler.cpp:90-206, n_types == 1)
// ---------------------------------------------------------------------------------------
struct FinalizeArgs {
  const double* dot_partial; int n_chunks; int64_t m, lo, n;
  const int32_t* n1; const int32_t* n2;
  const double* miss_corr;  // may be null
  const int64_t* loci; const double* beta_g; const double* tau_g; int k;
  double sum_r, sigma2, lmp_add, lmp_rem;
  int tau_mode; double tau_shared; const double* tau_snp;
  uint64_t seed, counter; double nu_tau2, s2_tau2, alpha2;
  double* dot; double* p_r;
};

__global__ void __launch_bounds__(256) k_scan_finalize(const FinalizeArgs a)
{
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= a.m) return;
  double dot = 0.0;
  for (int c = 0; c < a.n_chunks; ++c) dot += a.dot_partial[(int64_t)c * a.m + j];
  double s = (double)(a.n1[j] + 2 * a.n2[j]);
  double xx = (double)(a.n1[j] + 4 * a.n2[j]);
  const double dn = (double)a.n;
  double v = xx - s * s / dn;                       // precomputed_snp_covariances.hpp:116-117
  if (a.miss_corr) {                                 // data_model.cpp:105-167, type A
    const double dc = a.miss_corr[3 * j], sv = a.miss_corr[3 * j + 1], sv2 = a.miss_corr[3 * j + 2];
    dot += dc;
    v += sv2 - sv * (s * 2.0 + sv) / dn;
    s += sv;
    xx += sv2;
  }
  a.dot[j] = dot;

  double tau;
  if (a.tau_mode == 0) tau = a.tau_shared;
  else if (a.tau_mode == 1) tau = a.tau_snp[j];
  else {                                             // prior.hpp:192-199 with a counter-based generator
    Philox g(a.seed, a.counter, (uint64_t)(a.lo + j));
    double val = -1.0;
    while (!(val > 0.0) || !isfinite(val)) val = a.nu_tau2 * a.s2_tau2 / (2.0 * g.gamma(0.5 * a.nu_tau2));
    tau = 1.0 / (a.alpha2 * val);
  }
  int pos = -1;
  const int64_t gj = a.lo + j;
  for (int l = 0; l < a.k; ++l)
    if (a.loci[l] == gj) pos = l;
  double mr, lmp;
  if (pos < 0) {                                     // sampler.cpp:109-115
    mr = a.sum_r / dn;
    lmp = a.lmp_add;
  } else {                                           // sampler.cpp:116-149: residual without SNP j
    const double b = a.beta_g[pos];
    if (a.tau_mode != 0) tau = a.tau_g[pos];
    dot += b * xx;                                   // x.(r + b x) = x.r + b x.x
    mr = (a.sum_r + b * s) / dn;
    lmp = a.lmp_rem;
  }
  const double rx = dot - s * mr;                    // sampler.cpp:188
  const double det = v + tau;
  const double exp_term = (rx * rx) / det;
  double p = exp_term / (2.0 * a.sigma2) - 0.5 * (log(det) - log(tau)) + lmp;
  p = exp(p);
  a.p_r[j] = isfinite(p) ? p / (1.0 + p) : 1.0;      // sampler.cpp:200-206
}

// ---------------------------------------------------------------------------------------
// epilogue of Sampler::sample's rao block (src/sampler.cpp:739-803) + flat initialisation
// ---------------------------------------------------------------------------------------
__global__ void k_adapt(const double* __restrict__ p_r, int64_t m, int upd_rao, double rz1, double rz2, int upd_prop,
                        double pz1, double pz2, double q_add_min, double q_rem_min, double* __restrict__ p_rao,
                        double* __restrict__ p_prop, double* __restrict__ q_add, double* __restrict__ q_rem)
{
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= m) return;
  const double pr = p_r[j];
  if (upd_rao) p_rao[j] = __dadd_rn(__dmul_rn(rz2, p_rao[j]), __ddiv_rn(pr, rz1));
  if (upd_prop) {
    const double pp = __dadd_rn(__dmul_rn(pz2, p_prop[j]), __ddiv_rn(pr, pz1));
    p_prop[j] = pp;
    q_add[j] = fmax(pp, q_add_min);
    q_rem[j] = fmax(1.0 - pp, q_rem_min);
  }
}

__global__ void k_fill_flat(int64_t m, double value, double q_add_min, double q_rem_min, double* __restrict__ p_prop,
                            double* __restrict__ q_add, double* __restrict__ q_rem, double* __restrict__ p_rao)
{
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= m) return;
  p_prop[j] = value;
  q_add[j] = fmax(value, q_add_min);
  q_rem[j] = fmax(1.0 - value, q_rem_min);
  p_rao[j] = 0.0;
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
static void choose_scan_geometry(Chain* c)
{
  const Store* s = c->store;
  const int64_t W = s->W;
  double best = -1.0;
  int best_nw = 1;
  for (int nw = 1; nw <= kMaxWarps; ++nw) {
    const int64_t chunks = (W + 64 * nw - 1) / (64 * nw);
    const int ctas = 16 / nw;  // 128 registers/thread => 512 threads per SM
    const double eff = (double)W / (double)(chunks * 64 * nw) * (double)(ctas * nw) / 16.0;
    if (eff > best + 1e-9 || (eff > best - 1e-9 && nw > best_nw)) { best = eff; best_nw = nw; }
  }
  if (const char* env = getenv("BMG_SCAN_WARPS")) {   // development override
    const int v = atoi(env);
    if (v >= 1 && v <= kMaxWarps) best_nw = v;
  }
  c->scan_warps = best_nw;
  c->scan_chunk_words = 64 * best_nw;
  c->scan_chunks = (int)((W + c->scan_chunk_words - 1) / c->scan_chunk_words);
}

static size_t scan_smem_bytes(const Chain* c)
{
  const Store* s = c->store;
  const int64_t row_words = c->scan_chunks == 1 ? s->Wp : c->scan_chunk_words;
  return (size_t)kStages * kTileSnps * row_words * 4 + 2 * kMaxWarps * kTileSnps * sizeof(double) + kStages * sizeof(uint64_t);
}

// the per-SNP proposal / Rao-Blackwell arrays over [0, mw) (see Chain)
static void alloc_weight_arrays(Chain* c)
{
  const int64_t ma = c->mw_alloc, mw = c->mw;
  c->p_r.alloc(ma); c->p_rao.alloc(ma); c->p_proposal.alloc(ma); c->q_add.alloc(ma); c->q_rem.alloc(ma);
  BMG_CUDA(cudaMemset(c->p_r.p, 0, ma * sizeof(double)));
  BMG_CUDA(cudaMemset(c->p_rao.p, 0, ma * sizeof(double)));
  BMG_CUDA(cudaMemset(c->p_proposal.p, 0, ma * sizeof(double)));
  BMG_CUDA(cudaMemset(c->q_add.p, 0, ma * sizeof(double)));
  BMG_CUDA(cudaMemset(c->q_rem.p, 0, ma * sizeof(double)));
  c->cdf_blocks = (mw + c->cdf_block - 1) / c->cdf_block;
  c->cdf_add.alloc(c->cdf_blocks); c->cdf_rem.alloc(c->cdf_blocks);
  c->zero_add.alloc(mw); c->zero_rem.alloc(mw);
  BMG_CUDA(cudaMemset(c->zero_add.p, 0, mw));
  BMG_CUDA(cudaMemset(c->zero_rem.p, 0, mw));
  c->q_add_io.release(); c->q_rem_io.release(); c->cdf_eff_add.release(); c->cdf_eff_rem.release();
}

// SNP-sharded chain: rank r of `world` holds the packed SNPs [r stride, min(m_g, (r+1) stride)); the weight arrays
// are re-created over all m_g SNPs (padded to world x stride for the all-gather) with the GLOBAL in-order permutation.
void chain_set_sharded(Chain* c, int world, int rank, int64_t stride, AllGatherFn fn, void* ctx)
{
  Store* s = c->store;
  BMG_REQUIRE(world >= 1 && rank >= 0 && rank < world && stride > 0 && fn != nullptr, "sharded chain: invalid communicator");
  BMG_REQUIRE((int64_t)world * stride >= s->m_g, "sharded chain: world x stride does not cover m_g");
  BMG_REQUIRE(s->lo == std::min(s->m_g, (int64_t)rank * stride) && s->hi == std::min(s->m_g, (int64_t)(rank + 1) * stride),
              "sharded chain: the store's SNP range is not this rank's shard");
  BMG_CUDA(cudaSetDevice(s->device));
  c->world = world; c->rank = rank; c->shard_stride = stride; c->gather = fn; c->gather_ctx = ctx;
  c->mw = s->m_g; c->w_off = s->lo; c->mw_alloc = (int64_t)world * stride;
  alloc_weight_arrays(c);
  build_inorder_permutation(c->mw, c->h_inorder_own);
  c->inorder_own.alloc(c->mw);
  bmg::copy_h2d_sync(c->inorder_own.p, c->h_inorder_own.data(), c->mw * sizeof(int32_t));
  BMG_CUDA(cudaDeviceSynchronize());
  // missing calls: the lockstep ranks draw the same imputed values, so every rank keeps them for ALL SNPs and needs the
  // whole data set's index (collective)
  c->mv_own.reset(build_global_missing(s, world, rank, stride, fn, ctx));
  if (c->mv_own->n_missing > 0) chain_bind_missing(c, c->mv_own->view());
}

// One of several chains over a SNP-sharded store (group.cu): the chain's per-SNP arrays cover all m_g SNPs with the
// GLOBAL in-order permutation, as for chain_set_sharded, but nothing is replicated -- the chain exists on this rank only.
void chain_set_group(Chain* c, Group* g)
{
  Store* s = c->store;
  BMG_REQUIRE(g != nullptr && c->group == nullptr && c->world == 1, "shard group: chain already attached");
  BMG_CUDA(cudaSetDevice(s->device));
  c->group = g;
  c->mw = s->m_g; c->w_off = s->lo; c->mw_alloc = (int64_t)group_world(g) * group_stride(g);
  alloc_weight_arrays(c);
  c->dot.alloc(c->mw_alloc);
  build_inorder_permutation(c->mw, c->h_inorder_own);
  c->inorder_own.alloc(c->mw);
  bmg::copy_h2d_sync(c->inorder_own.p, c->h_inorder_own.data(), c->mw * sizeof(int32_t));
  BMG_CUDA(cudaDeviceSynchronize());
  const GlobalMissing* gm = group_missing(g);   // built collectively at group creation
  if (gm->n_missing > 0) chain_bind_missing(c, gm->view());
}

// every rank contributes elems [rank stride, (rank+1) stride) of dev_buffer (world x stride elements) in place
void chain_allgather(Chain* c, void* dev_buffer, int elem_bytes)
{
  if (c->world <= 1) return;
  const int rc = c->gather(c->gather_ctx, dev_buffer, c->shard_stride, elem_bytes, (void*)c->stream);
  if (rc != 0) throw Error("sharded chain: the host's all-gather callback failed");
}

// the chain's imputed values follow the index of `v`; every cell starts at 0 (data_model.hpp:80-84)
void chain_bind_missing(Chain* c, const MissView& v)
{
  BMG_CUDA(cudaSetDevice(c->store->device));
  c->mv = v;
  c->miss_val.release();
  c->miss_corr.release();
  if (v.n_missing > 0) {
    c->miss_val.alloc((size_t)v.n_missing);
    BMG_CUDA(cudaMemset(c->miss_val.p, 0, (size_t)v.n_missing));
    c->miss_corr.alloc(3 * (size_t)v.m);
    BMG_CUDA(cudaMemset(c->miss_corr.p, 0, 3 * (size_t)v.m * sizeof(double)));
    BMG_CUDA(cudaDeviceSynchronize());
  }
  chain_overlay_invalidate(c, nullptr, 0);
}

Chain* chain_create(Store* s)
{
  BMG_REQUIRE(s->m_e >= 1, "bmg_chain_create: call bmg_store_set_phenotype first");
  BMG_CUDA(cudaSetDevice(s->device));
  std::unique_ptr<Chain> c(new Chain());
  c->store = s;
  if (const char* env = getenv("BMG_SCAN_VARIANT")) {
    const int v = atoi(env);
    if (v >= 0 && v <= 2) c->scan_variant = v;
  }
  BMG_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  if (const char* env = getenv("BMG_COLSTATS_SERVER")) c->server_enabled = atoi(env) != 0;
  choose_scan_geometry(c.get());
  const int64_t n = s->n, m = s->m;
  const int64_t n_pad = 16 * (int64_t)c->scan_chunks * c->scan_chunk_words;
  c->y.alloc(n);
  BMG_CUDA(cudaMemcpy(c->y.p, s->y.p, n * sizeof(double), cudaMemcpyDeviceToDevice));
  c->yhat_e.alloc(n); c->yhat_g.alloc(n); c->r.alloc(n); c->r_scaled.alloc(n_pad);
  c->red_partial.alloc(1024 * 16); c->red_out.alloc(16); c->h_red.alloc(16);
  {
    MissView v;
    v.base = s->lo; v.m = s->m; v.n_missing = s->n_missing; v.off = s->miss_off.p; v.idx = s->miss_idx.p;
    v.n1 = s->n1.p; v.n2 = s->n2.p; v.nmiss = s->nmiss.p; v.h_off = s->h_miss_off.data();
    chain_bind_missing(c.get(), v);
  }
  c->dot_partial.alloc((size_t)c->scan_chunks * m);
  c->dot.alloc(m);
  c->loci_dev.alloc(2048); c->beta_dev.alloc(2048 + 64); c->taug_dev.alloc(2048);
  c->h_stage.alloc(8192); c->h_stage_i.alloc(4096);
  c->sample_out.alloc(4); c->h_sample.alloc(4);
  c->mw = m; c->w_off = 0; c->mw_alloc = m;
  alloc_weight_arrays(c.get());
  const size_t smem = scan_smem_bytes(c.get());
  BMG_REQUIRE(smem <= 227 * 1024, "scan tile does not fit in shared memory");
  BMG_CUDA(cudaFuncSetAttribute(k_scan_dots_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // CTAs per chunk: fill every SM with as many resident CTAs as registers/smem allow
  int per_sm = 1;
  BMG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_scan_dots_tma, 32 * c->scan_warps, smem));
  if (per_sm < 1) per_sm = 1;
  const int64_t tiles = (m + kTileSnps - 1) / kTileSnps;
  int64_t slices = ((int64_t)per_sm * s->sm_count) / c->scan_chunks;
  if (slices < 1) slices = 1;
  if (slices > tiles) slices = tiles;
  c->scan_ctas_per_chunk = (int)slices;
  BMG_CUDA(cudaDeviceSynchronize());
  return c.release();
}

void chain_destroy(Chain* c)
{
  if (!c) return;
  cudaSetDevice(c->store->device);
  chain_server_stop(c);
  chain_forget_server(c);
  if (c->server_stream) cudaStreamDestroy(c->server_stream);
  if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
  for (cudaEvent_t e : c->scan_ev) cudaEventDestroy(e);
  delete c;
}

void chain_set_missing(Chain* c, int64_t snp, const int8_t* vals, int64_t count)
{
  BMG_REQUIRE(c->mv.covers(snp), "bmg_chain_set_missing: SNP outside the chain's missing-call index");
  const int64_t cnt = c->mv.n_missing > 0 ? c->mv.count(snp) : 0;
  BMG_REQUIRE(cnt == count, "bmg_chain_set_missing: count does not match the number of missing cells of the SNP");
  if (cnt == 0) return;
  chain_set_missing_many(c, &snp, 1, &vals);   // overlay.cu: values to the device + the SNP's patched column, one launch
}

// DataModel::sample_missing (data_model.cpp:78-90) draws every cell again before a scan: one upload for the shard
void chain_set_missing_all(Chain* c, const int8_t* vals, int64_t count)
{
  Store* s = c->store;
  BMG_REQUIRE(count == c->mv.n_missing, "bmg_chain_set_missing_all: count does not match the number of missing cells of the store");
  if (count == 0) return;
  unsigned bad = 0;
  for (int64_t q = 0; q < count; ++q) bad |= (unsigned)((uint8_t)vals[q] > 2);
  BMG_REQUIRE(bad == 0, "bmg_chain_set_missing_all: values must be 0, 1 or 2");
  BMG_CUDA(cudaSetDevice(s->device));
  bmg::copy_h2d(c->miss_val.p, vals, (size_t)count, c->stream);
  BMG_CUDA(cudaStreamSynchronize(c->stream));
  chain_overlay_invalidate(c, nullptr, 0);   // the caller may have changed any SNP
}

// DataModel::sample_missing on the device (throughput mode; the parity mode draws on the host from the chain's own
// stream): one warp per SNP with missing calls, a counter-based uniform per cell, class from the cumulative counts of
// 0/1/2 among the SNP's observed cells (data.cpp:357-372).  SNPs of the model (sorted list) keep their values.
__global__ void k_impute_from_prior(const int64_t* __restrict__ off, const int32_t* __restrict__ n1,
                                    const int32_t* __restrict__ n2, const int32_t* __restrict__ nmiss, int64_t n, int64_t m,
                                    int64_t snp_lo, const int64_t* __restrict__ model_sorted, int k, uint64_t seed,
                                    uint64_t counter, int8_t* __restrict__ val)
{
  const int64_t j = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (j >= m) return;
  const int64_t lo = off[j], hi = off[j + 1];
  if (hi == lo) return;
  int a = 0, b = k;
  while (a < b) {
    const int mid = (a + b) >> 1;
    const int64_t v = model_sorted[mid];
    if (v == j + snp_lo) return;   // in the model: left to the Gibbs step
    if (v < j + snp_lo) a = mid + 1; else b = mid;
  }
  const double c0 = (double)(n - nmiss[j] - n1[j] - n2[j]), c1 = c0 + (double)n1[j], c2 = c1 + (double)n2[j];
  for (int64_t q = lo + lane; q < hi; q += 32) {
    Philox g(seed, counter, (uint64_t)q);
    const double r = g.u01() * c2;
    val[q] = (int8_t)(r < c0 ? 0 : (r < c1 ? 1 : 2));
  }
}

void chain_impute_from_prior(Chain* c, const int64_t* loci, int k, uint64_t seed, uint64_t counter)
{
  Store* s = c->store;
  const MissView& mv = c->mv;
  if (mv.n_missing == 0) return;
  BMG_REQUIRE(k >= 0 && k <= 2048, "bmg_chain_impute_from_prior: model size must be <= 2048");
  BMG_CUDA(cudaSetDevice(s->device));
  cudaStream_t st = c->stream;
  BMG_CUDA(cudaStreamSynchronize(st));   // the pinned staging buffer may still be in flight
  for (int l = 0; l < k; ++l) c->h_stage_i.p[l] = loci[l];
  std::sort(c->h_stage_i.p, c->h_stage_i.p + k);
  if (k) bmg::copy_h2d(c->loci_dev.p, c->h_stage_i.p, k * sizeof(int64_t), st);
  k_impute_from_prior<<<(unsigned)((mv.m * 32 + 127) / 128), 128, 0, st>>>(mv.off, mv.n1, mv.n2, mv.nmiss, s->n, mv.m, mv.base,
                                                                          c->loci_dev.p, k, seed, counter, c->miss_val.p);
  count_launch();
  BMG_CUDA(cudaGetLastError());
  BMG_CUDA(cudaStreamSynchronize(st));   // loci_dev / the staging buffer are reused by the scan that follows
  chain_overlay_invalidate(c, loci, k);  // the model's SNPs kept their values
}

// ---------------------------------------------------------------------------------------
// a few cells of a few columns, the chain's imputed values applied: what the reference reads as
// current_model->x(i_miss, col) in the missing-genotype Gibbs step (src/sampler.cpp:304-449)
// ---------------------------------------------------------------------------------------
struct CellArgs {
  const uint32_t* const* cols;   // k packed columns
  const int64_t* snp_local;      // local index of each column's SNP, -1 for a peer's
  const int32_t* rows;           // q individuals
  int k;
  int64_t q;
  const int64_t* off;            // store's missing index (CSR)
  const int32_t* idx;
  const int8_t* val;             // chain's imputed values
  int8_t* out;                   // k x q
};

__global__ void k_gather_cells(const CellArgs a)
{
  const int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (gid >= (int64_t)a.k * a.q) return;
  const int l = (int)(gid / a.q);
  const int32_t i = a.rows[gid - (int64_t)l * a.q];
  int v = (int)((a.cols[l][i >> 4] >> (2 * (i & 15))) & 3u);
  const int64_t j = a.snp_local[l];
  if (j >= 0) {
    int64_t lo = a.off[j], hi = a.off[j + 1];
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      const int32_t w = a.idx[mid];
      if (w == i) { v = a.val[mid] | 4; break; }   // bit 2: the cell is a missing call, its value is imputed
      if (w < i) lo = mid + 1; else hi = mid;
    }
  }
  a.out[gid] = (int8_t)v;
}

void chain_get_cells(Chain* c, const int64_t* loci, int k, const int32_t* rows, int64_t q, int8_t* out)
{
  Store* s = c->store;
  BMG_REQUIRE(k >= 0 && k <= 2048 && q >= 0, "bmg_chain_get_cells: bad sizes");
  if (k == 0 || q == 0) return;
  for (int64_t t = 0; t < q; ++t) BMG_REQUIRE(rows[t] >= 0 && rows[t] < s->n, "bmg_chain_get_cells: individual out of range");
  BMG_CUDA(cudaSetDevice(s->device));
  cudaStream_t st = c->stream;
  const size_t cells = (size_t)k * (size_t)q;
  if (c->gc_meta.n < (size_t)2 * k || c->gc_rows.n < (size_t)q || c->gc_out.n < cells) {
    chain_server_stop(c);   // allocations synchronise the device: a running server would stall them
    BMG_CUDA(cudaStreamSynchronize(st));
    if (c->gc_meta.n < (size_t)2 * k) { c->gc_meta.alloc(4096); c->gc_h_meta.alloc(4096); }
    if (c->gc_rows.n < (size_t)q) { c->gc_rows.alloc(2 * q + 1024); c->gc_h_rows.alloc(2 * q + 1024); }
    if (c->gc_out.n < cells) { c->gc_out.alloc(2 * cells + 4096); c->gc_h_out.alloc(2 * cells + 4096); }
  }
  for (int l = 0; l < k; ++l) {
    c->gc_h_meta.p[l] = (int64_t)(uintptr_t)s->column_ptr(loci[l]);
    const bool indexed = c->mv.n_missing > 0 && c->mv.covers(loci[l]);
    c->gc_h_meta.p[k + l] = indexed ? loci[l] - c->mv.base : -1;
  }
  memcpy(c->gc_h_rows.p, rows, (size_t)q * sizeof(int32_t));
  bmg::copy_h2d(c->gc_meta.p, c->gc_h_meta.p, (size_t)2 * k * sizeof(int64_t), st);
  bmg::copy_h2d(c->gc_rows.p, c->gc_h_rows.p, (size_t)q * sizeof(int32_t), st);
  CellArgs a;
  a.cols = reinterpret_cast<const uint32_t* const*>(c->gc_meta.p);
  a.snp_local = c->gc_meta.p + k;
  a.rows = c->gc_rows.p; a.k = k; a.q = q;
  a.off = c->mv.off; a.idx = c->mv.idx; a.val = c->miss_val.p; a.out = c->gc_out.p;
  k_gather_cells<<<(unsigned)((cells + 127) / 128), 128, 0, st>>>(a);
  count_launch();
  BMG_CUDA(cudaGetLastError());
  bmg::copy_d2h(c->gc_h_out.p, c->gc_out.p, cells, st);
  BMG_CUDA(cudaStreamSynchronize(st));
  memcpy(out, c->gc_h_out.p, cells);
}

// uploads loci / beta / tau of the current model into the chain's device scratch
static void upload_model(Chain* c, const int64_t* loci, const double* beta_g, const double* tau_g, int k)
{
  BMG_REQUIRE(k >= 0 && k <= 2048, "model size must be <= 2048");
  if (k == 0) return;
  // staged through pinned memory so the copies are truly asynchronous
  BMG_CUDA(cudaStreamSynchronize(c->stream));
  for (int l = 0; l < k; ++l) {
    c->h_stage_i.p[l] = loci[l];
    c->h_stage.p[l] = beta_g ? beta_g[l] : 0.0;
    c->h_stage.p[2048 + l] = tau_g ? tau_g[l] : 0.0;
  }
  bmg::copy_h2d(c->loci_dev.p, c->h_stage_i.p, k * sizeof(int64_t), c->stream);
  bmg::copy_h2d(c->beta_dev.p, c->h_stage.p, k * sizeof(double), c->stream);
  bmg::copy_h2d(c->taug_dev.p, c->h_stage.p + 2048, k * sizeof(double), c->stream);
}

void chain_residual(Chain* c, const int64_t* loci, const double* beta_e, const double* beta_g, int k, double* stats9,
                    const int32_t* term_types)
{
  chain_server_stop(c);   // precedes scans and latent sweeps (which may allocate and want all SMs)
  Store* s = c->store;
  BMG_CUDA(cudaSetDevice(s->device));
  BMG_REQUIRE(k >= 0 && k <= 2048, "bmg_chain_residual: model size must be <= 2048");
  cudaStream_t st = c->stream;
  upload_model(c, loci, beta_g, nullptr, k);
  // beta_e and column pointers
  BMG_CUDA(cudaStreamSynchronize(st));
  for (int j = 0; j < s->m_e; ++j) c->h_stage.p[4096 + j] = beta_e[j];
  bmg::copy_h2d(c->beta_dev.p + 2048, c->h_stage.p + 4096, s->m_e * sizeof(double), st);
  static_assert(sizeof(const uint32_t*) == sizeof(int64_t), "pointer size");
  {
    // columns as the chain sees them: a SNP with missing calls is read from its patched copy (overlay.cu)
    std::vector<const uint32_t*> cols(k);
    chain_overlay_columns(c, loci, k, cols.data(), nullptr, term_types);
    for (int l = 0; l < k; ++l) c->h_stage_i.p[2048 + l] = (int64_t)(uintptr_t)cols[l];
  }
  if (c->cs_idx.n < 4096) c->cs_idx.alloc(4096);
  if (k) bmg::copy_h2d(c->cs_idx.p, c->h_stage_i.p + 2048, k * sizeof(int64_t), st);
  YhatArgs ya;
  ya.cols = reinterpret_cast<const uint32_t* const*>(c->cs_idx.p);
  ya.beta_g = c->beta_dev.p; ya.beta_e = c->beta_dev.p + 2048; ya.e = s->e.p;
  ya.k = k; ya.m_e = s->m_e; ya.n = s->n; ya.W = s->W; ya.yhat_e = c->yhat_e.p; ya.yhat_g = c->yhat_g.p;
  k_yhat<<<(unsigned)((s->W + 127) / 128), 128, 0, st>>>(ya);
  count_launch();
  const int64_t n_pad = (int64_t)c->r_scaled.n;
  int blocks = (int)std::min<int64_t>(1024, (n_pad + 255) / 256);
  k_residual<<<blocks, 256, 0, st>>>(c->y.p, c->yhat_e.p, c->yhat_g.p, s->n, n_pad, c->r.p, c->r_scaled.p, c->red_partial.p);
  count_launch();
  k_reduce_final<<<1, 32, 0, st>>>(c->red_partial.p, blocks, kRed, c->red_out.p);
  count_launch();
  BMG_CUDA(cudaGetLastError());
  bmg::copy_d2h(c->h_red.p, c->red_out.p, kRed * sizeof(double), st);
  BMG_CUDA(cudaStreamSynchronize(st));
  c->sum_r = c->h_red.p[0];
  c->residual_valid = true;
  c->imma_q_valid = false;
  if (stats9) for (int q = 0; q < kRed; ++q) stats9[q] = c->h_red.p[q];
}

// optional CUDA-event pair around the scan's reduction kernel (bmg_chain_scan_kernel_time; bench.py roofline)
void scan_timer_begin(Chain* c, cudaStream_t st)
{
  if (!c->time_scan) return;
  if (c->scan_ev_used >= 4096) {   // fold finished pairs into the running total
    BMG_CUDA(cudaStreamSynchronize(st));
    for (size_t i = 0; i < c->scan_ev_used; ++i) {
      float ms = 0.f;
      BMG_CUDA(cudaEventElapsedTime(&ms, c->scan_ev[2 * i], c->scan_ev[2 * i + 1]));
      c->scan_ms_done += ms;
    }
    c->scan_launches_done += (int64_t)c->scan_ev_used;
    c->scan_ev_used = 0;
  }
  while (c->scan_ev.size() < 2 * (c->scan_ev_used + 1)) {
    cudaEvent_t e;
    BMG_CUDA(cudaEventCreate(&e));
    c->scan_ev.push_back(e);
  }
  BMG_CUDA(cudaEventRecord(c->scan_ev[2 * c->scan_ev_used], st));
}
void scan_timer_end(Chain* c, cudaStream_t st)
{
  if (!c->time_scan) return;
  BMG_CUDA(cudaEventRecord(c->scan_ev[2 * c->scan_ev_used + 1], st));
  ++c->scan_ev_used;
}

void chain_scan_dots(Chain* c)
{
  Store* s = c->store;
  BMG_REQUIRE(c->residual_valid, "scan: call bmg_chain_residual first");
  BMG_CUDA(cudaSetDevice(s->device));
  ScanArgs a;
  a.codes = s->codes.p; a.Wp = s->Wp; a.m = s->m; a.r_scaled = c->r_scaled.p;
  a.chunk_words = (int)c->scan_chunk_words; a.n_chunks = c->scan_chunks;
  a.tiles = (s->m + kTileSnps - 1) / kTileSnps; a.slices = c->scan_ctas_per_chunk; a.out = c->dot_partial.p;
  const unsigned grid = (unsigned)(c->scan_chunks * c->scan_ctas_per_chunk);
  if (c->scan_variant == 2) imma_quantize(c);   // residual -> fixed-point limbs, outside the timed pair
  scan_timer_begin(c, c->stream);
  if (c->scan_variant == 2) {
    imma_launch(c);
  } else {
    if (c->scan_variant == 1)
      k_scan_dots_tma<<<grid, 32 * c->scan_warps, scan_smem_bytes(c), c->stream>>>(a);
    else
      k_scan_dots_ldg<<<grid, 32 * c->scan_warps, 0, c->stream>>>(a);
    count_launch();
    c->last_partial = c->dot_partial.p;
    c->last_chunks = c->scan_chunks;
  }
  scan_timer_end(c, c->stream);
  BMG_CUDA(cudaGetLastError());
}

void chain_scan(Chain* c, const int64_t* loci, const double* beta_g, const double* tau_g, int k,
                const bmg_scan_params* prm, double* p_r_host)
{
  Store* s = c->store;
  BMG_REQUIRE(prm != nullptr, "bmg_chain_scan: params required");
  BMG_REQUIRE(prm->tau_mode >= 0 && prm->tau_mode <= 2, "bmg_chain_scan: tau_mode must be 0, 1 or 2");
  BMG_REQUIRE(prm->sigma2 > 0, "bmg_chain_scan: sigma2 must be positive");
  BMG_CUDA(cudaSetDevice(s->device));
  chain_server_stop(c);   // the scan's CTAs want every SM's registers; the server restarts with the next request
  cudaStream_t st = c->stream;
  upload_model(c, loci, beta_g, tau_g, k);
  if (prm->tau_mode == 1) {
    BMG_REQUIRE(prm->tau_host != nullptr, "bmg_chain_scan: tau_host required for tau_mode 1");
    const size_t need = c->group ? (size_t)c->mw : (size_t)s->m;
    if (c->tau_dev.n < need) { BMG_CUDA(cudaStreamSynchronize(st)); c->tau_dev.alloc(need); }
    if (!c->group) bmg::copy_h2d(c->tau_dev.p, prm->tau_host + c->w_off, s->m * sizeof(double), st);   // tau_host covers [0, mw)
  }
  if (c->group) {
    // several chains over the sharded store: every rank of the group scans its shard for this chain's residual and stores
    // the dot products into this GPU's memory (group.cu); the per-SNP algebra then runs here over all m_g SNPs
    BMG_REQUIRE(c->scan_variant == 2, "shard group: the tensor-core scan only");
    if (prm->tau_mode == 1) bmg::copy_h2d(c->tau_dev.p, prm->tau_host, c->mw * sizeof(double), st);
    if (c->mv.n_missing > 0) {   // the chain's imputed cells of ALL SNPs against its residual, on its own GPU
      k_miss_corr<<<(unsigned)((c->mv.m * 32 + 127) / 128), 128, 0, st>>>(c->mv.off, c->mv.idx, c->miss_val.p, c->r.p, c->mv.m, c->miss_corr.p);
      count_launch();
    }
    const double* dots = group_scan_round(c->group, c);
    FinalizeArgs f;
    f.dot_partial = dots; f.n_chunks = 1; f.m = c->mw; f.lo = 0; f.n = s->n;
    f.n1 = group_n1(c->group); f.n2 = group_n2(c->group); f.miss_corr = c->mv.n_missing > 0 ? c->miss_corr.p : nullptr;
    f.loci = c->loci_dev.p; f.beta_g = c->beta_dev.p; f.tau_g = c->taug_dev.p; f.k = k;
    f.sum_r = c->sum_r; f.sigma2 = prm->sigma2; f.lmp_add = prm->lmp_add; f.lmp_rem = prm->lmp_rem;
    f.tau_mode = prm->tau_mode; f.tau_shared = prm->tau_shared; f.tau_snp = c->tau_dev.p;
    f.seed = prm->tau_seed; f.counter = prm->tau_counter; f.nu_tau2 = prm->nu_tau2; f.s2_tau2 = prm->s2_tau2; f.alpha2 = prm->alpha2;
    f.dot = c->dot.p; f.p_r = c->p_r.p;
    k_scan_finalize<<<(unsigned)((c->mw + 255) / 256), 256, 0, st>>>(f);
    count_launch();
    BMG_CUDA(cudaGetLastError());
    if (p_r_host) {
      bmg::copy_d2h(p_r_host, c->p_r.p, c->mw * sizeof(double), st);
      BMG_CUDA(cudaStreamSynchronize(st));
    }
    return;
  }
  chain_scan_dots(c);
  if (c->mv.n_missing > 0) {   // over the SNPs of the chain's index (all of them when the store is sharded: cheap, and the same on every rank)
    k_miss_corr<<<(unsigned)((c->mv.m * 32 + 127) / 128), 128, 0, st>>>(c->mv.off, c->mv.idx, c->miss_val.p, c->r.p, c->mv.m, c->miss_corr.p);
    count_launch();
  }
  FinalizeArgs f;
  f.dot_partial = c->last_partial; f.n_chunks = c->last_chunks; f.m = s->m; f.lo = s->lo; f.n = s->n;
  f.n1 = s->n1.p; f.n2 = s->n2.p; f.miss_corr = c->mv.n_missing > 0 ? c->miss_corr.p + 3 * (s->lo - c->mv.base) : nullptr;
  f.loci = c->loci_dev.p; f.beta_g = c->beta_dev.p; f.tau_g = c->taug_dev.p; f.k = k;
  f.sum_r = c->sum_r; f.sigma2 = prm->sigma2; f.lmp_add = prm->lmp_add; f.lmp_rem = prm->lmp_rem;
  f.tau_mode = prm->tau_mode; f.tau_shared = prm->tau_shared; f.tau_snp = c->tau_dev.p;
  f.seed = prm->tau_seed; f.counter = prm->tau_counter; f.nu_tau2 = prm->nu_tau2; f.s2_tau2 = prm->s2_tau2; f.alpha2 = prm->alpha2;
  f.dot = c->dot.p; f.p_r = c->p_r.p + c->w_off;
  k_scan_finalize<<<(unsigned)((s->m + 255) / 256), 256, 0, st>>>(f);
  count_launch();
  BMG_CUDA(cudaGetLastError());
  chain_allgather(c, c->p_r.p, (int)sizeof(double));   // no-op unless SNP-sharded
  if (p_r_host) {
    bmg::copy_d2h(p_r_host, c->p_r.p, c->mw * sizeof(double), st);
    BMG_CUDA(cudaStreamSynchronize(st));
  }
}

// ---------------------------------------------------------------------------------------
// the scan with several effect types, or one type other than A (src/sampler.cpp:90-259)
//
// Both statistics every type needs come from the packed store: S1 = sum_{x=1} r (the heterozygote-indicator pass of
// the tensor-core kernel) and the additive dot S1 + 2 S2 (the ordinary pass); A = S1 + 2 S2, H = S1, D = S1 + S2,
// R = S2, and all moments follow from the genotype counts.
// ---------------------------------------------------------------------------------------
// out[4j..] = { sum_{val=1} r, sum_{val=2} r, #val=1, #val=2 } over the imputed cells of SNP j
__global__ void k_miss_corr4(const int64_t* __restrict__ off, const int32_t* __restrict__ idx, const int8_t* __restrict__ val,
                             const double* __restrict__ r, int64_t m, double* __restrict__ out)
{
  const int64_t j = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (j >= m) return;
  const int64_t lo = off[j], hi = off[j + 1];
  if (hi == lo) return;
  double d1 = 0.0, d2 = 0.0, c1 = 0.0, c2 = 0.0;
  for (int64_t q = lo + lane; q < hi; q += 32) {
    const int v = val[q];
    const double rv = r[idx[q]];
    if (v == 1) { d1 += rv; c1 += 1.0; }
    else if (v == 2) { d2 += rv; c2 += 1.0; }
  }
  d1 = warp_sum(d1); d2 = warp_sum(d2); c1 = warp_sum(c1); c2 = warp_sum(c2);
  if (lane == 0) { out[4 * j] = d1; out[4 * j + 1] = d2; out[4 * j + 2] = c1; out[4 * j + 3] = c2; }
}

struct TypedArgs {
  const double* part_a; const double* part_h; int n_chunks;
  int64_t m, lo, n;
  const int32_t* n1; const int32_t* n2; const int32_t* nmiss;
  const double* corr4;          // may be null
  // model: k SNPs, each with its effect type and one or two terms (AH: additive then heterozygous)
  const int64_t* loci; const int32_t* mtype; const double* mbeta; const double* mtau; int k;   // mbeta/mtau: [k][2]
  double sum_r, sigma2;
  int n_types; int types[5];
  int n_terms; int term_rank[4];  // -1: term not allowed
  int allow_ah, ref_offsets, tau_mode;
  double lmp_add[5], lmp_rem[25], tau_shared[4];
  const double* tau_snp;        // [m][n_terms] when tau_mode == 1
  double* p_r; double* p_r_types;
};

// value of a term type on genotype 1 and genotype 2
__device__ __forceinline__ void type_values(int t, double& a1, double& a2)
{
  a1 = (t == 3) ? 0.0 : 1.0;                    // A, H, D count a heterozygote
  a2 = (t == 0) ? 2.0 : ((t == 1) ? 0.0 : 1.0); // A counts 2, D and R count 1
}

__global__ void __launch_bounds__(128) k_scan_finalize_types(const __grid_constant__ TypedArgs a)
{
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= a.m) return;
  double dot_a = 0.0, dot_h = 0.0;
  for (int c = 0; c < a.n_chunks; ++c) { dot_a += a.part_a[(int64_t)c * a.m + j]; dot_h += a.part_h[(int64_t)c * a.m + j]; }
  const double dn = (double)a.n;
  const double n1 = (double)a.n1[j], n2 = (double)a.n2[j];
  double s1 = dot_h, s2 = 0.5 * (dot_a - dot_h);   // sums of r over the observed heterozygotes / minor homozygotes
  double c1 = 0.0, c2 = 0.0;
  const bool has_missing = a.corr4 != nullptr && a.nmiss[j] > 0;
  if (has_missing) { s1 += a.corr4[4 * j]; s2 += a.corr4[4 * j + 1]; c1 = a.corr4[4 * j + 2]; c2 = a.corr4[4 * j + 3]; }
  const double N1 = n1 + c1, N2 = n2 + c2;

  // the moment cache in the reference's layout (precomputed_snp_covariances.hpp:95-131), then update_prexx_cov
  // (data_model.cpp:105-167) for the imputed values
  double pre[9];
  const int offset = 2 * a.n_terms + a.allow_ah;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int rk = a.term_rank[t];
    if (rk < 0) continue;
    double a1, a2;
    type_values(t, a1, a2);
    const double s0 = a1 * n1 + a2 * n2, ss0 = a1 * a1 * n1 + a2 * a2 * n2;
    pre[2 * rk] = s0;
    pre[2 * rk + 1] = ss0 - s0 * s0 / dn;
  }
  if (a.allow_ah) pre[offset - 1] = n1 - pre[0] * pre[2] / dn;   // terms A and H are the first two
  if (has_missing) {
    const double sa_sh = a.allow_ah ? pre[0] * pre[2] : 0.0;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int rk = a.term_rank[t];
      if (rk < 0) continue;
      double a1, a2;
      type_values(t, a1, a2);
      const double sv = a1 * c1 + a2 * c2, sv2 = a1 * a1 * c1 + a2 * a2 * c2;
      const double old = pre[2 * rk];
      pre[2 * rk] += sv;
      pre[2 * rk + 1] += sv2 - sv * (old * 2.0 + sv) / dn;
    }
    if (a.allow_ah) {
      pre[offset - 1] += (sa_sh - pre[0] * pre[2]) / dn;
      pre[offset - 1] += c1;
    }
  }

  double tau[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int rk = a.term_rank[t];
    if (rk < 0) continue;
    tau[t] = a.tau_mode == 0 ? a.tau_shared[t] : a.tau_snp[j * a.n_terms + rk];
  }
  int pos = -1;
  const int64_t gj = a.lo + j;
  for (int l = 0; l < a.k; ++l)
    if (a.loci[l] == gj) pos = l;
  double mr = a.sum_r / dn;
  const double* lmp = a.lmp_add;
  // residual without SNP j (sampler.cpp:116-149): r + b1 x^(U1) + b2 x^(H); its dot with x^(T) and its mean follow from counts
  double add1 = 0.0, add2 = 0.0;   // what r_om adds to sum_{x=1} and to sum_{x=2}
  if (pos >= 0) {
    const int mt = a.mtype[pos];
    const int u1 = mt == 4 ? 0 : mt;
    double b1 = a.mbeta[2 * pos], b2 = mt == 4 ? a.mbeta[2 * pos + 1] : 0.0;
    if (a.tau_mode != 0) {
      tau[u1] = a.mtau[2 * pos];
      if (mt == 4) tau[1] = a.mtau[2 * pos + 1];
    }
    double u_a1, u_a2;
    type_values(u1, u_a1, u_a2);
    add1 = (b1 * u_a1 + b2) * N1;            // the H term is 1 on heterozygotes only
    add2 = (b1 * u_a2) * N2;
    mr = (a.sum_r + add1 + add2) / dn;
    lmp = a.lmp_rem + 5 * mt;
  }
  const double S1 = s1 + add1, S2 = s2 + add2;

  double p[5];
  double max_types = -INFINITY;
  for (int ti = 0; ti < a.n_types; ++ti) {
    const int t = a.types[ti];
    double det, exp_term, sum_log_q;
    if (t == 4) {
      const double sa = pre[0], va = pre[1] + tau[0], sh = pre[2], vh = pre[3] + tau[1], vah = pre[offset - 1];
      sum_log_q = -log(tau[0]) - log(tau[1]);
      const double rxa = (S1 + 2.0 * S2) - sa * mr, rxh = S1 - sh * mr;
      det = va * vh - vah * vah;
      exp_term = (rxa * rxa * vh - 2.0 * rxa * rxh * vah + rxh * rxh * va) / det;
    } else {
      const int o = a.ref_offsets ? a.term_rank[t] : 2 * a.term_rank[t];   // see TypedArgs / include/bmagwa_b200.h
      double a1, a2;
      type_values(t, a1, a2);
      const double sT = pre[o];
      det = pre[o + 1] + tau[t];
      sum_log_q = -log(tau[t]);
      const double rx = (a1 * S1 + a2 * S2) - sT * mr;
      exp_term = (rx * rx) / det;
    }
    p[ti] = exp_term / (2.0 * a.sigma2) - 0.5 * (log(det) + sum_log_q) + lmp[t];
    if (p[ti] > max_types) max_types = p[ti];   // a NaN never wins this comparison (nor does it in the reference)
  }
  double* prt = a.p_r_types ? a.p_r_types + j * a.n_types : nullptr;
  if (a.n_types == 1) {                              // sampler.cpp:199-206
    const double e = exp(p[0]);
    a.p_r[j] = isfinite(e) ? e / (1.0 + e) : 1.0;
    return;
  }
  if (!isfinite(max_types)) {                        // sampler.cpp:210-237
    if (max_types > 0) {
      double sum = 0.0;
      for (int ti = 0; ti < a.n_types; ++ti) { prt[ti] = (!isfinite(p[ti]) && p[ti] > 0) ? 1.0 : 0.0; sum += prt[ti]; }
      for (int ti = 0; ti < a.n_types; ++ti) prt[ti] /= sum;
      a.p_r[j] = 1.0;
    } else {
      for (int ti = 0; ti < a.n_types; ++ti) prt[ti] = 1.0 / a.n_types;
      a.p_r[j] = 0.0;
    }
    return;
  }
  double sum_types = 0.0;                            // sampler.cpp:240-257
  for (int ti = 0; ti < a.n_types; ++ti) { p[ti] = exp(p[ti] - max_types); sum_types += p[ti]; }
  for (int ti = 0; ti < a.n_types; ++ti) prt[ti] = p[ti] / sum_types;
  sum_types = exp(log(sum_types) + max_types);
  a.p_r[j] = isfinite(sum_types) ? sum_types / (1.0 + sum_types) : 1.0;
}

void chain_scan_types(Chain* c, const int64_t* loci, const int32_t* loci_type, const double* beta2, const double* tau2, int k,
                      const bmg_scan_types_params* prm, double* p_r_host, double* p_r_types_host)
{
  Store* s = c->store;
  BMG_REQUIRE(prm != nullptr, "bmg_chain_scan_types: params required");
  BMG_REQUIRE(c->residual_valid, "bmg_chain_scan_types: call bmg_chain_residual[_types] first");
  BMG_REQUIRE(c->scan_variant == 2, "bmg_chain_scan_types: needs the tensor-core scan (variant 2)");
  BMG_REQUIRE(c->world == 1 && c->group == nullptr, "bmg_chain_scan_types: not available on a SNP-sharded chain");
  BMG_REQUIRE(prm->n_types >= 1 && prm->n_types <= 5, "bmg_chain_scan_types: n_types must be 1..5");
  BMG_REQUIRE(prm->tau_mode == 0 || prm->tau_mode == 1, "bmg_chain_scan_types: tau_mode must be 0 (shared) or 1 (host draws)");
  BMG_REQUIRE(prm->sigma2 > 0, "bmg_chain_scan_types: sigma2 must be positive");
  BMG_REQUIRE(k >= 0 && k <= 1024, "bmg_chain_scan_types: model size must be <= 1024");
  TypedArgs a;
  memset(&a, 0, sizeof(a));
  bool allow_type[5] = {false, false, false, false, false}, allow_term[4] = {false, false, false, false};
  for (int i = 0; i < prm->n_types; ++i) {
    const int t = prm->types[i];
    BMG_REQUIRE(t >= 0 && t <= 4 && !allow_type[t], "bmg_chain_scan_types: types must be distinct codes 0..4");
    BMG_REQUIRE(i == 0 || t > prm->types[i - 1], "bmg_chain_scan_types: types must be in increasing order (the reference sorts model.types)");
    allow_type[t] = true;
    if (t == 4) allow_term[0] = allow_term[1] = true; else allow_term[t] = true;
    a.types[i] = t;
  }
  a.n_types = prm->n_types;
  for (int t = 0; t < 4; ++t) a.term_rank[t] = allow_term[t] ? a.n_terms++ : -1;
  a.allow_ah = allow_type[4] ? 1 : 0;
  a.ref_offsets = prm->reference_offsets ? 1 : 0;
  for (int l = 0; l < k; ++l) BMG_REQUIRE(loci_type[l] >= 0 && loci_type[l] <= 4 && allow_type[loci_type[l]], "bmg_chain_scan_types: a model SNP has a type that is not configured");
  BMG_CUDA(cudaSetDevice(s->device));
  chain_server_stop(c);
  cudaStream_t st = c->stream;
  // model arrays
  if (c->tm_d.n == 0) { c->tm_d.alloc(4 * 1024); c->tm_i.alloc(1024); c->tm_h.alloc(4 * 1024 + 1024); }
  BMG_CUDA(cudaStreamSynchronize(st));
  for (int l = 0; l < k; ++l) {
    c->h_stage_i.p[l] = loci[l];
    c->tm_h.p[2 * l] = beta2[2 * l]; c->tm_h.p[2 * l + 1] = beta2[2 * l + 1];
    c->tm_h.p[2048 + 2 * l] = tau2[2 * l]; c->tm_h.p[2048 + 2 * l + 1] = tau2[2 * l + 1];
    reinterpret_cast<int32_t*>(c->tm_h.p + 4096)[l] = loci_type[l];
  }
  if (k) {
    bmg::copy_h2d(c->loci_dev.p, c->h_stage_i.p, k * sizeof(int64_t), st);
    bmg::copy_h2d(c->tm_d.p, c->tm_h.p, 2 * k * sizeof(double), st);
    bmg::copy_h2d(c->tm_d.p + 2048, c->tm_h.p + 2048, 2 * k * sizeof(double), st);
    bmg::copy_h2d(c->tm_i.p, c->tm_h.p + 4096, k * sizeof(int32_t), st);
  }
  if (prm->tau_mode == 1) {
    BMG_REQUIRE(prm->tau_host != nullptr, "bmg_chain_scan_types: tau_host required for tau_mode 1");
    const size_t need = (size_t)s->m * a.n_terms;
    if (c->tau_dev.n < need) c->tau_dev.alloc(need);
    bmg::copy_h2d(c->tau_dev.p, prm->tau_host, need * sizeof(double), st);
  }
  if (prm->n_types > 1 && c->p_r_types.n < (size_t)s->m * prm->n_types) c->p_r_types.alloc((size_t)s->m * 5);
  imma_quantize(c);
  imma_launch(c, false);
  imma_launch(c, true);
  if (s->n_missing > 0) {
    if (c->miss_corr4.n == 0) { c->miss_corr4.alloc(4 * s->m); c->miss_corr4.zero(st); }
    k_miss_corr4<<<(unsigned)((s->m * 32 + 127) / 128), 128, 0, st>>>(s->miss_off.p, s->miss_idx.p, c->miss_val.p, c->r.p, s->m,
                                                                      c->miss_corr4.p);
    count_launch();
  }
  a.part_a = c->imma_partial.p; a.part_h = c->imma_partial_h.p; a.n_chunks = c->imma_chunks;
  a.m = s->m; a.lo = s->lo; a.n = s->n; a.n1 = s->n1.p; a.n2 = s->n2.p; a.nmiss = s->nmiss.p;
  a.corr4 = s->n_missing > 0 ? c->miss_corr4.p : nullptr;
  a.loci = c->loci_dev.p; a.mtype = c->tm_i.p; a.mbeta = c->tm_d.p; a.mtau = c->tm_d.p + 2048; a.k = k;
  a.sum_r = c->sum_r; a.sigma2 = prm->sigma2; a.tau_mode = prm->tau_mode;
  for (int t = 0; t < 5; ++t) a.lmp_add[t] = prm->lmp_add[t];
  for (int t = 0; t < 25; ++t) a.lmp_rem[t] = prm->lmp_rem[t];
  for (int t = 0; t < 4; ++t) a.tau_shared[t] = prm->tau_shared[t];
  a.tau_snp = c->tau_dev.p;
  a.p_r = c->p_r.p; a.p_r_types = prm->n_types > 1 ? c->p_r_types.p : nullptr;
  k_scan_finalize_types<<<(unsigned)((s->m + 127) / 128), 128, 0, st>>>(a);
  count_launch();
  BMG_CUDA(cudaGetLastError());
  if (p_r_host) bmg::copy_d2h(p_r_host, c->p_r.p, s->m * sizeof(double), st);
  if (p_r_types_host && prm->n_types > 1) bmg::copy_d2h(p_r_types_host, c->p_r_types.p, (size_t)s->m * prm->n_types * sizeof(double), st);
  BMG_CUDA(cudaStreamSynchronize(st));
}

void chain_adapt(Chain* c, int update_rao, int64_t n_rao_mean, int update_prop, int64_t n_prop_mean, double q_add_min,
                 double q_rem_min)
{
  Store* s = c->store;
  BMG_CUDA(cudaSetDevice(s->device));
  const double rz1 = (double)(n_rao_mean + 1), rz2 = (double)n_rao_mean / rz1;      // sampler.cpp:740-741
  const double pz1 = (double)(n_prop_mean + 1), pz2 = (double)n_prop_mean / pz1;    // sampler.cpp:762-763
  k_adapt<<<(unsigned)((c->mw + 255) / 256), 256, 0, c->stream>>>(c->p_r.p, c->mw, update_rao, rz1, rz2, update_prop, pz1, pz2,
                                                                 q_add_min, q_rem_min, c->p_rao.p, c->p_proposal.p,
                                                                 c->q_add.p, c->q_rem.p);
  count_launch();
  BMG_CUDA(cudaGetLastError());
  if (update_prop) chain_partial_cdf(c);
}

void chain_init_flat(Chain* c, double value, double q_add_min, double q_rem_min)
{
  Store* s = c->store;
  BMG_CUDA(cudaSetDevice(s->device));
  k_fill_flat<<<(unsigned)((c->mw + 255) / 256), 256, 0, c->stream>>>(c->mw, value, q_add_min, q_rem_min, c->p_proposal.p,
                                                                     c->q_add.p, c->q_rem.p, c->p_rao.p);
  count_launch();
  BMG_CUDA(cudaGetLastError());
  chain_partial_cdf(c);
}

}  // namespace bmg
