// scan_imma.cu -- the genotype scan on the integer tensor cores (scan variant 2).
//
// Same contract as k_scan_dots_tma in scan.cu -- dot_j = sum_i x_ij r_i for every SNP of the shard,
// each packed byte read once -- but the multiply-adds leave the FP64 pipe, which caps a
// one-DFMA-per-genotype kernel at ~1/3 of HBM bandwidth on B200 (profiles/round1_notes.md):
//
//   * the residual is converted ONCE per scan to 62-bit fixed point, q_i = round(r_i 2^S) with
//     S chosen from max|r|, and split into eight signed base-256 digits ("limbs"),
//       q_i = sum_b d_ib 256^b,  d_ib in [-128, 127];
//   * a tile of 16 SNPs x 32 individuals (int8 genotypes 0/1/2) times 32 individuals x 8 limbs is one
//     mma.sync.aligned.m16n8k32.s8.s8.s32 (SASS IMMA.16832.S8.S8): the 8 limb sums per SNP accumulate
//     exactly in int32; dot_j = 2^-S sum_b D_jb 256^b.  Integer arithmetic is associative, so the
//     result is independent of tiling, warp order and chunking (bit-reproducible), and its only error
//     is the 2^-62 max|r| quantisation -- smaller than the rounding of an fp64 accumulation.
//   * per thread, the limb fragments of its 512-individual slice stay in registers for the whole
//     kernel (8 x uint4); genotype words come from the bulk-async (TMA) stage ring;
//   * the 2-bit -> 8-bit expansion costs ONE integer instruction per four genotypes and no memory access:
//     w & (0x03030303 << 2p) leaves, in byte lane b, field p of packed byte b times 4^p (an unsigned byte
//     <= 128: the MMA's A operand is .u8).  Which individual sits at which k of the MMA is free as long as
//     the limb operand agrees, so the limb bytes of a word are stored in the order the four masks produce
//     (individuals 0,4,8,12 | 1,5,9,13 | 2,6,10,14 | 3,7,11,15).  The factor 4^p is paid back in the
//     quantisation: the individual at position p of its packed byte is quantised with exponent S - 2p,
//     so every product still carries 2^S.  Costs 6 of the 62 fixed-point bits.
//
// This is not a GEMM re-shaping of the problem: the "N" dimension of the MMA is the eight digits of ONE
// right-hand side, which is what makes a single-RHS GEMV fill the 16x8 tile.
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>
#include "common.cuh"
#include "store.cuh"

namespace bmg {

constexpr int kImmaTile = 16;     // SNPs per tile = M of the MMA
constexpr int kImmaStages = 4;
constexpr int kImmaStages2 = 6;   // two-residual kernel: two resident CTAs per SM instead of three, so a deeper ring keeps as many tiles in flight
constexpr int kImmaGroups = 8;    // 64-individual groups per warp kept in registers
constexpr int kImmaWarpWords = 4 * kImmaGroups;   // 32 packed words = 512 individuals per warp
constexpr int kImmaMaxWarps = 16;
constexpr int kImmaNonFinite = -(1 << 30);   // scale exponent standing for "the residual holds a NaN or an infinity"

__device__ __forceinline__ uint32_t smem_u32i(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init_i(uint64_t* bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32i(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx_i(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32i(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_i(uint64_t* bar, uint32_t parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32i(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s_i(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32i(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32i(bar))
               : "memory");
}
// field P of each of the four packed bytes of w, scaled by 4^P, one per byte lane: a single LOP3
// HET: only the low bit of each field, i.e. the heterozygote indicator [x == 1] (fields hold 0, 1 or 2): the same
// kernel then yields S1 = sum_{x=1} r, from which every effect type follows (A = S1 + 2 S2, H = S1,
// D = S1 + S2, R = S2)
template <int P, bool HET>
__device__ __forceinline__ uint32_t expand_field(uint32_t w) { return w & ((HET ? 0x01010101u : 0x03030303u) << (2 * P)); }

__device__ __forceinline__ void imma16832(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// ---- residual -> fixed point limbs ------------------------------------------------------------
// out[0] = S (as int), chosen so that |r_i| 2^S < 2^61 for every i
__global__ void __launch_bounds__(1024) k_absmax_exp(const double* __restrict__ r, int64_t n, int* __restrict__ scale_exp)
{
  __shared__ double sm[32];
  __shared__ int any_bad;
  if (threadIdx.x == 0) any_bad = 0;
  __syncthreads();
  double mx = 0.0;
  int bad = 0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = r[i];
    bad |= !isfinite(v);       // fmax drops NaNs: a diverged chain must not look healthy
    mx = fmax(mx, fabs(v));
  }
  if (bad) any_bad = 1;
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) mx = fmax(mx, sm[w]);
    int s = 0;
    if (mx > 0.0 && isfinite(mx)) s = 60 - ilogb(mx);
    scale_exp[0] = any_bad ? kImmaNonFinite : s;   // a non-finite residual makes every dot product NaN, as in the fp64 variants
  }
}

// Q layout: [word][limb 0..7][16 slots] bytes, i.e. one uint4 per (word, limb); individual j = 4 b + p of the word
// (packed byte b, position p) sits in slot 4 p + b, the order expand_field<0..3> delivers them
__global__ void k_quantize(const double* __restrict__ r, int64_t n, int64_t n_pad, const int* __restrict__ scale_exp,
                           int8_t* __restrict__ q)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  long long v = 0;
  const int se = scale_exp[0];
  if (i < n && se != kImmaNonFinite) v = __double2ll_rn(scalbn(r[i], se - 2 * (int)(i & 3)));   // 4^p is carried by the genotype operand
  int8_t* dst = q + ((i >> 4) * 8) * 16 + (4 * (i & 3) + ((i >> 2) & 3));
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const int8_t d = (int8_t)(v & 0xFF);   // low byte read as signed
    dst[b * 16] = d;
    v = (v - (long long)d) >> 8;
  }
}

struct ImmaArgs {
  const uint32_t* codes;
  int64_t Wp, m;
  const uint4* q;          // [words][8]
  const int* scale_exp;
  int chunk_words;         // 32 * warps
  int n_chunks;
  int64_t tiles;
  int slices;
  int row_stride;          // shared-memory row stride in words (chunk_words + 16: two rows x 16 words per LDS.128 phase hit 32 banks)
  double* out;             // [n_chunks][m]
  // second right-hand side of the two-residual kernel (k_scan_dots_imma2; unused otherwise)
  const uint4* q1;
  const int* scale_exp1;
  double* out1;
  int no_proxy_fence;      // development (tools/proxy_fence_ab.py): leave out the fence before a stage is handed back
};

__device__ __forceinline__ void mbar_arrive_i(uint64_t* bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32i(bar)) : "memory");
}

__device__ __forceinline__ void fence_cta() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }

// Warp-specialised: blockDim = 32 * (consumer warps + 1).  The last warp is the producer (one lane issues the
// bulk-async copies, gated by per-stage "empty" mbarriers); consumer warps never meet at a CTA barrier: each adds its
// int32 limb sums into a shared accumulator and the LAST warp to finish a tile (atomic ticket) combines and stores it.
// dynamic smem: NS * 16 * row_stride words | int acc[NS + 1][R][16][8] | int ticket[8] | mbarriers (NS = stages of the ring)
//
// R = 2 (k_scan_dots_imma2, shard groups): the same genotype words against the limbs of TWO residuals -- every packed byte
// is read and expanded once per pair of chains; each residual's sums are the integers the one-residual kernel forms, so
// the results are the same bits.
template <bool HET, int R, int NS>
__device__ __forceinline__ void scan_dots_imma_body(const ImmaArgs& a)
{
  constexpr int kAccBufs = NS + 1;
  static_assert(kAccBufs <= 8, "ticket array");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int n_cons = (blockDim.x >> 5) - 1;
  const int g = lane >> 2, tig = lane & 3;
  const int chunk = blockIdx.x % a.n_chunks, slice = blockIdx.x / a.n_chunks;
  const int64_t c0 = (int64_t)chunk * a.chunk_words;
  const int row_copy_words = (int)min((int64_t)a.chunk_words, a.Wp - c0);
  const int RS = a.row_stride;
  const int stage_words = kImmaTile * RS;

  uint32_t* stage0 = reinterpret_cast<uint32_t*>(smem_raw);
  int* acc = reinterpret_cast<int*>(stage0 + (size_t)NS * stage_words);
  int* ticket = acc + kAccBufs * R * kImmaTile * 8;
  uint64_t* full = reinterpret_cast<uint64_t*>(ticket + 8);
  uint64_t* empty = full + NS;

  const int64_t tile_lo = a.tiles * slice / a.slices, tile_hi = a.tiles * (slice + 1) / a.slices;
  const int my_tiles = (int)(tile_hi - tile_lo);

  if (t == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init_i(&full[s], 1); mbar_init_i(&empty[s], (uint32_t)n_cons); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int idx = t; idx < kAccBufs * R * kImmaTile * 8 + 8; idx += blockDim.x) acc[idx] = 0;   // accumulators and tickets
  __syncthreads();

  if (warp == n_cons) {
    // ---------------- producer warp ----------------
    // One lane issues the 16 row copies of a tile.  Its addresses are per-thread values, so ptxas moves them to uniform
    // registers copy by copy (~100 cycles per UBLKCP); a variant in which the whole warp runs this loop on warp-uniform
    // values and an elected lane issues straight from uniform registers was 4-6 % faster at 2.5 GB and no faster at the
    // C2 size, and the repeated-launch guard test failed once in six full runs with it (never in isolation, never again
    // once the proxy fence below was in): not shipped, see profiles/round2_notes.md 10.
    if (lane == 0) {
      const uint32_t rb = (uint32_t)row_copy_words * 4u;
      int s = 0;
      uint32_t round = 0;   // number of times the ring has wrapped
      for (int it = 0; it < my_tiles; ++it) {
        if (round) mbar_wait_i(&empty[s], (round - 1) & 1);
        const int64_t snp0 = (tile_lo + it) * kImmaTile;
        const int rows = (int)min((int64_t)kImmaTile, a.m - snp0);
        uint32_t* dst = stage0 + (size_t)s * stage_words;
        mbar_expect_tx_i(&full[s], rb * (uint32_t)rows);
        for (int rr = 0; rr < rows; ++rr) bulk_g2s_i(dst + (size_t)rr * RS, a.codes + (snp0 + rr) * a.Wp + c0, rb, &full[s]);
        if (++s == NS) { s = 0; ++round; }
      }
    }
    return;
  }

  // ---------------- consumer warps ----------------
  // this thread's words of a warp slice: 16 x + 4 tig + e (x < 2, e < 4) = one uint4 per x; limb g of each
  uint4 bq[R][kImmaGroups];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const uint4* q = r == 0 ? a.q : a.q1;
#pragma unroll
    for (int grp = 0; grp < kImmaGroups; ++grp)
      bq[r][grp] = q[(c0 + warp * kImmaWarpWords + 16 * (grp >> 2) + 4 * tig + (grp & 3)) * 8 + g];
  }
  int scale_exp[R];
#pragma unroll
  for (int r = 0; r < R; ++r) scale_exp[r] = (r == 0 ? a.scale_exp : a.scale_exp1)[0];
  const int word_off = warp * kImmaWarpWords + 4 * tig;

  int s = 0, buf = 0;
  uint32_t phase = 0;
  for (int it = 0; it < my_tiles; ++it) {
    mbar_wait_i(&full[s], phase);
    const uint32_t* row_lo = stage0 + (size_t)s * stage_words + (size_t)g * RS + word_off;
    const uint32_t* row_hi = row_lo + 8 * RS;
    uint32_t wl[kImmaGroups], wh[kImmaGroups];
#pragma unroll
    for (int x = 0; x < kImmaGroups / 4; ++x) {
      const uint4 vl = *reinterpret_cast<const uint4*>(row_lo + 16 * x);
      const uint4 vh = *reinterpret_cast<const uint4*>(row_hi + 16 * x);
      wl[4 * x] = vl.x; wl[4 * x + 1] = vl.y; wl[4 * x + 2] = vl.z; wl[4 * x + 3] = vl.w;
      wh[4 * x] = vh.x; wh[4 * x + 1] = vh.y; wh[4 * x + 2] = vh.z; wh[4 * x + 3] = vh.w;
    }
    // The refill of this stage is an async-proxy write (bulk copy) to memory this thread has just read through the generic
    // proxy; PTX orders accesses made through different proxies only across a proxy fence, in this write-after-read
    // direction too (CUTLASS places the same fence before releasing a TMA-loaded buffer its threads read with ld.shared).
    if (!a.no_proxy_fence) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) mbar_arrive_i(&empty[s]);   // this warp's words are in registers: the stage may be refilled
    int c[R][4], c2[R][4];   // two independent IMMA chains per residual
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int e = 0; e < 4; ++e) c[r][e] = c2[r][e] = 0;
#pragma unroll
    for (int grp = 0; grp < kImmaGroups; ++grp) {
      const uint32_t l0 = expand_field<0, HET>(wl[grp]), h0 = expand_field<0, HET>(wh[grp]), l1 = expand_field<1, HET>(wl[grp]),
                     h1 = expand_field<1, HET>(wh[grp]), l2 = expand_field<2, HET>(wl[grp]), h2 = expand_field<2, HET>(wh[grp]),
                     l3 = expand_field<3, HET>(wl[grp]), h3 = expand_field<3, HET>(wh[grp]);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        imma16832(c[r], l0, h0, l1, h1, bq[r][grp].x, bq[r][grp].y);
        imma16832(c2[r], l2, h2, l3, h3, bq[r][grp].z, bq[r][grp].w);
      }
    }
    int* tile_acc = acc + buf * R * kImmaTile * 8;
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
      for (int e = 0; e < 4; ++e) c[r][e] += c2[r][e];
      int* ta = tile_acc + r * kImmaTile * 8;
      atomicAdd(&ta[g * 8 + 2 * tig], c[r][0]);
      atomicAdd(&ta[g * 8 + 2 * tig + 1], c[r][1]);
      atomicAdd(&ta[(g + 8) * 8 + 2 * tig], c[r][2]);
      atomicAdd(&ta[(g + 8) * 8 + 2 * tig + 1], c[r][3]);
    }
    fence_cta();
    __syncwarp();
    int last = 0;
    if (lane == 0) last = (atomicAdd(&ticket[buf], 1) == n_cons - 1);
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last) {   // every other warp's sums are in (their fences precede their tickets)
      fence_cta();
      if (lane < R * kImmaTile) {   // lane = 16 r + SNP of the tile
        const int r = R == 1 ? 0 : lane >> 4;
        const int64_t snp = (tile_lo + it) * kImmaTile + (lane & (kImmaTile - 1));
        volatile int* d = tile_acc + lane * 8;
        const long long lo = (long long)d[0] + ((long long)d[1] << 8) + ((long long)d[2] << 16) + ((long long)d[3] << 24);
        const long long hi = (long long)d[4] + ((long long)d[5] << 8) + ((long long)d[6] << 16) + ((long long)d[7] << 24);
#pragma unroll
        for (int b = 0; b < 8; ++b) d[b] = 0;
        // 2^-S in two always-normal factors (S spans about +-1100)
        const int se = R == 1 ? scale_exp[0] : (r == 0 ? scale_exp[0] : scale_exp[R - 1]);
        const int h0 = se / 2, h1 = se - h0;
        const double f0 = __hiloint2double((1023 - h0) << 20, 0), f1 = __hiloint2double((1023 - h1) << 20, 0);
        double* out = R == 1 || r == 0 ? a.out : a.out1;
        if (snp < a.m)
          out[(int64_t)chunk * a.m + snp] = se == kImmaNonFinite ? __longlong_as_double(0x7ff8000000000000ll)
                                                                 : fma((double)hi, 4294967296.0, (double)lo) * f0 * f1;
      }
      __syncwarp();
      if (lane == 0) { fence_cta(); ticket[buf] = 0; }
    }
    if (++s == NS) { s = 0; phase ^= 1u; }
    if (++buf == kAccBufs) buf = 0;
  }
}

template <bool HET>
__global__ void __maxnreg__(80) k_scan_dots_imma(const __grid_constant__ ImmaArgs a)
{
  scan_dots_imma_body<HET, 1, kImmaStages>(a);
}
// two residuals per pass: 64 limb registers per thread, so fewer resident warps than the one-residual kernel
__global__ void __maxnreg__(128) k_scan_dots_imma2(const __grid_constant__ ImmaArgs a)
{
  scan_dots_imma_body<false, 2, kImmaStages2>(a);
}

// ---- host side ----------------------------------------------------------------------------------
void imma_choose_geometry(Chain* c)
{
  const int64_t W = c->store->W;
  double best = -1.0;
  int best_nw = 1;
  for (int nw = 1; nw <= kImmaMaxWarps; ++nw) {
    const int64_t chunks = (W + (int64_t)kImmaWarpWords * nw - 1) / ((int64_t)kImmaWarpWords * nw);
    double eff = (double)W / (double)(chunks * kImmaWarpWords * nw);
    // resident warps per SM: 80 registers/thread and the stage ring in shared memory
    const int by_regs = 65536 / (80 * 32 * (nw + 1));
    const size_t smem = (size_t)kImmaStages * kImmaTile * (kImmaWarpWords * nw + 16) * 4 + 2048;
    const int by_smem = (int)((227 * 1024) / smem);
    const int ctas = by_regs < by_smem ? by_regs : by_smem;
    if (ctas < 1) continue;
    const double occ = (double)(ctas * nw) / 20.0;
    eff *= occ < 1.0 ? occ : 1.0;
    if (eff > best + 1e-9) { best = eff; best_nw = nw; }
  }
  if (const char* env = getenv("BMG_IMMA_WARPS")) {
    const int v = atoi(env);
    if (v >= 1 && v <= kImmaMaxWarps) best_nw = v;
  }
  c->imma_warps = best_nw;
  c->imma_chunk_words = kImmaWarpWords * best_nw;
  c->imma_chunks = (int)((W + c->imma_chunk_words - 1) / c->imma_chunk_words);
}

static size_t imma_smem_bytes(const Chain* c, int n_rhs = 1)
{
  const int RS = (int)c->imma_chunk_words + 16;
  const int stages = n_rhs == 2 ? kImmaStages2 : kImmaStages;
  return (size_t)stages * kImmaTile * RS * 4 + (size_t)(stages + 1) * n_rhs * kImmaTile * 8 * sizeof(int) + 8 * sizeof(int) +
         2 * stages * sizeof(uint64_t);
}

void imma_prepare(Chain* c)
{
  Store* s = c->store;
  imma_choose_geometry(c);
  const int64_t n_pad = 16 * (int64_t)c->imma_chunks * c->imma_chunk_words;
  c->imma_q.alloc((size_t)n_pad * 8);
  c->imma_exp.alloc(4);
  if ((int64_t)c->imma_partial.n < (int64_t)c->imma_chunks * s->m) c->imma_partial.alloc((size_t)c->imma_chunks * s->m);
  const size_t smem = imma_smem_bytes(c);
  BMG_REQUIRE(smem <= 227 * 1024, "IMMA scan tile does not fit in shared memory");
  BMG_CUDA(cudaFuncSetAttribute(k_scan_dots_imma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  BMG_CUDA(cudaFuncSetAttribute(k_scan_dots_imma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  BMG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_scan_dots_imma<false>, 32 * (c->imma_warps + 1), smem));
  if (per_sm < 1) per_sm = 1;
  if (const char* env = getenv("BMG_IMMA_CTAS_PER_SM")) {
    const int v = atoi(env);
    if (v >= 1 && v <= per_sm) per_sm = v;
  }
  const int64_t tiles = (s->m + kImmaTile - 1) / kImmaTile;
  int64_t slices = ((int64_t)per_sm * s->sm_count) / c->imma_chunks;
  if (slices < 1) slices = 1;
  if (slices > tiles) slices = tiles;
  c->imma_slices = (int)slices;
  // the two-residual kernel: same warps and chunks (so the same limb layout), its own number of resident CTAs
  const size_t smem2 = imma_smem_bytes(c, 2);
  c->imma_slices2 = 0;
  if (smem2 <= 227 * 1024 && getenv("BMG_IMMA_SINGLE") == nullptr) {
    BMG_CUDA(cudaFuncSetAttribute(k_scan_dots_imma2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    int per_sm2 = 0;
    BMG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm2, k_scan_dots_imma2, 32 * (c->imma_warps + 1), smem2));
    if (per_sm2 >= 1) {
      int64_t slices2 = ((int64_t)per_sm2 * s->sm_count) / c->imma_chunks;
      if (slices2 < 1) slices2 = 1;
      if (slices2 > tiles) slices2 = tiles;
      c->imma_slices2 = (int)slices2;
    }
  }
  c->imma_ready = true;
}

// quantise the current residual into limbs (once per residual)
void imma_quantize(Chain* c)
{
  Store* s = c->store;
  if (!c->imma_ready) imma_prepare(c);
  if (c->imma_q_valid) return;
  cudaStream_t st = c->stream;
  const int64_t n_pad = 16 * (int64_t)c->imma_chunks * c->imma_chunk_words;
  k_absmax_exp<<<1, 1024, 0, st>>>(c->r.p, s->n, c->imma_exp.p);
  k_quantize<<<(unsigned)((n_pad + 255) / 256), 256, 0, st>>>(c->r.p, s->n, n_pad, c->imma_exp.p,
                                                              reinterpret_cast<int8_t*>(c->imma_q.p));
  count_launch(2);
  c->imma_q_valid = true;
}

static int imma_no_proxy_fence()
{
  static const int v = getenv("BMG_IMMA_NO_PROXY_FENCE") != nullptr;
  return v;
}

// the tensor-core scan of this store's SNPs against ONE quantised right-hand side (q, scale_exp: the layout k_quantize
// writes; they may belong to another chain, group.cu) into out ([chunks][m] doubles), with the geometry of `geom`
void imma_launch_on(Chain* geom, const uint4* q, const int* scale_exp, double* out, bool het, cudaStream_t st, Chain* timed)
{
  Store* s = geom->store;
  if (!geom->imma_ready) imma_prepare(geom);
  ImmaArgs a;
  a.codes = s->codes.p; a.Wp = s->Wp; a.m = s->m; a.q = q; a.scale_exp = scale_exp;
  a.chunk_words = (int)geom->imma_chunk_words; a.n_chunks = geom->imma_chunks; a.tiles = (s->m + kImmaTile - 1) / kImmaTile;
  a.slices = geom->imma_slices; a.row_stride = (int)geom->imma_chunk_words + 16;
  a.out = out;
  a.q1 = nullptr; a.scale_exp1 = nullptr; a.out1 = nullptr;
  a.no_proxy_fence = imma_no_proxy_fence();
  const unsigned grid = (unsigned)(geom->imma_chunks * geom->imma_slices);
  if (timed) scan_timer_begin(timed, st);
  if (het) k_scan_dots_imma<true><<<grid, 32 * (geom->imma_warps + 1), imma_smem_bytes(geom), st>>>(a);
  else k_scan_dots_imma<false><<<grid, 32 * (geom->imma_warps + 1), imma_smem_bytes(geom), st>>>(a);
  if (timed) scan_timer_end(timed, st);
  count_launch();
}

// the same scan against TWO quantised right-hand sides in one pass over the shard (shard groups: two chains' residuals);
// false when the geometry has no room for the two-residual kernel (the caller then launches twice)
bool imma_launch2_on(Chain* geom, const uint4* q0, const int* scale_exp0, double* out0, const uint4* q1, const int* scale_exp1,
                     double* out1, cudaStream_t st, Chain* timed)
{
  Store* s = geom->store;
  if (!geom->imma_ready) imma_prepare(geom);
  if (geom->imma_slices2 < 1) return false;
  ImmaArgs a;
  a.codes = s->codes.p; a.Wp = s->Wp; a.m = s->m; a.q = q0; a.scale_exp = scale_exp0; a.out = out0;
  a.q1 = q1; a.scale_exp1 = scale_exp1; a.out1 = out1;
  a.no_proxy_fence = imma_no_proxy_fence();
  a.chunk_words = (int)geom->imma_chunk_words; a.n_chunks = geom->imma_chunks; a.tiles = (s->m + kImmaTile - 1) / kImmaTile;
  a.slices = geom->imma_slices2; a.row_stride = (int)geom->imma_chunk_words + 16;
  const unsigned grid = (unsigned)(geom->imma_chunks * geom->imma_slices2);
  if (timed) scan_timer_begin(timed, st);
  k_scan_dots_imma2<<<grid, 32 * (geom->imma_warps + 1), imma_smem_bytes(geom, 2), st>>>(a);
  if (timed) scan_timer_end(timed, st);
  count_launch();
  return true;
}

// the chain's own scan into imma_partial; het: the heterozygote-indicator pass into imma_partial_h
void imma_launch(Chain* c, bool het)
{
  Store* s = c->store;
  if (het) {
    if ((int64_t)c->imma_partial_h.n < (int64_t)c->imma_chunks * s->m) {
      BMG_CUDA(cudaStreamSynchronize(c->stream));
      c->imma_partial_h.alloc((size_t)c->imma_chunks * s->m);
    }
    imma_launch_on(c, reinterpret_cast<const uint4*>(c->imma_q.p), c->imma_exp.p, c->imma_partial_h.p, true, c->stream, nullptr);
    return;
  }
  imma_launch_on(c, reinterpret_cast<const uint4*>(c->imma_q.p), c->imma_exp.p, c->imma_partial.p, false, c->stream, nullptr);
  c->last_partial = c->imma_partial.p;
  c->last_chunks = c->imma_chunks;
}

}  // namespace bmg
