// store.cu -- device-resident genotype store: load, recode, counts, moment cache, missing
// index, typed column decode.
//
// Replaces (reference paths): Data::read_g + recode_g_to_minor_allele_count + handle_missing_g
// + compute_g_var_and_mean (src/data.cpp:245-273,324-376,403-434), the four get_genotypes_*
// unpackers (src/data.cpp:56-138) and PrecomputedSNPCovariances::precompute
// (src/precomputed_snp_covariances.hpp:95-131).  The reference makes four per-element passes
// over the genotypes on one core at start-up; here it is one kernel over the raw .bed bytes
// (counts + recode decision + re-coding) plus one for the sparse missing index.
#include "common.cuh"
#include "store.cuh"

namespace bmg {

// ---------------------------------------------------------------------------------------
// raw .bed word for individuals 16w .. 16w+15 of one SNP (B bytes, possibly unaligned)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t raw_word(const uint8_t* __restrict__ col, int64_t B, int64_t w)
{
  const int64_t b0 = 4 * w;
  uint32_t x = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (b0 + k < B) x |= (uint32_t)col[b0 + k] << (8 * k);
  return x;
}

// mask with both bits of every 2-bit field whose individual index is < n
__device__ __forceinline__ uint32_t valid_mask(int64_t n, int64_t w)
{
  const int64_t first = 16 * w;
  if (first + 16 <= n) return 0xFFFFFFFFu;
  if (first >= n) return 0u;
  const int cnt = (int)(n - first);
  return (1u << (2 * cnt)) - 1u;
}

__device__ __forceinline__ int block_sum_int(int v, int* smem)
{
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  int t = 0;
  for (int i = 0; i < nw; ++i) t += smem[i];
  return t;
}

// One CTA per SNP.  PLINK codes as the reference reads them (src/data.cpp:36): 00 -> 0,
// 01 -> missing, 10 -> 1, 11 -> 2.  Output codes are the VALUE (00,01,10), missing -> 00.
__global__ void __launch_bounds__(128)
k_transcode_count(const uint8_t* __restrict__ raw, int64_t n, int64_t m, int64_t B, int64_t W, int64_t Wp,
                  int recode, uint32_t* __restrict__ codes, int32_t* __restrict__ n1, int32_t* __restrict__ n2,
                  int32_t* __restrict__ nmiss, uint8_t* __restrict__ swapped)
{
  __shared__ int red[4];
  const int64_t j = blockIdx.x;
  const uint8_t* col = raw + j * B;
  int c1 = 0, c2 = 0, cm = 0;
  for (int64_t w = threadIdx.x; w < W; w += blockDim.x) {
    const uint32_t x = raw_word(col, B, w) & valid_mask(n, w);
    const uint32_t H = (x >> 1) & 0x55555555u, L = x & 0x55555555u;
    c1 += __popc(H & ~L);
    c2 += __popc(H & L);
    cm += __popc(~H & L);
  }
  c1 = block_sum_int(c1, red);
  c2 = block_sum_int(c2, red);
  cm = block_sum_int(cm, red);
  const int ng = (int)n - cm;
  // src/utils.cpp:33-44 + src/data.cpp:330: macount / (2.0 * ngenos) > 0.5  <=>  macount > ngenos
  const bool swap = recode && ((c1 + 2 * c2) > ng);
  for (int64_t w = threadIdx.x; w < Wp; w += blockDim.x) {
    uint32_t out = 0;
    if (w < W) {
      const uint32_t vm = valid_mask(n, w);
      const uint32_t x = raw_word(col, B, w) & vm;
      const uint32_t H = (x >> 1) & 0x55555555u, L = x & 0x55555555u;
      const uint32_t one = H & ~L;
      const uint32_t two = swap ? (~H & ~L & 0x55555555u & vm) : (H & L);
      out = one | (two << 1);
    }
    codes[j * Wp + w] = out;
  }
  if (threadIdx.x == 0) {
    n1[j] = c1;
    n2[j] = swap ? (ng - c1 - c2) : c2;
    nmiss[j] = cm;
    swapped[j] = swap ? 1 : 0;
  }
}

// per SNP: moment cache (precomputed_snp_covariances.hpp:113-117, missing = 0) and the per-SNP
// mean / unbiased variance terms of Data::compute_g_var_and_mean (data.cpp:403-434)
__global__ void k_moments(const int32_t* __restrict__ n1, const int32_t* __restrict__ n2,
                          const int32_t* __restrict__ nmiss, int64_t n, int64_t m, double* __restrict__ mom,
                          double* __restrict__ snp_mean, double* __restrict__ snp_var)
{
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= m) return;
  const double s = (double)(n1[j] + 2 * n2[j]);
  const double xx = (double)(n1[j] + 4 * n2[j]);
  mom[2 * j] = s;
  mom[2 * j + 1] = __dsub_rn(xx, __ddiv_rn(__dmul_rn(s, s), (double)n));
  const int ng = (int)n - nmiss[j];
  snp_mean[j] = ng >= 1 ? __ddiv_rn(s, (double)ng) : nan("");
  snp_var[j] = ng > 1 ? __ddiv_rn(__dsub_rn(xx, __ddiv_rn(__dmul_rn(s, s), (double)ng)), (double)(ng - 1)) : nan("");
}

// One warp per SNP that has missing cells: ascending row indices of its 01 codes.
__global__ void k_missing_extract(const uint8_t* __restrict__ raw, int64_t n, int64_t m, int64_t B, int64_t W,
                                  const int64_t* __restrict__ miss_off, int32_t* __restrict__ miss_idx)
{
  const int64_t j = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (j >= m) return;
  int64_t pos = miss_off[j];
  if (miss_off[j + 1] == pos) return;
  const uint8_t* col = raw + j * B;
  for (int64_t w0 = 0; w0 < W; w0 += 32) {
    const int64_t w = w0 + lane;
    uint32_t mm = 0;
    if (w < W) {
      const uint32_t x = raw_word(col, B, w) & valid_mask(n, w);
      mm = ~((x >> 1) & 0x55555555u) & (x & 0x55555555u);
    }
    const int cnt = __popc(mm);
    int incl = cnt;
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    int64_t at = pos + incl - cnt;
    while (mm) {
      const int b = __ffs(mm) - 1;
      mm &= mm - 1;
      miss_idx[at++] = (int32_t)(16 * w + (b >> 1));
    }
    pos += __shfl_sync(0xffffffffu, incl, 31);
  }
}

// typed decode of one column (data.cpp:56-138 / genotype_tables.hpp): value v in {0,1,2}
__device__ __forceinline__ double typed(int v, int type)
{
  switch (type) {
    case 0: return (double)v;
    case 1: return v == 1 ? 1.0 : 0.0;
    case 2: return v > 0 ? 1.0 : 0.0;
    default: return v == 2 ? 1.0 : 0.0;
  }
}

__global__ void k_decode_column(const uint32_t* __restrict__ col, int64_t n, int type, double* __restrict__ out)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int v = (col[i >> 4] >> (2 * (i & 15))) & 3;
  out[i] = typed(v, type);
}

// missing cells: -1 (Data semantics) when vals == nullptr, else the typed imputed value
// (DataModel semantics, data_model.cpp:30-72)
__global__ void k_decode_fix_missing(const int32_t* __restrict__ idx, const int8_t* __restrict__ vals, int64_t cnt,
                                     int type, double* __restrict__ out)
{
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= cnt) return;
  out[idx[k]] = vals ? typed(vals[k], type) : -1.0;
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
void build_inorder_permutation(int64_t m, std::vector<int32_t>& order)
{
  // in-order traversal of the implicit heap (children 2i+1, 2i+2) the reference's
  // DiscreteDistribution builds over items 0..m-1 (discrete_distribution.hpp:217-261)
  order.resize(m);
  std::vector<int64_t> stack;
  stack.reserve(128);
  int64_t pos = 0, node = 0;
  while (!stack.empty() || node < m) {
    while (node < m) {
      stack.push_back(node);
      node = 2 * node + 1;
    }
    node = stack.back();
    stack.pop_back();
    order[pos++] = (int32_t)node;
    node = 2 * node + 2;
  }
}

// Genotype counts and missing-call index of ALL SNPs of a sharded data set, on this rank's device (collective: every rank
// calls it with the same arguments; `fn` is the in-place all-gather of the communicator).  The CSR is the one an unsharded
// store of the same file would hold -- cells in SNP order -- so per-cell counters (k_impute_from_prior) agree with it.
GlobalMissing* build_global_missing(Store* s, int world, int rank, int64_t stride, AllGatherFn fn, void* ctx)
{
  BMG_CUDA(cudaSetDevice(s->device));
  std::unique_ptr<GlobalMissing> gm(new GlobalMissing());
  gm->m_g = s->m_g;
  const int64_t total = (int64_t)world * stride, off0 = (int64_t)rank * stride;
  cudaStream_t st;
  BMG_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  auto gather_i32 = [&](const int32_t* local, DevBuf<int32_t>& all) {
    all.alloc(total);
    BMG_CUDA(cudaMemsetAsync(all.p, 0, total * sizeof(int32_t), st));
    BMG_CUDA(cudaMemcpyAsync(all.p + off0, local, s->m * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    if (world > 1 && fn(ctx, all.p, stride, (int)sizeof(int32_t), (void*)st) != 0) throw Error("sharded data set: all-gather of the genotype counts failed");
    BMG_CUDA(cudaStreamSynchronize(st));
  };
  gather_i32(s->n1.p, gm->n1);
  gather_i32(s->n2.p, gm->n2);
  gather_i32(s->nmiss.p, gm->nmiss);
  std::vector<int32_t> h_nm(total);
  BMG_CUDA(cudaMemcpy(h_nm.data(), gm->nmiss.p, total * sizeof(int32_t), cudaMemcpyDeviceToHost));
  gm->h_off.assign(s->m_g + 1, 0);
  for (int64_t j = 0; j < s->m_g; ++j) gm->h_off[j + 1] = gm->h_off[j] + h_nm[j];
  gm->n_missing = gm->h_off[s->m_g];
  gm->off.alloc(s->m_g + 1);
  BMG_CUDA(cudaMemcpy(gm->off.p, gm->h_off.data(), (s->m_g + 1) * sizeof(int64_t), cudaMemcpyHostToDevice));
  if (gm->n_missing > 0) {
    // the cells: every rank contributes its shard's list, padded to the longest
    auto cells_of = [&](int r) {
      const int64_t a = std::min<int64_t>(s->m_g, (int64_t)r * stride), b = std::min<int64_t>(s->m_g, (int64_t)(r + 1) * stride);
      return gm->h_off[b] - gm->h_off[a];
    };
    int64_t longest = 1;
    for (int r = 0; r < world; ++r) longest = std::max(longest, cells_of(r));
    BMG_REQUIRE(cells_of(rank) == s->n_missing, "sharded data set: the shard's missing-call index disagrees with the gathered counts");
    DevBuf<int32_t> pad;
    pad.alloc((size_t)world * longest);
    BMG_CUDA(cudaMemsetAsync(pad.p, 0, (size_t)world * longest * sizeof(int32_t), st));
    if (s->n_missing > 0)
      BMG_CUDA(cudaMemcpyAsync(pad.p + (int64_t)rank * longest, s->miss_idx.p, s->n_missing * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    if (world > 1 && fn(ctx, pad.p, longest, (int)sizeof(int32_t), (void*)st) != 0) throw Error("sharded data set: all-gather of the missing-call index failed");
    gm->idx.alloc((size_t)gm->n_missing);
    for (int r = 0; r < world; ++r) {
      const int64_t cnt = cells_of(r);
      if (cnt > 0)
        BMG_CUDA(cudaMemcpyAsync(gm->idx.p + gm->h_off[std::min<int64_t>(s->m_g, (int64_t)r * stride)], pad.p + (int64_t)r * longest,
                                 cnt * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    }
    BMG_CUDA(cudaStreamSynchronize(st));
  }
  cudaStreamDestroy(st);
  return gm.release();
}

const uint32_t* Store::column_ptr(int64_t snp) const
{
  if (snp >= lo && snp < hi) return codes.p + (snp - lo) * Wp;
  for (const PeerShard& p : peers)
    if (snp >= p.lo && snp < p.hi) return p.codes + (snp - p.lo) * Wp;
  throw Error("SNP " + std::to_string(snp) + " is neither in the local shard nor in an attached peer shard");
}

static std::unique_ptr<Store> store_open(int64_t n, int64_t m_g, int64_t lo, int64_t hi, bool recode, int device)
{
  BMG_REQUIRE(n > 0 && m_g > 0 && lo >= 0 && hi > lo && hi <= m_g, "bmg_store_create: invalid sizes");
  BMG_REQUIRE(n < (int64_t)1 << 31, "bmg_store_create: n must be < 2^31");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    throw Error("no CUDA device available: libbmagwa_b200 has no CPU fallback (" + std::string(cudaGetErrorString(e)) + ")");
  BMG_REQUIRE(device >= 0 && device < ndev, "bmg_store_create: invalid device index");
  BMG_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  BMG_CUDA(cudaGetDeviceProperties(&prop, device));
  BMG_REQUIRE(prop.major >= 10, "libbmagwa_b200 is built for sm_100a (B200) only; found sm_" +
                                    std::to_string(prop.major) + std::to_string(prop.minor));
  std::unique_ptr<Store> s(new Store());
  s->device = device;
  s->sm_count = prop.multiProcessorCount;
  s->n = n; s->m_g = m_g; s->lo = lo; s->hi = hi; s->m = hi - lo;
  s->W = words_for(n); s->Wp = stride_words_for(n);
  s->recode = recode;
  return s;
}

static Store* store_build(std::unique_ptr<Store> s, const uint8_t* raw);

Store* store_create(const uint8_t* bed, bool on_device, int64_t n, int64_t m_g, int64_t lo, int64_t hi, bool recode,
                    int device)
{
  std::unique_ptr<Store> s = store_open(n, m_g, lo, hi, recode, device);
  const int64_t m = s->m, B = (n + 3) / 4;
  // raw upload (freed at the end of this function)
  DevBuf<uint8_t> raw_own;
  const uint8_t* raw = bed;
  if (!on_device) {
    raw_own.alloc((size_t)(m * B));
    bmg::copy_h2d_sync(raw_own.p, bed, (size_t)(m * B));
    raw = raw_own.p;
  }
  return store_build(std::move(s), raw);
}

// The .bed file streamed to the device: two pinned 2 MiB buffers, the read of block b+1 overlaps the copy of
// block b, and no host copy of the payload is ever held (data.cpp:245-273 reads it genotype by genotype into
// n x m_g doubles).  Header checks as Data::read_g (data.cpp:250-262).
Store* store_create_from_bed(const char* path, int64_t n, int64_t m_g, int64_t lo, int64_t hi, bool recode, int device)
{
  BMG_REQUIRE(path != nullptr, "bmg_store_create_from_bed: null path");
  FILE* f = fopen(path, "rb");
  if (!f) throw Error("BED file could not be opened");
  struct Closer { FILE* f; ~Closer() { fclose(f); } } closer{f};
  unsigned char hdr[3] = {0, 0, 0};
  if (fread(hdr, 1, 3, f) != 3 || hdr[0] != 0x6C || hdr[1] != 0x1B) throw Error("BED file not recognised (magic number does not match)");
  if (hdr[2] != 0x01) throw Error("BED file not in snp-major format");
  std::unique_ptr<Store> s = store_open(n, m_g, lo, hi, recode, device);
  const int64_t m = s->m, B = (n + 3) / 4;
  const size_t total = (size_t)(m * B);
  if (fseeko(f, (off_t)(3 + lo * B), SEEK_SET) != 0) throw Error("Reading the BED file failed");
  const bool timing = getenv("BMG_TIMING") != nullptr;
  auto now = [] { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; };
  const double t_a = now();
  DevBuf<uint8_t> raw_own;
  raw_own.alloc(total);
  const double t_b = now();
  const size_t kBlock = (size_t)2 << 20;   // pinning costs ~0.7 ms per MiB: small staging buffers, many blocks
  PinnedBuf<uint8_t> stage[2];
  cudaEvent_t done[2];
  cudaStream_t up;
  BMG_CUDA(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    stage[i].alloc(std::min(kBlock, total));
    BMG_CUDA(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
  }
  const double t_c = now();
  bool ok = true;
  size_t off = 0;
  for (int b = 0; off < total; ++b) {
    const int i = b & 1;
    if (b >= 2) BMG_CUDA(cudaEventSynchronize(done[i]));
    const size_t len = std::min(kBlock, total - off);
    if (fread(stage[i].p, 1, len, f) != len) { ok = false; break; }
    BMG_CUDA(cudaMemcpyAsync(raw_own.p + off, stage[i].p, len, cudaMemcpyHostToDevice, up));
    BMG_CUDA(cudaEventRecord(done[i], up));
    g_h2d_bytes.fetch_add(len, std::memory_order_relaxed);
    off += len;
  }
  BMG_CUDA(cudaStreamSynchronize(up));
  for (int i = 0; i < 2; ++i) cudaEventDestroy(done[i]);
  cudaStreamDestroy(up);
  if (!ok) throw Error("Reading the BED file failed");
  if (timing)
    fprintf(stderr, "[bmg timing] store from bed: device alloc %.1f ms, pinned staging + stream %.1f ms, read + upload %.1f ms\n",
            1e3 * (t_b - t_a), 1e3 * (t_c - t_b), 1e3 * (now() - t_c));
  return store_build(std::move(s), raw_own.p);
}

static Store* store_build(std::unique_ptr<Store> s, const uint8_t* raw)
{
  const int64_t n = s->n, m = s->m, B = (n + 3) / 4;
  const bool recode = s->recode;
  s->codes.alloc((size_t)(m * s->Wp + 4096));  // slack: tiles may over-read past the last column
  BMG_CUDA(cudaMemset(s->codes.p + m * s->Wp, 0, 4096 * sizeof(uint32_t)));
  s->n1.alloc(m); s->n2.alloc(m); s->nmiss.alloc(m); s->swapped.alloc(m); s->mom.alloc(2 * m);
  k_transcode_count<<<(unsigned)m, 128>>>(raw, n, m, B, s->W, s->Wp, recode ? 1 : 0, s->codes.p, s->n1.p, s->n2.p,
                                          s->nmiss.p, s->swapped.p);
  count_launch();
  BMG_CUDA(cudaGetLastError());
  DevBuf<double>& snp_mean = s->snp_mean;
  DevBuf<double>& snp_var = s->snp_var;
  snp_mean.alloc(m); snp_var.alloc(m);
  k_moments<<<(unsigned)((m + 255) / 256), 256>>>(s->n1.p, s->n2.p, s->nmiss.p, n, m, s->mom.p, snp_mean.p, snp_var.p);
  count_launch();
  BMG_CUDA(cudaGetLastError());

  // missing index: exclusive scan of the counts on the host (m ints), then one extraction kernel
  std::vector<int32_t> h_nmiss(m);
  bmg::copy_d2h_sync(h_nmiss.data(), s->nmiss.p, m * sizeof(int32_t));
  s->h_miss_off.resize(m + 1);
  int64_t tot = 0;
  for (int64_t j = 0; j < m; ++j) { s->h_miss_off[j] = tot; tot += h_nmiss[j]; }
  s->h_miss_off[m] = tot;
  s->n_missing = tot;
  s->miss_off.alloc(m + 1);
  bmg::copy_h2d_sync(s->miss_off.p, s->h_miss_off.data(), (m + 1) * sizeof(int64_t));
  if (tot > 0) {
    s->miss_idx.alloc((size_t)tot);
    const int threads = 128;
    const int64_t blocks = (m * 32 + threads - 1) / threads;
    k_missing_extract<<<(unsigned)blocks, threads>>>(raw, n, m, B, s->W, s->miss_off.p, s->miss_idx.p);
    count_launch();
    BMG_CUDA(cudaGetLastError());
  }

  // Data::compute_g_var_and_mean: sequential sums over SNPs in index order, as the reference does
  std::vector<double> hm(m), hv(m);
  bmg::copy_d2h_sync(hm.data(), snp_mean.p, m * sizeof(double));
  bmg::copy_d2h_sync(hv.data(), snp_var.p, m * sizeof(double));
  double tm = 0, tv = 0, nm = 0, nv = 0;
  for (int64_t j = 0; j < m; ++j) {
    const int ng = (int)n - h_nmiss[j];
    if (ng > 1) { tm += hm[j]; tv += hv[j]; nm += 1; nv += 1; }
    else if (ng == 1) { tm += hm[j]; nm += 1; }
  }
  s->summaries[0] = tm; s->summaries[1] = nm; s->summaries[2] = tv; s->summaries[3] = nv;

  build_inorder_permutation(m, s->h_inorder);
  s->inorder.alloc(m);
  bmg::copy_h2d_sync(s->inorder.p, s->h_inorder.data(), m * sizeof(int32_t));
  BMG_CUDA(cudaDeviceSynchronize());
  return s.release();
}

void store_set_phenotype(Store* s, const double* y, const double* e, int m_e)
{
  BMG_REQUIRE(m_e >= 1, "bmg_store_set_phenotype: m_e must include the column of ones (>= 1)");
  BMG_CUDA(cudaSetDevice(s->device));
  s->m_e = m_e;
  s->y.alloc(s->n);
  s->e.alloc((size_t)s->n * m_e);
  bmg::copy_h2d_sync(s->y.p, y, s->n * sizeof(double));
  bmg::copy_h2d_sync(s->e.p, e, (size_t)s->n * m_e * sizeof(double));
  s->h_y.assign(y, y + s->n);
  // Data ctor (data.hpp:67-70): yy = y'y (ddot), var_y = VectorView::var (vector.cpp:112-123)
  double sq = 0, sum = 0;
  for (int64_t i = 0; i < s->n; ++i) { sq += y[i] * y[i]; sum += y[i]; }
  s->summaries[4] = (sq - sum * sum / (double)s->n) / (double)(s->n - 1);
  double yy = 0;
  for (int64_t i = 0; i < s->n; ++i) yy += y[i] * y[i];
  s->summaries[5] = yy;
}

void store_get_column(const Store* s, int64_t snp, int type, const int8_t* miss_vals_dev, bool overlay, double* out_host,
                      cudaStream_t st)
{
  BMG_REQUIRE(type >= 0 && type <= 3, "get_column: type must be 0..3 (A,H,D,R)");
  BMG_REQUIRE(s->is_local(snp), "get_column: SNP not in the local shard");
  BMG_CUDA(cudaSetDevice(s->device));
  DevBuf<double> out;
  out.alloc(s->n);
  const int64_t j = snp - s->lo;
  k_decode_column<<<(unsigned)((s->n + 255) / 256), 256, 0, st>>>(s->codes.p + j * s->Wp, s->n, type, out.p);
  count_launch();
  const int64_t cnt = s->h_miss_off[j + 1] - s->h_miss_off[j];
  if (cnt > 0) {
    const int8_t* v = overlay ? miss_vals_dev + s->h_miss_off[j] : nullptr;
    k_decode_fix_missing<<<(unsigned)((cnt + 127) / 128), 128, 0, st>>>(s->miss_idx.p + s->h_miss_off[j], v, cnt, type, out.p);
    count_launch();
  }
  BMG_CUDA(cudaGetLastError());
  bmg::copy_d2h(out_host, out.p, s->n * sizeof(double), st);
  BMG_CUDA(cudaStreamSynchronize(st));
}

}  // namespace bmg
