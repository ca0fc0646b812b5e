// weights.cu -- proposal weights as device partial CDFs, and DiscreteDistribution::sample()
// evaluated on the device.
//
// The reference's DiscreteDistribution (src/discrete_distribution.hpp:64-330) is a threaded
// binary tree over heap indices whose sample() is an in-order cumulative search.
// Here the weights stay in natural SNP order; `inorder[pos]` maps an in-order position to the
// SNP, block sums over 256 consecutive in-order positions are the "partial CDFs" (what shards
// exchange), and zeroing an item (adddate/remdate, :156-201) flips a byte flag and refreshes
// one block sum.
#include "common.cuh"
#include "store.cuh"

namespace bmg {

__device__ __forceinline__ double block_reduce_256(double v, double* sm)
{
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sm[w];
  return s;
}

// block b: sums of q over in-order positions [256 b, 256 (b+1)); raw and zero-aware
__global__ void __launch_bounds__(256) k_block_sums(const double* __restrict__ q, const uint8_t* __restrict__ zeroed,
                                                    const int32_t* __restrict__ inorder, int64_t m,
                                                    double* __restrict__ raw, double* __restrict__ eff,
                                                    double* __restrict__ q_inorder)
{
  __shared__ double sm[8];
  const int64_t pos = (int64_t)blockIdx.x * 256 + threadIdx.x;
  double w = 0.0, we = 0.0;
  if (pos < m) {
    const int32_t j = inorder[pos];
    w = q[j];
    we = zeroed[j] ? 0.0 : w;
    q_inorder[pos] = w;
  }
  const double s = block_reduce_256(w, sm);
  const double se = block_reduce_256(we, sm);
  if (threadIdx.x == 0) { raw[blockIdx.x] = s; eff[blockIdx.x] = se; }
}

__global__ void k_set_zero(uint8_t* __restrict__ zeroed, int64_t j, int flag) { zeroed[j] = (uint8_t)flag; }

// single CTA of 256 threads.  out[0] = sampled local SNP (-1 if every weight is zero), out[1] = total
__global__ void __launch_bounds__(256) k_sample(const double* __restrict__ q, const uint8_t* __restrict__ zeroed,
                                                const int32_t* __restrict__ inorder, const double* __restrict__ eff,
                                                int64_t m, int64_t nb, double u01, double* __restrict__ out)
{
  __shared__ double sm[8];
  __shared__ double scan[256];
  __shared__ double s_total, s_carry;
  __shared__ long long s_block;
  __shared__ int s_found;
  const int t = threadIdx.x;
  // 1. per-thread sums of contiguous groups of blocks, then an inclusive scan across threads
  const int64_t G = (nb + 255) / 256;
  const int64_t b0 = t * G, b1 = min(nb, b0 + G);
  double mine = 0.0;
  for (int64_t b = b0; b < b1; ++b) mine += eff[b];
  scan[t] = mine;
  __syncthreads();
  if (t == 0) {
    double run = 0.0;
    for (int i = 0; i < 256; ++i) { run += scan[i]; scan[i] = run; }
    s_total = run;
    s_block = -1;
  }
  __syncthreads();
  const double total = s_total;
  const double r = u01 * total;                      // discrete_distribution.hpp:127
  // 2. the thread whose group contains r walks its blocks
  const double before = t == 0 ? 0.0 : scan[t - 1];
  if (b0 < b1 && r >= before && r < scan[t]) {
    double run = before;
    long long found = b1 - 1;
    for (int64_t b = b0; b < b1; ++b) {
      if (r < run + eff[b]) { found = b; break; }
      run += eff[b];
    }
    s_block = found;
    s_carry = run;
  }
  __syncthreads();
  long long blk = s_block;
  double carry = s_carry;
  if (blk < 0) {  // r >= total through rounding: fall back to the last block (tree's final "upright parent" return)
    blk = nb - 1;
    carry = total - eff[nb - 1];
  }
  // 3. in-block search over 256 in-order positions (strict "r < cumulative", :131-152)
  const int64_t pos = blk * 256 + t;
  double w = 0.0;
  bool live = false;
  if (pos < m) {
    const int32_t j = inorder[pos];
    live = !zeroed[j];
    w = live ? q[j] : 0.0;
  }
  scan[t] = w;
  sm[0] = 0.0;
  __shared__ unsigned char live_s[256];
  live_s[t] = live ? 1 : 0;
  if (t == 0) s_found = -1;
  __syncthreads();
  if (t == 0) {
    double run = carry;
    int last = -1;
    for (int i = 0; i < 256; ++i) {
      if (!live_s[i]) continue;
      run += scan[i];
      last = i;
      if (r < run) { s_found = i; break; }
    }
    if (s_found < 0) s_found = last;
  }
  __syncthreads();
  if (t == 0) {
    const int f = s_found;
    out[0] = f >= 0 ? (double)inorder[blk * 256 + f] : -1.0;
    out[1] = total;
  }
  (void)sm;
}

void chain_partial_cdf(Chain* c)
{
  Store* s = c->store;
  BMG_CUDA(cudaSetDevice(s->device));
  if (c->cdf_eff_add.n == 0) {
    c->cdf_eff_add.alloc(c->cdf_blocks); c->cdf_eff_rem.alloc(c->cdf_blocks);
    c->q_add_io.alloc(c->mw); c->q_rem_io.alloc(c->mw);
  }
  k_block_sums<<<(unsigned)c->cdf_blocks, 256, 0, c->stream>>>(c->q_add.p, c->zero_add.p, c->inorder_dev(), c->mw, c->cdf_add.p,
                                                              c->cdf_eff_add.p, c->q_add_io.p);
  k_block_sums<<<(unsigned)c->cdf_blocks, 256, 0, c->stream>>>(c->q_rem.p, c->zero_rem.p, c->inorder_dev(), c->mw, c->cdf_rem.p,
                                                              c->cdf_eff_rem.p, c->q_rem_io.p);
  count_launch(2);
  BMG_CUDA(cudaGetLastError());
}

void chain_set_zeroed(Chain* c, int which, int64_t snp, int flag)
{
  Store* s = c->store;
  BMG_REQUIRE(which == 0 || which == 1, "which must be 0 (dd_add) or 1 (dd_rem)");
  const int64_t base = s->lo - c->w_off;   // global index of element 0 of the weight arrays
  BMG_REQUIRE(snp >= base && snp < base + c->mw, "bmg_chain_set_zeroed: SNP outside the chain's weight arrays");
  BMG_CUDA(cudaSetDevice(s->device));
  if (c->cdf_eff_add.n == 0) chain_partial_cdf(c);
  uint8_t* z = which == 0 ? c->zero_add.p : c->zero_rem.p;
  k_set_zero<<<1, 1, 0, c->stream>>>(z, snp - base, flag ? 1 : 0);
  count_launch();
  // refresh the zero-aware partial sums (all blocks: m/256 tiny CTAs; keeps the kernel count at two)
  if (which == 0)
    k_block_sums<<<(unsigned)c->cdf_blocks, 256, 0, c->stream>>>(c->q_add.p, c->zero_add.p, c->inorder_dev(), c->mw, c->cdf_add.p,
                                                                c->cdf_eff_add.p, c->q_add_io.p);
  else
    k_block_sums<<<(unsigned)c->cdf_blocks, 256, 0, c->stream>>>(c->q_rem.p, c->zero_rem.p, c->inorder_dev(), c->mw, c->cdf_rem.p,
                                                                c->cdf_eff_rem.p, c->q_rem_io.p);
  count_launch();
  BMG_CUDA(cudaGetLastError());
}

void chain_fill_zeroed(Chain* c, int which, int flag)
{
  Store* s = c->store;
  BMG_REQUIRE(which == 0 || which == 1, "which must be 0 (dd_add) or 1 (dd_rem)");
  BMG_CUDA(cudaSetDevice(s->device));
  uint8_t* z = which == 0 ? c->zero_add.p : c->zero_rem.p;
  BMG_CUDA(cudaMemsetAsync(z, flag ? 1 : 0, c->mw, c->stream));
  chain_partial_cdf(c);
}

void chain_sample(Chain* c, int which, double u01, int64_t* snp, double* total)
{
  Store* s = c->store;
  BMG_REQUIRE(which == 0 || which == 1, "which must be 0 (dd_add) or 1 (dd_rem)");
  BMG_REQUIRE(u01 >= 0.0 && u01 < 1.0, "bmg_chain_sample: u01 must be in [0,1)");
  BMG_CUDA(cudaSetDevice(s->device));
  if (c->cdf_eff_add.n == 0) chain_partial_cdf(c);
  const double* q = which == 0 ? c->q_add.p : c->q_rem.p;
  const uint8_t* z = which == 0 ? c->zero_add.p : c->zero_rem.p;
  const double* eff = which == 0 ? c->cdf_eff_add.p : c->cdf_eff_rem.p;
  k_sample<<<1, 256, 0, c->stream>>>(q, z, c->inorder_dev(), eff, c->mw, c->cdf_blocks, u01, c->sample_out.p);
  count_launch();
  BMG_CUDA(cudaGetLastError());
  bmg::copy_d2h(c->h_sample.p, c->sample_out.p, 2 * sizeof(double), c->stream);
  BMG_CUDA(cudaStreamSynchronize(c->stream));
  const double j = c->h_sample.p[0];
  BMG_REQUIRE(j >= 0, "bmg_chain_sample: every item is zeroed");
  *snp = (int64_t)j + (s->lo - c->w_off);
  if (total) *total = c->h_sample.p[1];
}

}  // namespace bmg
