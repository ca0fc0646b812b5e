// overlay.cu -- per-chain cache of packed columns with the chain's imputed genotypes filled in.
//
// The store keeps a missing call as 00 plus an entry in a sparse index (common.cuh); the chain's imputed values
// (DataModel::miss_val_, src/data_model.hpp:75-101) are one byte per missing cell.  The reference materialises a
// column with its imputed values every time it unpacks one (DataModel::get_genotypes_*, src/data_model.cpp:30-72).
// Here a column that is about to be used by the per-move statistics or the fitted values is materialised ONCE, in
// its packed 2-bit form, into a slot of this cache (copy of the store column + OR of the imputed 2-bit values), and
// the statistics kernels read it exactly like a store column: no sparse correction pass, same latency path.
//   * a slot is (re)built when the SNP's imputed values are set (bmg_chain_set_missing: a proposed addition,
//     data_model.cpp:95-103, or the Gibbs step, sampler.cpp:264-453) -- the kernel takes the values straight from
//     mapped pinned host memory, stores them into the chain's value array and patches the slot, one launch for all
//     SNPs of the move;
//   * or on first use after the values changed in bulk (bmg_chain_set_missing_all / _impute_from_prior drop the
//     slots of every SNP outside the model), from the values already on the device.
// Slots are handed out round-robin; a request never evicts a column it is itself using.
#include <string.h>
#include <algorithm>
#include "common.cuh"
#include "store.cuh"

namespace bmg {

constexpr int kPatchPerLaunch = 24;
constexpr int kOverlaySlots = 2048 + 256 + 8;   // largest model + largest candidate list of one request

struct PatchArgs {
  PatchDesc d[kPatchPerLaunch];
  int64_t quads;         // 16-byte units per column (padded stride / 4)
};

__global__ void __launch_bounds__(256) k_patch_columns(const __grid_constant__ PatchArgs a)
{
  const PatchDesc& d = a.d[blockIdx.x];
  const uint4* s4 = reinterpret_cast<const uint4*>(d.src);
  uint4* d4 = reinterpret_cast<uint4*>(d.dst);
  for (int64_t w = threadIdx.x; w < a.quads; w += blockDim.x) d4[w] = s4[w];
  __syncthreads();   // the copy (global writes of this CTA) is ordered before the ORs below
  for (int64_t q = threadIdx.x; q < d.cnt; q += blockDim.x) {
    const int v = d.vals[q];
    if (d.keep) d.keep[q] = (int8_t)v;
    const int32_t i = d.idx[q];
    if (v) atomicOr(&d.dst[i >> 4], (uint32_t)v << (2 * (i & 15)));
  }
  if (d.type == 0) return;
  // typed column (DataModel::get_genotypes_heterozygous / dominant / recessive, data_model.cpp:41-72) as 0/1 fields:
  // a field is 00, 01 or 10, so [x == 1] is its low bit, [x == 2] its high bit, [x > 0] either
  __syncthreads();
  uint32_t* w32 = d.dst;
  for (int64_t w = threadIdx.x; w < 4 * a.quads; w += blockDim.x) {
    const uint32_t v = w32[w], lo = v & 0x55555555u, hi = (v >> 1) & 0x55555555u;
    w32[w] = d.type == 1 ? lo : (d.type == 2 ? (lo | hi) : hi);
  }
}

static void overlay_prepare(Chain* c)
{
  if (c->pc_slots) return;
  Store* s = c->store;
  chain_server_stop(c);   // allocations synchronise the device: a running server would stall them
  BMG_CUDA(cudaStreamSynchronize(c->stream));
  c->pc_cols.alloc((size_t)kOverlaySlots * s->Wp);
  c->pc_slots = kOverlaySlots;
  c->pc_snp.assign(kOverlaySlots, -1);
  c->pc_use.assign(kOverlaySlots, 0);
  c->pc_seq = 0;
  c->pc_next = 0;
}

// Round-robin with a second chance: the model's SNPs take part in every request, so their slots always carry a recent
// sequence number and survive; the slots of proposed SNPs that were not accepted go stale and are reused.
static int overlay_take_slot(Chain* c)
{
  constexpr unsigned int kRecent = 128;   // requests
  for (int tries = 0; tries < 2 * c->pc_slots; ++tries) {
    const int slot = c->pc_next;
    c->pc_next = (slot + 1) % c->pc_slots;
    if (c->pc_use[slot] == c->pc_seq) continue;   // in use by the request being resolved
    if (tries < c->pc_slots && c->pc_snp[slot] >= 0 && c->pc_seq - c->pc_use[slot] < kRecent) continue;
    if (c->pc_snp[slot] >= 0) c->pc_map.erase(c->pc_snp[slot]);
    return slot;
  }
  throw Error("overlay cache: more columns in one request than slots");
}

static void overlay_launch(Chain* c, const std::vector<PatchDesc>& pend)
{
  Store* s = c->store;
  for (size_t done = 0; done < pend.size(); done += kPatchPerLaunch) {
    const int cnt = (int)std::min<size_t>(kPatchPerLaunch, pend.size() - done);
    PatchArgs a;
    memset(&a, 0, sizeof(a));
    for (int i = 0; i < cnt; ++i) a.d[i] = pend[done + i];
    a.quads = s->Wp / 4;
    k_patch_columns<<<cnt, 256, 0, c->stream>>>(a);
    count_launch();
  }
  BMG_CUDA(cudaGetLastError());
}

// Device pointers of `count` packed columns as the chain sees them (imputed values applied).  host_vals == nullptr:
// columns whose slot is missing are rebuilt from the values on the device.  host_vals[i] != nullptr: SNP i gets these
// new values (count = its number of missing cells), stored on the device and patched into its slot in the same launch.
// Returns the number of columns (re)built; the work is queued on the chain's stream.
int chain_overlay_columns(Chain* c, const int64_t* snps, int count, const uint32_t** out, const int8_t* const* host_vals,
                          const int32_t* types)
{
  Store* s = c->store;
  bool any_typed = false;
  if (types)
    for (int i = 0; i < count; ++i) {
      BMG_REQUIRE(types[i] >= 0 && types[i] <= 3, "effect type of a column must be 0 (A), 1 (H), 2 (D) or 3 (R)");
      any_typed = any_typed || types[i] != 0;
    }
  const MissView& mv = c->mv;
  if (mv.n_missing == 0 && !any_typed) {
    for (int i = 0; i < count; ++i) out[i] = s->column_ptr(snps[i]);
    return 0;
  }
  BMG_REQUIRE(count <= kOverlaySlots - 8, "overlay cache: too many columns in one request");
  overlay_prepare(c);
  if (++c->pc_seq == 0) {   // wrapped: forget the marks
    std::fill(c->pc_use.begin(), c->pc_use.end(), 0u);
    c->pc_seq = 1;
  }
  std::vector<PatchDesc>& pend = c->pc_pending;
  pend.clear();
  size_t staged = 0;
  if (host_vals) {
    size_t total = 0;
    for (int i = 0; i < count; ++i)
      if (host_vals[i] && mv.covers(snps[i]) && mv.n_missing > 0) total += (size_t)mv.count(snps[i]);
    if (c->pc_h_vals.n < total) {
      chain_server_stop(c);
      BMG_CUDA(cudaStreamSynchronize(c->stream));
      c->pc_h_vals.alloc(2 * total + 4096);
    }
  }
  // first the columns that already have a slot (so that none of them is evicted below) ...
  for (int i = 0; i < count; ++i) {
    const int64_t snp = snps[i];
    out[i] = nullptr;
    if (!mv.covers(snp)) {   // a column outside the chain's index (a peer's, on a chain that keeps no global index)
      BMG_REQUIRE(!types || types[i] == 0, "typed columns of a peer shard are not available");
      out[i] = s->column_ptr(snp);
      continue;
    }
    const int ty = types ? types[i] : 0;
    BMG_REQUIRE(!(host_vals && host_vals[i] && ty != 0), "new imputed values are given for the additive column");
    const int64_t j = snp - mv.base, cnt = mv.n_missing > 0 ? mv.h_off[j + 1] - mv.h_off[j] : 0;
    if (cnt == 0 && ty == 0) { out[i] = s->column_ptr(snp); continue; }
    auto it = c->pc_map.find(4 * j + ty);
    if (it != c->pc_map.end()) {
      c->pc_use[it->second] = c->pc_seq;
      out[i] = c->pc_cols.p + (size_t)it->second * s->Wp;
    }
  }
  // ... then slots for the others, and the patch list
  for (int i = 0; i < count; ++i) {
    const int64_t snp = snps[i];
    if (!mv.covers(snp)) continue;
    const int ty = types ? types[i] : 0;
    const int64_t j = snp - mv.base, lo = mv.n_missing > 0 ? mv.h_off[j] : 0, cnt = mv.n_missing > 0 ? mv.h_off[j + 1] - lo : 0;
    if (cnt == 0 && ty == 0) continue;
    const bool fresh = host_vals != nullptr && host_vals[i] != nullptr;
    if (out[i] != nullptr && !fresh) continue;
    if (fresh)   // the typed copies of this SNP are stale now: they are rebuilt on their next use
      for (int t2 = 1; t2 < 4; ++t2) {
        auto ot = c->pc_map.find(4 * j + t2);
        if (ot != c->pc_map.end()) { c->pc_snp[ot->second] = -1; c->pc_map.erase(ot); }
      }
    uint32_t* dst;
    if (out[i] == nullptr) {
      auto again = c->pc_map.find(4 * j + ty);   // the same column twice in one request
      if (again != c->pc_map.end()) {
        out[i] = c->pc_cols.p + (size_t)again->second * s->Wp;
        if (!fresh) continue;
        dst = const_cast<uint32_t*>(out[i]);
      } else {
        const int slot = overlay_take_slot(c);
        c->pc_map[4 * j + ty] = slot;
        c->pc_snp[slot] = 4 * j + ty;
        c->pc_use[slot] = c->pc_seq;
        dst = c->pc_cols.p + (size_t)slot * s->Wp;
        out[i] = dst;
      }
    } else {
      dst = const_cast<uint32_t*>(out[i]);
    }
    PatchDesc d;
    d.src = s->column_ptr(snp);
    d.dst = dst;
    d.idx = cnt ? mv.idx + lo : nullptr;
    d.cnt = cnt;
    d.type = ty;
    if (fresh) {
      memcpy(c->pc_h_vals.p + staged, host_vals[i], (size_t)cnt);
      d.vals = c->pc_h_vals.p + staged;   // mapped pinned memory: read by the kernel over PCIe, no separate copy
      d.keep = c->miss_val.p + lo;
      staged += (size_t)cnt;
      g_h2d_bytes.fetch_add((uint64_t)cnt, std::memory_order_relaxed);
    } else {
      d.vals = cnt ? c->miss_val.p + lo : nullptr;
      d.keep = nullptr;
    }
    pend.push_back(d);
  }
  if (!pend.empty()) {
    BMG_CUDA(cudaSetDevice(s->device));
    overlay_launch(c, pend);
  }
  return (int)pend.size();
}

// DataModel::miss_val of several SNPs at once (a move's proposed additions, or the Gibbs step's in-model SNPs)
void chain_set_missing_many(Chain* c, const int64_t* snps, int count, const int8_t* const* vals)
{
  if (count == 0) return;
  BMG_REQUIRE(count <= 2048, "bmg_chain_set_missing: too many SNPs in one call");
  for (int i = 0; i < count; ++i) {
    BMG_REQUIRE(c->mv.covers(snps[i]), "bmg_chain_set_missing: SNP outside the chain's missing-call index");
    const int64_t cnt = c->mv.n_missing > 0 ? c->mv.count(snps[i]) : 0;
    BMG_REQUIRE(vals[i] != nullptr || cnt == 0, "bmg_chain_set_missing: null values");
    for (int64_t q = 0; q < cnt; ++q) BMG_REQUIRE(vals[i][q] >= 0 && vals[i][q] <= 2, "bmg_chain_set_missing: values must be 0, 1 or 2");
  }
  std::vector<const uint32_t*> ptrs(count);
  if (chain_overlay_columns(c, snps, count, ptrs.data(), vals) > 0)
    BMG_CUDA(cudaStreamSynchronize(c->stream));   // the staging buffer is free again, and the slots are complete for
                                                  // readers on other streams (the column-statistics server)
}

// The imputed values changed in bulk: drop every slot except those of keep[0..k) (SNPs whose values did not change)
void chain_overlay_invalidate(Chain* c, const int64_t* keep, int k)
{
  if (c->pc_slots == 0) return;
  std::vector<std::pair<int64_t, int>> kept;
  for (int l = 0; l < k; ++l) {
    if (!c->mv.covers(keep[l])) continue;
    for (int ty = 0; ty < 4; ++ty) {
      auto it = c->pc_map.find(4 * (keep[l] - c->mv.base) + ty);
      if (it != c->pc_map.end()) kept.push_back(*it);
    }
  }
  c->pc_map.clear();
  std::fill(c->pc_snp.begin(), c->pc_snp.end(), (int64_t)-1);
  for (const auto& e : kept) { c->pc_map[e.first] = e.second; c->pc_snp[e.second] = e.first; }
}

}  // namespace bmg
