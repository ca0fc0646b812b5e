// group.cuh -- several chains over one SNP-sharded store (group.cu)
#pragma once
#include <stdint.h>
#include "store.cuh"

namespace bmg {

constexpr int kGroupMaxRanks = 16;
struct Group;

Group* group_create(Store* shard, int world, int rank, int n_chains, int64_t stride, const char* shm_name);
void group_destroy(Group* g);
void group_barrier(Group* g);
// bmg_allgather_fn over the group's peer mappings (ctx = Group*)
int group_allgather(void* ctx, void* dev_buffer, int64_t elems_per_rank, int elem_bytes, void* cuda_stream);
const double* group_scan_round(Group* g, Chain* mine);
void group_serve(Group* g, int64_t n_rounds);
int group_world(const Group* g);
int group_rank(const Group* g);
int group_chains(const Group* g);
int64_t group_stride(const Group* g);
const int32_t* group_n1(const Group* g);
const int32_t* group_n2(const Group* g);
const GlobalMissing* group_missing(const Group* g);
Chain* group_scan_chain(Group* g);
void group_stats(const Group* g, double* out4);

}  // namespace bmg
