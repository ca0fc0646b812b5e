// sampler_capi.cpp -- extern "C" entry points of the host sampler (placeholder until sampler.cpp lands)
#include "../common.cuh"
#include "../../../include/bmagwa_b200.h"
#define BMG_API extern "C" __attribute__((visibility("default")))
static int nyi() { bmg::set_last_error("host sampler not built into this library yet"); return 1; }
BMG_API int bmg_sampler_create(const char*, int, int, bmg_sampler**) { return nyi(); }
BMG_API int bmg_sampler_create_on_store(const char*, int, bmg_store*, bmg_sampler**) { return nyi(); }
BMG_API int bmg_sampler_set_option(bmg_sampler*, const char*, const char*) { return nyi(); }
BMG_API int bmg_sampler_begin(bmg_sampler*) { return nyi(); }
BMG_API int bmg_sampler_run(bmg_sampler*, int64_t) { return nyi(); }
BMG_API int bmg_sampler_end(bmg_sampler*) { return nyi(); }
BMG_API int bmg_sampler_stats(bmg_sampler*, double*) { return nyi(); }
BMG_API bmg_store* bmg_sampler_store(bmg_sampler*) { return nullptr; }
BMG_API bmg_chain* bmg_sampler_chain(bmg_sampler*) { return nullptr; }
BMG_API int bmg_sampler_destroy(bmg_sampler*) { return 0; }
