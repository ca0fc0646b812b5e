// main.cpp -- the `bmagwa` command line: ./bmagwa config.ini
//
// Same contract as the reference binary (src/main.cpp:36-120): one INI file, n_threads independent
// chains (one host thread per chain) over ONE shared device-resident genotype store, output files
// named basename<chain>_*.  Talks to the library only through the C ABI of include/bmagwa_b200.h.
#include <pthread.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "bmagwa_b200.h"

struct Job {
  bmg_sampler* sampler;
  long n_iter;
  int status;
  std::string error;
};

static void* run_chain(void* arg)
{
  Job* j = static_cast<Job*>(arg);
  j->status = bmg_sampler_begin(j->sampler);
  if (!j->status) j->status = bmg_sampler_run(j->sampler, j->n_iter);
  if (!j->status) j->status = bmg_sampler_end(j->sampler);
  if (j->status) j->error = bmg_last_error();
  return NULL;
}

// n_threads / do_n_iter as the library reads them (same parser as the samplers: bmg_ini_lookup)
static long ini_long(const char* path, const char* section, const char* key, const char* dflt)
{
  char buf[256];
  if (bmg_ini_lookup(path, section, key, dflt, buf, (int)sizeof buf)) {
    fprintf(stderr, "terminate called after throwing an instance of 'std::runtime_error'\n  what():  %s\n", bmg_last_error());
    exit(134);
  }
  char* end = NULL;
  const long v = strtol(buf, &end, 10);
  if (end == buf) {
    fprintf(stderr, "terminate called after throwing an instance of 'std::runtime_error'\n  what():  Config error: invalid value for %s.%s\n", section, key);
    exit(134);
  }
  return v;
}

int main(int argc, char* argv[])
{
  printf("-------------------------------------------------------------\n"
         "BMAGWA hot path, B200-native implementation (reference: BMAGWA software version 2.0)\n"
         "-------------------------------------------------------------\n\n");
  if (argc != 2) {
    printf("Usage: %s INIFILE\n", argv[0]);
    return 0;
  }
  const char* ini = argv[1];
  const long n_threads = ini_long(ini, "thread", "n_threads", "1");
  const long do_n_iter = ini_long(ini, "sampler", "do_n_iter", "0");
  std::vector<Job> jobs(n_threads > 0 ? (size_t)n_threads : 1);
  bmg_sampler* first = NULL;
  for (size_t t = 0; t < jobs.size(); ++t) {
    // chain 0 loads the data and builds the device store (the reference's Data + PrecomputedSNPCovariances,
    // src/main.cpp:54-68); its summaries are printed where the reference prints them
    int rc = t == 0 ? bmg_sampler_create(ini, 0, -1, &jobs[t].sampler)
                    : bmg_sampler_create_on_store(ini, (int)t, bmg_sampler_store(first), &jobs[t].sampler);
    if (rc) {
      fprintf(stderr, "terminate called after throwing an instance of 'std::runtime_error'\n  what():  %s\n", bmg_last_error());
      return 134;
    }
    if (t == 0) {
      first = jobs[t].sampler;
      double sm[6];
      if (bmg_store_summaries(bmg_sampler_store(first), sm) == 0)   // src/main.cpp:60-62 (iostream default precision: %g)
        printf("var y = %g\nvar x = %g\nmean x = %g\nPrecomputing SNP covariances\n", sm[4], sm[2] / sm[3], sm[0] / sm[1]);
    }
    printf("Initializing sampler %zu\n", t);
    jobs[t].n_iter = do_n_iter;
    jobs[t].status = 0;
  }
  std::vector<pthread_t> threads(jobs.size());
  for (size_t t = 0; t < jobs.size(); ++t) {
    printf("Creating thread %zu\n", t);
    if (pthread_create(&threads[t], NULL, &run_chain, &jobs[t])) {
      fprintf(stderr, "Creating thread failed\n");
      return 1;
    }
  }
  int rc = 0;
  for (size_t t = 0; t < jobs.size(); ++t) {
    pthread_join(threads[t], NULL);
    printf("Completed join with thread %zu having a status of %d\n", t, jobs[t].status);
    if (jobs[t].status) { fprintf(stderr, "chain %zu: %s\n", t, jobs[t].error.c_str()); rc = 1; }
  }
  for (size_t t = jobs.size(); t-- > 1;) bmg_sampler_destroy(jobs[t].sampler);
  bmg_sampler_destroy(jobs[0].sampler);
  return rc;
}
