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

// minimal lookup of "key = value" in a section of the INI file (only for n_threads / do_n_iter here;
// the library parses the file in full)
static std::string ini_lookup(const char* path, const char* section, const char* key, const char* dflt)
{
  FILE* f = fopen(path, "r");
  if (!f) return dflt;
  char line[256];
  std::string cur, out = dflt;
  while (fgets(line, sizeof line, f)) {
    char* s = line;
    while (*s == ' ' || *s == '\t') ++s;
    if (*s == '[') { char* e = strchr(s, ']'); if (e) cur.assign(s + 1, e - s - 1); continue; }
    if (*s == ';' || *s == '#' || cur != section) continue;
    char* eq = strchr(s, '=');
    if (!eq) continue;
    std::string name(s, eq - s);
    while (!name.empty() && (name.back() == ' ' || name.back() == '\t')) name.pop_back();
    if (name != key) continue;
    std::string v(eq + 1);
    size_t c = v.find(" ;");
    if (c != std::string::npos) v.resize(c);
    size_t a = v.find_first_not_of(" \t\r\n"), b = v.find_last_not_of(" \t\r\n");
    out = a == std::string::npos ? "" : v.substr(a, b - a + 1);
  }
  fclose(f);
  return out;
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
  const int n_threads = atoi(ini_lookup(ini, "thread", "n_threads", "1").c_str());
  const long do_n_iter = atol(ini_lookup(ini, "sampler", "do_n_iter", "0").c_str());
  std::vector<Job> jobs(n_threads > 0 ? n_threads : 1);
  bmg_sampler* first = NULL;
  for (size_t t = 0; t < jobs.size(); ++t) {
    printf("Initializing sampler %zu\n", t);
    int rc = t == 0 ? bmg_sampler_create(ini, 0, -1, &jobs[t].sampler)
                    : bmg_sampler_create_on_store(ini, (int)t, bmg_sampler_store(first), &jobs[t].sampler);
    if (rc) {
      fprintf(stderr, "terminate called after throwing an instance of 'std::runtime_error'\n  what():  %s\n", bmg_last_error());
      return 134;
    }
    if (t == 0) first = jobs[t].sampler;
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
