// Library-wide pieces of the C ABI: version, thread-local error message, device attribute cache.
#include <stdlib.h>
#include <atomic>
#include <mutex>
#include <string>
#include <string.h>
#include <vector>

#include "common.cuh"

namespace nsc {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    if (const char* e = getenv("NSC_SMS")) {   // experiments: persistent grids of fewer CTAs (two streams side by side)
      const int cap = atoi(e);
      if (cap >= 1 && cap < n) n = cap;
    }
    cached[dev] = n;   // immutable once written; a benign race writes the same value
  }
  return cached[dev];
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- profiling records ---------------------------------------------------------------------------
struct ProfRec {
  char name[32];
  double flops, bytes;
  cudaEvent_t e0, e1;
};
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static int g_prof_cap = 0;
static std::vector<ProfRec> g_prof;

ProfScope::ProfScope(cudaStream_t s, const char* name, double flops, double bytes) : slot(-1), st(s) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_prof_on || (int)g_prof.size() >= g_prof_cap) return;
  ProfRec r;
  strncpy(r.name, name, sizeof(r.name) - 1);
  r.name[sizeof(r.name) - 1] = 0;
  r.flops = flops;
  r.bytes = bytes;
  if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
  cudaEventRecord(r.e0, st);
  g_prof.push_back(r);
  slot = (int)g_prof.size() - 1;
}

ProfScope::~ProfScope() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (slot < (int)g_prof.size()) cudaEventRecord(g_prof[slot].e1, st);
}

}  // namespace nsc

extern "C" {

long long nsc_launch_count(void) { return nsc::g_launches.load(); }

int nsc_profile_begin(int32_t max_records) {
  std::lock_guard<std::mutex> lk(nsc::g_prof_mu);
  for (auto& r : nsc::g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  nsc::g_prof.clear();
  nsc::g_prof_cap = max_records;
  nsc::g_prof_on = max_records > 0;
  return NSC_OK;
}

int nsc_profile_end(int32_t* n_records, char* names, float* ms, double* flops, double* bytes, int32_t cap) {
  std::lock_guard<std::mutex> lk(nsc::g_prof_mu);
  nsc::g_prof_on = false;
  int n = 0;
  for (auto& r : nsc::g_prof) {
    if (n < cap) {
      float t = 0.f;
      if (cudaEventSynchronize(r.e1) != cudaSuccess || cudaEventElapsedTime(&t, r.e0, r.e1) != cudaSuccess) t = -1.f;
      if (names) memcpy(names + (size_t)n * 32, r.name, 32);
      if (ms) ms[n] = t;
      if (flops) flops[n] = r.flops;
      if (bytes) bytes[n] = r.bytes;
      ++n;
    }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  nsc::g_prof.clear();
  if (n_records) *n_records = n;
  return NSC_OK;
}

int nsc_version(void) { return 100; /* 0.1.0 */ }

const char* nsc_last_error(void) { return nsc::g_err; }

}  // extern "C"
