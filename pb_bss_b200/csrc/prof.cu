#include <atomic>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "prof.cuh"

namespace pbb {

static std::atomic<long long> g_launches{0};
static std::atomic<int> g_enabled{0};
static std::mutex g_mu;
struct Rec { std::string name; cudaEvent_t a, b; };
static std::vector<Rec> g_recs;
static thread_local int g_open = -1;

void prof_begin(const char* name, cudaStream_t st) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (!g_enabled.load(std::memory_order_relaxed)) return;
  Rec r;
  r.name = name;
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
  cudaEventRecord(r.a, st);
  std::lock_guard<std::mutex> lk(g_mu);
  g_recs.push_back(r);
  g_open = (int)g_recs.size() - 1;
}

void prof_end(cudaStream_t st) {
  if (g_open < 0) return;
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_open < (int)g_recs.size()) cudaEventRecord(g_recs[g_open].b, st);
  g_open = -1;
}

static void clear_locked() {
  for (auto& r : g_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  g_recs.clear();
}

}  // namespace pbb

using namespace pbb;

extern "C" {

long long pbb_launch_count(void) { return g_launches.load(); }

void pbb_profile_enable(int on) { g_enabled.store(on ? 1 : 0); }

void pbb_profile_reset(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  clear_locked();
}

// Sums the recorded launches per kernel name and returns the one with the
// largest total device time (ms) and its launch count; clears the records.
int pbb_profile_dominant(char* name, int name_len, double* total_ms, int* launches) {
  std::lock_guard<std::mutex> lk(g_mu);
  std::map<std::string, std::pair<double, int>> acc;
  for (auto& r : g_recs) {
    if (cudaEventSynchronize(r.b) != cudaSuccess) continue;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) continue;
    auto& e = acc[r.name];
    e.first += ms;
    e.second += 1;
  }
  std::string best;
  double best_ms = -1.0;
  int best_n = 0;
  for (auto& kv : acc)
    if (kv.second.first > best_ms) { best = kv.first; best_ms = kv.second.first; best_n = kv.second.second; }
  if (name && name_len > 0) { strncpy(name, best.c_str(), name_len - 1); name[name_len - 1] = 0; }
  if (total_ms) *total_ms = best_ms < 0 ? 0.0 : best_ms;
  if (launches) *launches = best_n;
  clear_locked();
  return (int)acc.size();
}

// Prints every recorded launch (name, device ms) to stderr in launch order; keeps the records.
void pbb_profile_dump(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& r : g_recs) {
    float ms = 0.f;
    if (cudaEventSynchronize(r.b) != cudaSuccess) continue;
    if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) continue;
    fprintf(stderr, "[pbb] %-32s %9.4f ms\n", r.name.c_str(), (double)ms);
  }
}

}  // extern "C"
