#include "bnf_prof.h"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/bnf.h"

namespace bnf {

static std::atomic<unsigned long long> g_launch_count{0};
static std::atomic<bool> g_prof_on{false};
static std::mutex g_mu;
struct Rec { const char* name; cudaEvent_t a, b; };
static std::vector<Rec> g_recs;

void prof_count() { g_launch_count.fetch_add(1, std::memory_order_relaxed); }
void prof_add_launches(long long n) { g_launch_count.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
bool prof_enabled() { return g_prof_on.load(std::memory_order_relaxed); }

void prof_begin(const char* name, cudaStream_t st, int* slot) {
  std::lock_guard<std::mutex> lk(g_mu);
  Rec r;
  r.name = name;
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
  cudaEventRecord(r.a, st);
  g_recs.push_back(r);
  *slot = (int)g_recs.size() - 1;
}
void prof_end(cudaStream_t st, int slot) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (slot >= 0 && slot < (int)g_recs.size()) cudaEventRecord(g_recs[slot].b, st);
}

// PDL scope (see bnf_prof.h)
static thread_local bool g_pdl = false;
static bool pdl_env() {
  const char* e = getenv("BNF_PDL");      // read per scope: tests flip it
  return !(e && e[0] == '0');
}
bool pdl_active() { return g_pdl && !prof_enabled(); }
PdlScope::PdlScope(bool on) : prev(g_pdl) { g_pdl = on && pdl_env(); }
PdlScope::~PdlScope() { g_pdl = prev; }

}  // namespace bnf

using namespace bnf;

extern "C" uint64_t bnf_debug_launch_count(void) { return (uint64_t)g_launch_count.load(); }

extern "C" int bnf_debug_profile(int32_t enable) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& r : g_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  g_recs.clear();
  g_prof_on.store(enable != 0);
  return BNF_OK;
}

// "name count total_ms\n" per kernel class; synchronises the device.
extern "C" int bnf_debug_profile_report(char* buf, int32_t len) {
  if (!buf || len < 1) return BNF_ERR_INVALID;
  cudaDeviceSynchronize();
  std::lock_guard<std::mutex> lk(g_mu);
  std::map<std::string, std::pair<long long, double>> agg;
  for (auto& r : g_recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      auto& e = agg[r.name];
      e.first += 1;
      e.second += ms;
    }
  }
  std::string out;
  char line[256];
  for (auto& kv : agg) {
    snprintf(line, sizeof(line), "%s %lld %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
    out += line;
  }
  snprintf(buf, len, "%s", out.c_str());
  return BNF_OK;
}
