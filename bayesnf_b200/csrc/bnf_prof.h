// Launch accounting + optional per-kernel CUDA-event timing (bnf_debug_profile).
#pragma once
#include <cuda_runtime.h>

namespace bnf {

void prof_count();
void prof_add_launches(long long n);
bool prof_enabled();
void prof_begin(const char* name, cudaStream_t st, int* slot);
void prof_end(cudaStream_t st, int slot);

// Counts the launch; when profiling is on also brackets it with CUDA events
// recorded on the launching stream.
struct ProfScope {
  cudaStream_t st; int slot;
  ProfScope(const char* name, cudaStream_t s) : st(s), slot(-1) {
    prof_count();
    if (prof_enabled()) prof_begin(name, st, &slot);
  }
  ~ProfScope() { if (slot >= 0) prof_end(st, slot); }
};

#define BNF_PROF_CAT2(a, b) a##b
#define BNF_PROF_CAT(a, b) BNF_PROF_CAT2(a, b)
#define BNF_PROF(name, st) ::bnf::ProfScope BNF_PROF_CAT(prof_scope_, __LINE__)(name, st)

}  // namespace bnf
