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

// ---- launch helper ------------------------------------------------------------
// pdl_active(): inside a PdlScope (bnf_map_steps' step sequence) and BNF_PDL != 0.  Only kernels
// that begin with pdl_wait() may be launched through launch_k with pdl = true.
bool pdl_active();
struct PdlScope {
  bool prev;
  explicit PdlScope(bool on);
  ~PdlScope();
};

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                            Args&&... args) {
  cudaLaunchConfig_t cfg;
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr;
  cfg.numAttrs = 0;
  if (pdl_active()) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 1;
  }
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

#define BNF_PROF_CAT2(a, b) a##b
#define BNF_PROF_CAT(a, b) BNF_PROF_CAT2(a, b)
#define BNF_PROF(name, st) ::bnf::ProfScope BNF_PROF_CAT(prof_scope_, __LINE__)(name, st)

}  // namespace bnf
