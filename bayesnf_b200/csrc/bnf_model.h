// Internal model description shared by host bookkeeping and device kernels.
// Follows models.py:197-273 (BayesianNeuralField1D) of the reference.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/bnf.h"

namespace bnf {

constexpr int kMaxD = 16;       // input dims
constexpr int kMaxSeasonal = 64;
constexpr int kMaxInter = 32;
constexpr int kMaxLayers = 16;  // hidden layers
constexpr int kFeatPad = 64;
constexpr int kDerivedStride = 64;  // floats of derived scalars per network    // K-block of the tensor-core path (128B of bf16)

// Passed by value as a __grid_constant__ kernel parameter (< 4 KB).
struct DevModel {
  int D, F, Fp, W, L, P;
  int likelihood;
  int n_seasonal, n_inter;
  // feature column bases (column index inside the F-wide feature row) and the
  // parameter offset of each group's feature_inv_sp_scale{i} (-1: group absent)
  int col_x, off_scale_x;
  int fourier_deg[kMaxD], fourier_col[kMaxD], fourier_scale_off[kMaxD];
  int col_seasonal, off_scale_seasonal;
  int col_inter, off_scale_inter;
  float input_scales[kMaxD];          // f32(input_scales), models.py:221
  float seasonal_w[kMaxSeasonal];     // f32(2*pi) * f32(freq), models.py:73
  float seasonal_h[kMaxSeasonal];     // harmonic index (denominator)
  int inter_a[kMaxInter], inter_b[kMaxInter];
  // parameter offsets
  int off_lsa, off_actw, off_out_scale;
  int off_layer_scale[kMaxLayers];
  int off_bias[kMaxLayers + 1], off_kernel[kMaxLayers + 1];
  float inv_sqrt_F, inv_sqrt_W;       // 1/sqrt(fan_in) (models.py:267,272)
};

struct Leaf {
  std::string name;
  int64_t offset;
  int rows, cols;  // cols==0: 1-D of length rows; rows==0: scalar
};

}  // namespace bnf

struct bnf_plan {
  bnf::DevModel m;
  std::vector<bnf::Leaf> leaves;
  int n_groups;
  int sm_count;
  // CUDA-graph replay of one full-batch MAP step (see bnf_map_steps); mutable
  // cache, so graph mode is single-threaded per plan.
  // Cache of kGraphWays captured steps per kind (kind 0: the MAP/MLE step, 1: the VI step; entry =
  // kind * kGraphWays + way): a caller that alternates between two input buffers (double-buffered
  // host->device prefetch) replays both without re-capturing.  graph_age orders the ways (LRU).
  static constexpr int kGraphWays = 2;
  mutable void* graph_stream = nullptr;   // capture stream (nothing executes on it)
  mutable void* graph_exec[2 * kGraphWays] = {};
  mutable void* graph_key[2 * kGraphWays] = {};     // StepGraphKey of graph_exec
  // the same step captured kGraphUnroll times back to back (one launch = several steps: the
  // programmatic-dependent-launch chain then also spans the step boundary); built on demand
  mutable void* graph_exec_multi[2 * kGraphWays] = {};
  mutable int graph_multi_steps[2 * kGraphWays] = {};
  mutable void* last_key[2] = {nullptr, nullptr};   // StepGraphKey of the previous replayable call of the kind
  mutable long long graph_launches[2 * kGraphWays] = {};   // kernel nodes per replay
  mutable unsigned long long graph_age[2 * kGraphWays] = {};
  mutable unsigned long long graph_clock = 0;
};
