// tcgen05 (bf16 operands, f32 TMEM accumulators) GEMMs of the dense stack.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "bnf_model.h"

namespace bnf {

// NULL when the plan can run on the tensor-core path, else a reason string.
const char* tc_unsupported_reason(const DevModel& m);
const char* tc_last_error();
// bf16 elements of staged weights per network (all hidden layers, padded K)
size_t tc_weight_elems(const DevModel& m);
// f32 master kernels -> bf16 staging: wt = [layer][N=W][Kp] (K-major for the
// forward GEMM), wn = [layer][Kp][W] (natural (in,out); K-major for dgrad).
void tc_cast_weights(const DevModel& m, const float* params, __nv_bfloat16* wt, __nv_bfloat16* wn,
                     int n_net, cudaStream_t st);
// B operand: `wn` read MN-major (default) or `wt` read K-major (BNF_FWD_WT=1 / wn == NULL)
int tc_fwd_layer(const bnf_plan* p, int layer, const float* params, const float* derived,
                 const __nv_bfloat16* a_in, const __nv_bfloat16* wt, const __nv_bfloat16* wn,
                 __nv_bfloat16* z, __nv_bfloat16* h, int n_net, int B, cudaStream_t st,
                 bool x3 = false, float* zf = nullptr);
bool tc_fwd_uses_wt();
// ---- bf16x3 (split-operand, f32-parity) mode: every activation / weight tensor fed to a GEMM
// holds three bf16 planes side by side (a = a0 + a1 + a2); z stays f32.  wn3 = [layer][Kp][3*W].
void tc_cast_weights_x3(const DevModel& m, const float* params, __nv_bfloat16* wn3, int n_net, cudaStream_t st);
int tc_dgrad_act_x3(const bnf_plan* p, int layer, const __nv_bfloat16* wn3, const __nv_bfloat16* dU,
                    __nv_bfloat16* out, const float* z_prev, const float* params, const float* derived,
                    float* grad, int n_net, int B, cudaStream_t st);
// training step, last hidden layer: fwd GEMM + head + log-likelihood + activation backward
bool tc_fwd_head_supported(const DevModel& m);
int tc_fwd_head(const bnf_plan* p, const float* params, const float* derived, const __nv_bfloat16* a_in,
                const __nv_bfloat16* wn, const float* y, const int32_t* idx, int64_t idx_stride,
                __nv_bfloat16* dU, float* ll, float* grad, int n_net, int B, cudaStream_t st, bool x3 = false);
// fused feature encode + Dense_0 (the default first layer of the bf16 path when Fp <= 128)
bool tc_fused_encode_supported(const DevModel& m);
bool tc_fused_encode_wanted(const DevModel& m, bool training);
int tc_fwd_layer0_fused(const bnf_plan* p, const float* params, const float* derived, const float* x,
                        const int32_t* idx, int64_t idx_stride, const __nv_bfloat16* wn,
                        __nv_bfloat16* feat, __nv_bfloat16* z, __nv_bfloat16* h, int n_net, int B,
                        cudaStream_t st);
// out_bf (hidden layers) or out_f32 (layer 0 -> dfeat [n_net,B,Fp]) receives isf * dU @ K^T
// z_prev != NULL (hidden layers): fuse the activation backward of layer-1 into the epilogue
// (out_bf receives dU of layer-1; bias / activation-mix / layer-scale grads go to `grad`).
int tc_dgrad(const bnf_plan* p, int layer, const __nv_bfloat16* wn, const __nv_bfloat16* dU,
             __nv_bfloat16* out_bf, float* out_f32, int n_net, int B, cudaStream_t st,
             const __nv_bfloat16* z_prev = nullptr, const float* params = nullptr,
             const float* derived = nullptr, float* grad = nullptr);
// the fused dgrad + activation backward keeps a layer's bias column sums in shared memory
bool tc_dgrad_act_supported(const DevModel& m);
// layer-0 dgrad with the feature-encode backward fused into its epilogue (dfeat stays on chip)
bool tc_dgrad0_enc_supported(const DevModel& m);
int tc_dgrad0_enc(const bnf_plan* p, const __nv_bfloat16* wn, const __nv_bfloat16* dU, const float* x,
                  const int32_t* idx, int64_t idx_stride, const float* params, const float* derived,
                  float* grad, int n_net, int B, cudaStream_t st, bool x3 = false);
// bias0: (layer 0) feat carries the constant-one column -> also emit the Dense_0 bias gradient
int tc_wgrad(const bnf_plan* p, int layer, const __nv_bfloat16* a_in, const __nv_bfloat16* dU,
             float* grad, int n_net, int B, cudaStream_t st, bool x3 = false, bool bias0 = false);
bool tc_bias0_via_wgrad(const DevModel& m);

int tc_debug_gemm(int mn_major, const __nv_bfloat16* A, const __nv_bfloat16* Bm, float* C, int n_net,
                  int M, int N, int K, int sm_count, cudaStream_t st);

}  // namespace bnf
