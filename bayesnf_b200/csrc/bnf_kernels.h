// Launch wrappers implemented in bnf_kernels.cu (SIMT) and bnf_tc.cu (tcgen05).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "bnf_model.h"

namespace bnf {

// derived scalars of every network; optionally also zeroes two n_net-float accumulators and
// two int32 cursors (the MAP prologue)
void launch_prep(const DevModel& m, const float* params, float* derived, int n_net, float* zero_acc,
                 float* zero_acc2, int32_t* zero_cursors, cudaStream_t st, float** loss_slot = nullptr,
                 float* out_loss = nullptr, float* zero_rows = nullptr);

template <typename T>
void launch_encode(const DevModel& m, const float* derived, const float* x, const int32_t* idx,
                   int64_t idx_stride, int B, T* feat, int n_net, cudaStream_t st);
void launch_encode_bwd(const DevModel& m, const float* params, const float* derived, const float* x,
                       const int32_t* idx, int64_t idx_stride, int B, const float* dfeat, float* grad,
                       int n_net, bool fast_trig, bool col_major, cudaStream_t st);
template <typename T>
void launch_head(const DevModel& m, const float* params, const float* derived, const T* h,
                 const float* y, const int32_t* idx, int64_t idx_stride, int B, float* out_loc,
                 float* opre, float* r, float* ll, float* grad, int n_net, cudaStream_t st);
template <typename T>
bool launch_head_fused(const DevModel& m, const float* params, const float* derived, const T* h, const T* z,
                       const float* y, const int32_t* idx, int64_t idx_stride, int B, T* dU, float* ll,
                       float* grad, int n_net, cudaStream_t st);
template <typename T>
void launch_act_bwd(const DevModel& m, int layer, bool is_head, const float* params,
                    const float* derived, const T* z, const T* h, const float* r, T* dU, int B,
                    float* grad, int n_net, cudaStream_t st);

// SIMT GEMMs; T = float (parity mode) or bf16 storage (debug mode 2)
template <typename T>
void launch_fwd_layer_simt_t(const DevModel& m, int layer, const float* params, const float* derived,
                             const T* a_in, int K, int lda, T* z, T* h, int n_net, int B, cudaStream_t st);
template <typename T, typename TO>
void launch_dgrad_simt_t(const DevModel& m, int layer, const float* params, const T* dU, TO* out,
                         int Kout, int ld_out, int n_net, int B, cudaStream_t st);
template <typename T>
void launch_wgrad_simt_t(const DevModel& m, int layer, const T* a_in, int Kin, int lda, const T* dU,
                         float* grad, int n_net, int B, cudaStream_t st);

void launch_tick(int32_t* step_count, int32_t* slot, cudaStream_t st);
void launch_arm_loss(float** loss_slot, float* out_loss, int32_t* cursor, cudaStream_t st);
void launch_map_adam(int P, float* params, float* am, float* av, const float* g_ll,
                     const int32_t* step_count, float c_ll, float prior_weight, float lr,
                     float* prior_out, int n_net, cudaStream_t st);
// fused tail of a MAP step (adam + grad zero + bf16 restage + loss row + ticks + next derived)
void launch_map_update(const DevModel& m, float* params, float* am, float* av, float* grad,
                       int32_t* step_count, float c_ll, float prior_weight, float lr, float* prior,
                       float* ll, float* const* loss_slot, int32_t* slot, unsigned int* counter, float* derived,
                       __nv_bfloat16* wn, size_t w_per_net, int wn_planes, int n_net, cudaStream_t st);
// ---- bf16x3 (split-operand) mode: activations fed to a GEMM are rows of three bf16 planes
void launch_encode_x3(const DevModel& m, const float* derived, const float* x, const int32_t* idx,
                      int64_t idx_stride, int B, __nv_bfloat16* feat3, int n_net, cudaStream_t st);
void launch_head_x3(const DevModel& m, const float* params, const float* derived, const __nv_bfloat16* h3,
                    const float* y, const int32_t* idx, int64_t idx_stride, int B, float* out_loc,
                    float* opre, float* r, float* ll, float* grad, int n_net, cudaStream_t st);
bool head_fused_x3_supported(const DevModel& m);
bool launch_head_fused_x3(const DevModel& m, const float* params, const float* derived, const __nv_bfloat16* h3,
                          const float* z, const float* y, const int32_t* idx, int64_t idx_stride, int B,
                          __nv_bfloat16* dU3, float* ll, float* grad, int n_net, cudaStream_t st);
void launch_map_loss(int n_net, const float* ll, const float* prior, float c_ll, float prior_weight,
                     float* out, const int32_t* slot, cudaStream_t st);
void launch_vi_sample(int P, int E, int S, const float* mu, const float* rho, const float* eps_in,
                      float* eps_out, uint64_t seed, uint64_t stream_id, const int32_t* step_ptr, float* z,
                      cudaStream_t st);
// device-drawn minibatch windows: idx_win [n_rows, B] <- permutation of (member, epoch) at the
// window given by the device-side count of completed steps
void launch_batch_window(int n_total, int B, int steps_per_epoch, int n_rows, uint64_t seed, int64_t first_member,
                         const int32_t* step_count, int32_t* idx_win, cudaStream_t st);
void launch_vi_adam(int P, int E, int S, float* mu, float* rho, float* am, float* av, const float* z,
                    const float* eps, const float* g_ll, const int32_t* step_count, float c, float lr,
                    float* loss_acc, cudaStream_t st);
void launch_vi_loss(int E, int S, const float* loss_acc, const float* ll, float c, float* out,
                    float* const* out_slot, const int32_t* slot, cudaStream_t st);
void launch_init_params(const DevModel& m, float lns_init, uint64_t seed, int64_t first_member,
                        int n_net, float* params, cudaStream_t st);
void launch_quantiles(const float* means, const float* scales, int M, int N, const double* q, int nq,
                      bool approximate, const float* ndtri_q, float* out, float* mm, cudaStream_t st);

void launch_nb_quantiles(const float* loc, const float* shape_raw, const float* pi_logit, int M, int N,
                         const double* q, int nq, float* means, float* out, float* ws, cudaStream_t st);

}  // namespace bnf
