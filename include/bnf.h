/*
 * bnf.h -- C ABI of libbnf_sm100.so, the B200 (sm_100a) implementation of the
 * BayesNF ensemble-training hot path.
 *
 * The reference (google/bayesnf) has no FFI: its seam is three Python functions
 * taking plain arrays:
 *     inference.fit_map      src/bayesnf/inference.py:376-458
 *     inference.fit_vi       src/bayesnf/inference.py:336-373
 *     inference.predict_bnf  src/bayesnf/inference.py:461-507
 * Every entry point below names the reference lines it replaces.  The Python
 * host layer (bayesnf_b200/inference.py) binds these with ctypes; see
 * INTEGRATION.md for the stub a reference maintainer would add.
 *
 * Conventions
 *   - All data pointers are DEVICE pointers owned by the caller unless the
 *     comment says "host".  No torch types cross this boundary.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it and
 *     nothing synchronises unless stated.
 *   - Return value: 0 on success, non-zero error code; bnf_last_error() returns
 *     a thread-local message.  There is NO CPU fallback: on a machine without
 *     an sm_100 device every compute entry point fails with BNF_ERR_CUDA.
 *   - A plan is immutable after creation and may be shared by streams; the
 *     workspace passed to a call must not be used concurrently by another call.
 *   - "network" = one parameter vector pushed through the MLP.  MAP/MLE: one
 *     network per ensemble member.  VI: S Monte-Carlo draws per member.
 *
 * Parameter vector (one network, P floats, f32), in the order of the
 * reference's params tuple (models.py:94-103, jax.tree_util.tree_leaves of the
 * Flax dict = sorted keys):
 *     [0] log_noise_scale  [1] shape  [2] inflated_loc_probs (logit)
 *     then Dense_0/bias (W), Dense_0/kernel (F x W, row-major (in,out)), ...,
 *     Dense_L/bias (1), Dense_L/kernel (W x 1), feature_inv_sp_scale{i}...,
 *     inv_sp_layer_scale{l}..., inv_sp_output_scale, log_scale_adjustment (D),
 *     logit_activation_weight        -- names sorted as strings.
 * bnf_plan_leaf() reports name/offset/shape of every leaf.
 */
#ifndef BNF_H_
#define BNF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BNF_ABI_VERSION 1

/* models.py:30-33 LikelihoodDist */
enum { BNF_NORMAL = 0, BNF_NB = 1, BNF_ZINB = 2 };

/* Arithmetic mode of the dense stack.
 *   BNF_PREC_FP32 : f32 SIMT FMA, f32 activations.  The <=1e-5 parity mode.
 *   BNF_PREC_BF16 : tcgen05 kind::f16 (bf16 operands, f32 TMEM accumulators),
 *                   bf16 activations in HBM, f32 master weights/grads/Adam.
 *   BNF_PREC_BF16X3 : see below -- the tensor-core mode at the parity tolerance. */
enum {
  BNF_PREC_FP32 = 0,
  BNF_PREC_BF16 = 1,
  BNF_PREC_BF16_SIMT = 2, /* debug: bf16 storage, SIMT f32 FMA GEMMs (no tensor cores) */
  /* f32-class arithmetic ON the tensor cores: every forward GEMM operand a is carried as three bf16
   * planes a0 + a1 + a2 (|a - a0 - a1 - a2| <= 2^-27 |a|) and every f32 GEMM runs as the six
   * tcgen05 kind::f16 products a_i.b_j, i + j <= 2, into one f32 TMEM accumulator (smallest
   * products first: the accumulation rounds toward zero); the backpropagated dU carries two planes
   * (five products); pre-activations stay f32 in HBM; activation math with <= 3e-7 absolute
   * error.  Meets the 1e-5 parity bar (tests/test_gpu_parity.py) -- the default precision of the
   * Python layer.                                                                            */
  BNF_PREC_BF16X3 = 3
};

enum {
  BNF_OK = 0,
  BNF_ERR_INVALID = 1,   /* bad argument (maps to ValueError on the host side) */
  BNF_ERR_CUDA = 2,      /* CUDA runtime/driver failure, or no sm_100 device   */
  BNF_ERR_WORKSPACE = 3, /* workspace too small                                */
  BNF_ERR_UNSUPPORTED = 4
};

typedef struct bnf_plan bnf_plan_t;

/* Static description of the model = the reference's `model_args`
 * (spatiotemporal.py:360-370) after host-side bookkeeping.  All pointers are
 * HOST pointers and are copied.                                              */
typedef struct bnf_config {
  int32_t abi_version;         /* = BNF_ABI_VERSION                           */
  int32_t input_dim;           /* D                                           */
  int32_t width;               /* W  (models.py:200)                          */
  int32_t depth;               /* number of hidden layers (models.py:201)     */
  int32_t likelihood;          /* BNF_NORMAL / BNF_NB / BNF_ZINB              */
  int32_t n_seasonal;          /* unique seasonal frequencies (models.py:36-59)*/
  const float* seasonal_freq;  /* [n_seasonal] f32 frequencies h/p            */
  const float* seasonal_harm;  /* [n_seasonal] harmonic index (denominator)   */
  const int32_t* fourier_degrees; /* [D] (spatiotemporal.py:296-310)          */
  int32_t n_interactions;
  const int32_t* interactions; /* [n_interactions*2] column pairs             */
  const double* input_scales;  /* [D] (spatiotemporal.py:189-192)             */
} bnf_config_t;

typedef struct bnf_plan_info {
  int32_t num_params;          /* P                                           */
  int32_t num_features;        /* F, true fan-in of Dense_0                   */
  int32_t padded_features;     /* F rounded up for the tensor-core path       */
  int32_t num_leaves;          /* leaves after the 3 scalar heads             */
  int32_t num_feature_groups;
  int32_t sm_count;            /* SMs of the current device (0 if none)       */
} bnf_plan_info_t;

int bnf_abi_version(void);
const char* bnf_last_error(void);

/* make_model / make_prior bookkeeping (inference.py:234-268; models.py:216-252).
 * Pure host work: succeeds without a GPU.                                    */
int bnf_plan_create(const bnf_config_t* cfg, bnf_plan_t** out_plan);
void bnf_plan_destroy(bnf_plan_t* plan);
int bnf_plan_info(const bnf_plan_t* plan, bnf_plan_info_t* out);
/* leaf 0.. in reference order (after the three scalars at offsets 0,1,2).     */
int bnf_plan_leaf(const bnf_plan_t* plan, int32_t leaf, char* name, int32_t name_len,
                  int64_t* offset, int32_t* rows, int32_t* cols);

/* 0 when `precision` can run this plan (tensor-core modes need width % 64 == 0; BNF_PREC_BF16X3
 * also width in {64,128,256,512,1024} and <= 128 encoded features), else the error code the
 * compute entry points would return (message in bnf_last_error()).  Host-only.          */
int bnf_precision_supported(const bnf_plan_t* plan, int32_t precision);

/* Bytes of scratch a call needs for `n_networks` x `batch_rows`.  `mode` says
 * which entry point will use it (for BNF_WS_VI n_networks = S * members).     */
enum { BNF_WS_FORWARD = 0, BNF_WS_GRAD = 1, BNF_WS_MAP = 2, BNF_WS_VI = 3 };
size_t bnf_workspace_bytes(const bnf_plan_t* plan, int32_t precision,
                           int32_t n_networks, int32_t batch_rows, int32_t mode);

/* Row selection shared by the calls below: network j reads row
 * idx[j*idx_stride + i] of (x, y) for i in [0, batch_rows); idx == NULL means
 * rows 0..batch_rows-1 for every network; idx_stride == 0 shares one index row
 * (VI sub-batch, inference.py:704-709).  x is (n_rows_total, D) f32 row-major
 * -- the jnp.array() f32 cast of inference.py:553-554 is done by the caller. */

/* mlp.apply for every network (models.py:213-273; forecast_inner,
 * inference.py:103-126): out_loc[j*batch_rows + i] = network output.          */
int bnf_forward(const bnf_plan_t* plan, int32_t precision, const float* params,
                int32_t n_networks, const float* x, const int32_t* idx,
                int64_t idx_stride, int32_t batch_rows, float* out_loc,
                void* workspace, size_t workspace_bytes, void* stream);

/* make_likelihood_model(...).log_prob(y) and its gradient
 * (models.py:106-194; jax.value_and_grad at inference.py:602):
 * out_loglik[j] = sum_i log p(y_i | network j); out_grad[j*P + p] = d/dparam.
 * out_grad may be NULL (value only).                                         */
int bnf_loglik_grad(const bnf_plan_t* plan, int32_t precision, const float* params,
                    int32_t n_networks, const float* x, const float* y,
                    const int32_t* idx, int64_t idx_stride, int32_t batch_rows,
                    float* out_loglik, float* out_grad, void* workspace,
                    size_t workspace_bytes, void* stream);

/* `n_steps` consecutive `_one_step`s of ensemble_map._run (inference.py:599-608):
 * loss_j = -(loglik_j * n_total/batch_rows + prior_weight * logprior_j)
 * (inference.py:558-569; prior models.py:94-103), value_and_grad, optax.adam
 * (b1=.9, b2=.999, eps=1e-8, inference.py:580,605-606).  Step s uses index rows
 * [s*batch_rows, (s+1)*batch_rows) of each network's index row.
 * params/adam_m/adam_v: [n_networks, P] updated in place; step_count: device
 * int32 shared by all networks, incremented per step.  out_loss[s*n_networks+j]
 * is the loss BEFORE update s.                                               */
int bnf_map_steps(const bnf_plan_t* plan, int32_t precision, float* params,
                  float* adam_m, float* adam_v, int32_t* step_count,
                  int32_t n_networks, const float* x, const float* y,
                  const int32_t* idx, int64_t idx_stride, int32_t batch_rows,
                  int32_t n_rows_total, int32_t n_steps, float learning_rate,
                  float prior_weight, float* out_loss, void* workspace,
                  size_t workspace_bytes, void* stream);

/* `n_epochs` epochs of `_one_epoch` (inference.py:583-614) with the minibatch order drawn ON THE
 * DEVICE: every epoch each member walks a fresh permutation of the n_rows_total rows
 * (permute_dataset, :35-39, vmapped over members :593-597) in windows of batch_rows rows, the
 * ragged tail dropped (:583-589) -- n_rows_total / batch_rows steps per epoch.  The permutation of
 * (seed; first_member + j, epoch) is a keyed bijection evaluated per batch window (no sort, no
 * index array; bnf_debug_permutation evaluates the same function on the host), the epoch /
 * window follow from the device-side step_count, so the whole call replays one CUDA graph.
 * out_loss [n_epochs * steps_per_epoch, n_networks]: loss before each update (the epoch loss of
 * :614 is the mean over an epoch's rows).                                              */
int bnf_map_epochs(const bnf_plan_t* plan, int32_t precision, float* params, float* adam_m,
                   float* adam_v, int32_t* step_count, int32_t n_networks, const float* x,
                   const float* y, int32_t batch_rows, int32_t n_rows_total, int32_t n_epochs,
                   float learning_rate, float prior_weight, uint64_t seed, int64_t first_member,
                   float* out_loss, void* workspace, size_t workspace_bytes, void* stream);

/* One step of tfp.vi.fit_surrogate_posterior_stateless as driven by
 * ensemble_vi (inference.py:687-739): q = prod N(mu, 1e-4 + softplus(rho)),
 * z_s = mu + sigma*eps_s, loss_e = mean_s[log q(z_s) - logprior(z_s)
 * - loglik(z_s)*(n_total/batch_rows)/kl_weight], reparameterised gradient, Adam
 * on (mu, rho).  eps: [S, E, P] standard normals, or NULL to draw them on the
 * device from (seed, step_count).  out_loss[e] is the loss before the update,
 * NOT yet multiplied by kl_weight (inference.py:758 does that on the host).   */
int bnf_vi_step(const bnf_plan_t* plan, int32_t precision, float* mu, float* rho,
                float* adam_m, float* adam_v, int32_t* step_count,
                int32_t n_members, int32_t n_mc_samples, const float* eps,
                uint64_t seed, const float* x, const float* y, const int32_t* idx,
                int32_t batch_rows, int32_t n_rows_total, float learning_rate,
                float kl_weight, float* out_loss, void* workspace,
                size_t workspace_bytes, void* stream);

/* `n_steps` such steps with everything drawn on the device (the production path; bnf_vi_step
 * with injected eps / idx is the test hook): eps from the Philox stream of (seed, step count);
 * when batch_rows < n_rows_total one shared random sub-batch per step = the first batch_rows
 * entries of the permutation keyed by (seed; device_id, step count) (inference.py:704-709).
 * The step sequence replays one CUDA graph.  out_loss [n_steps, n_members].               */
int bnf_vi_steps(const bnf_plan_t* plan, int32_t precision, float* mu, float* rho,
                 float* adam_m, float* adam_v, int32_t* step_count, int32_t n_members,
                 int32_t n_mc_samples, uint64_t seed, int64_t device_id, const float* x,
                 const float* y, int32_t batch_rows, int32_t n_rows_total, int32_t n_steps,
                 float learning_rate, float kl_weight, float* out_loss, void* workspace,
                 size_t workspace_bytes, void* stream);

/* surrogate.sample(num_samples) (inference.py:741-753): out[s,e,:] = mu_e +
 * (1e-4+softplus(rho_e)) * eps[s,e,:]; eps NULL -> device Philox from seed.    */
int bnf_vi_sample(const bnf_plan_t* plan, const float* mu, const float* rho,
                  int32_t n_members, int32_t n_samples, const float* eps,
                  uint64_t seed, float* out_params, void* stream);

/* _make_init_fn / make_vi_init (inference.py:399-427, :203-231): leaf 0 =
 * `log_noise_scale_init`, every 2-D kernel ~ TruncatedNormal(0,1,[-2,2]) from a
 * device Philox stream keyed by (seed, first_member + j), all else 0.          */
int bnf_init_params(const bnf_plan_t* plan, float log_noise_scale_init, uint64_t seed,
                    int64_t first_member, int32_t n_networks, float* out_params,
                    void* stream);

/* _approximate_normal_quantile / _normal_quantile_via_root (inference.py:42-84)
 * over a mixture of n_components Normals per point: means [n_components,
 * n_points], scales [n_components]; q: HOST array of n_q quantiles;
 * out [n_q, n_points].  Root mode: Chandrupatla on the global bracket
 * [min mu - 5 max sigma, max mu + 5 max sigma], value tolerance 1e-5, <=60 its. */
int bnf_mixture_quantiles(const float* means, const float* scales,
                          int32_t n_components, int32_t n_points, const double* q,
                          int32_t n_q, int32_t approximate, float* out,
                          void* workspace, size_t workspace_bytes, void* stream);
size_t bnf_quantile_workspace_bytes(int32_t n_components, int32_t n_points);

/* NB / ZINB predictive mean and quantiles of the ensemble mixture
 * (_build_observation_distribution + _get_nb_quantiles_root, inference.py:271-333).
 * loc [n_components, n_points] = network outputs (bnf_forward); shape_raw /
 * pi_logit [n_components] = params[1] / params[2] (pi_logit NULL -> plain NB).
 * out_means [n_components, n_points] = distribution mean (obs_d.mean()); out_q
 * [n_q, n_points] = min{k >= 0 integer : mean-CDF(k) >= q} capped at ceil(high), high =
 * max(mean) + 1.1*rsqrt(1-q)*max(stddev) -- the integer the reference's
 * ceil(Chandrupatla root) lands on.  CDF = regularised incomplete beta, in f64.  */
int bnf_nb_mixture_quantiles(const float* loc, const float* shape_raw, const float* pi_logit,
                             int32_t n_components, int32_t n_points, const double* q, int32_t n_q,
                             float* out_means, float* out_q, void* workspace, size_t workspace_bytes,
                             void* stream);

/* ---- test / profiling hooks (not part of the reference seam) ----------------
 * bnf_debug_gemm: run the tcgen05 GEMM kernel alone, C[net][M][N] (f32) =
 *   mn_major == 0: A[net][M][K] x B[net][N][K]^T   (both K-major, as fwd/dgrad use it)
 *   mn_major == 1: A[net][K][M]^T x B[net][K][N]   (both MN-major, as wgrad uses it)
 *   mn_major == 2: A[net][M][K] x B[net][K][N]     (A K-major, B MN-major: forward)
 *   mn_major == 3, 4, 5: the bf16x3 (split-operand) versions of 2, 1, 0 -- every operand row
 *     holds three bf16 planes side by side (A[net][M][3K] ...), C = the f32-accurate product.
 * A, B are bf16.  Used by tests/test_gpu_tc.py against a plain f32 matmul.      */
int bnf_debug_gemm(int32_t mn_major, const void* a, const void* b, float* c,
                   int32_t n_networks, int32_t m, int32_t n, int32_t k, void* stream);
/* The device's permute_dataset for (seed; member, epoch), evaluated on the HOST: out[i] = row at
 * position i (a permutation of 0..n-1).                                                  */
int bnf_debug_permutation(uint64_t seed, int64_t member, int32_t epoch, int32_t n, int32_t* out);
/* Philox4x32-10 block function of the device RNG (init / VI draws / minibatch permutations),
 * evaluated on the HOST: counter4 -> out4 under key2.  Known-answer tested against the
 * Random123 vectors without a GPU.                                              */
void bnf_debug_philox(const uint32_t* counter4, const uint32_t* key2, uint32_t* out4);
/* Kernel launches issued by this library since load (all threads).             */
uint64_t bnf_debug_launch_count(void);
/* enable != 0: bracket every kernel launch with CUDA events on its stream (and
 * forget earlier records); 0: stop.  bnf_debug_profile_report synchronises the
 * device and writes "name count total_ms\n" per kernel class into buf.         */
int bnf_debug_profile(int32_t enable);
int bnf_debug_profile_report(char* buf, int32_t len);

#ifdef __cplusplus
}
#endif
#endif  /* BNF_H_ */
