// Device math shared by the SIMT (fp32) and tcgen05 (bf16) paths.
// No -use_fast_math anywhere: the f32 trig arguments reach 1e3..1e5 rad
// (models.py:73) and need the accurate sinf/cosf slow path.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "bnf_model.h"

namespace bnf {

// ---- per-network derived scalars (written by prep_kernel) -------------------
constexpr int kDvActW = 0;      // sigmoid(logit_activation_weight)   models.py:253
constexpr int kDvSOut = 1;      // softplus(inv_sp_output_scale)      models.py:270
constexpr int kDvSigma = 2;     // 0.01 + exp(log_noise_scale)        models.py:163
constexpr int kDvShape = 3;     // softplus(params[1])                models.py:171
constexpr int kDvPi = 4;        // sigmoid(params[2])                 models.py:184
constexpr int kDvSLayer = 8;    // softplus(inv_sp_layer_scale{l})    models.py:265
constexpr int kDvDenom = 24;    // input_scales[i]*exp(lsa[i])        models.py:221
constexpr int kDvSX = 40;       // softplus(feature_inv_sp_scale) of scaled_x
constexpr int kDvSSeas = 41;
constexpr int kDvSInter = 42;
constexpr int kDvSFourier = 43; // + input dim i

// ---- programmatic dependent launch (PDL) --------------------------------------
// Every kernel of a training step starts with pdl_trigger(); <prologue>; pdl_wait():
// launched with cudaLaunchAttributeProgrammaticStreamSerialization the next kernel's CTAs may
// become resident (barrier init, TMEM alloc, tensor-map prefetch) while this one drains, and
// griddepcontrol.wait blocks them until every prerequisite grid has completed and flushed.
// Without the launch attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// `const __restrict__` loads are treated as kernel-lifetime invariants and may be hoisted ABOVE
// an asm memory barrier (seen in SASS: LDG.E.CONSTANT before ACQBULK) -- with PDL that reads
// data the previous kernel has not written yet.  pdl_enter() therefore runs first in a kernel
// and passes every pointer through an opaque asm after the wait: loads through the laundered
// pointers carry a true dependency on it.
// (an opaque ZERO OFFSET is added rather than the pointer itself being laundered, so the
// compiler still knows the pointer derives from a kernel parameter = global address space)
template <typename T> __device__ __forceinline__ void pdl_launder(T& p) {   // T = any pointer type
  long long z = 0;
  asm volatile("" : "+l"(z) :: "memory");
  p = (T)((const char*)p + z);
}
template <typename... P> __device__ __forceinline__ void pdl_enter(P&... ptrs) {
  pdl_trigger();
  pdl_wait();
  (pdl_launder(ptrs), ...);
}

__device__ __forceinline__ float softplus_f(float x) {
  // jax.nn.softplus == logaddexp(x, 0)
  return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));
}
__device__ __forceinline__ float sigmoid_f(float x) {
  // numerically safe on both tails
  if (x >= 0.f) return 1.f / (1.f + expf(-x));
  float e = expf(x);
  return e / (1.f + e);
}
__device__ __forceinline__ float log_sigmoid_f(float x) { return -softplus_f(-x); }

// activation models.py:255-258 : w*elu(z) + (1-w)*tanh(z)
__device__ __forceinline__ float act_f(float z, float w) {
  float elu = z > 0.f ? z : expm1f(z);
  return w * elu + (1.f - w) * tanhf(z);
}
// returns act'(z); *diff = elu(z) - tanh(z)  (d act / d w)
__device__ __forceinline__ float act_grad_f(float z, float w, float* diff) {
  float t = tanhf(z);
  float elu, delu;
  if (z > 0.f) { elu = z; delu = 1.f; } else { delu = expf(z); elu = expm1f(z); }
  *diff = elu - t;
  return w * delu + (1.f - w) * (1.f - t * t);
}

// bf16-mode variants: one MUFU.TANH + one MUFU.EX2 per element.  Their error
// (~2^-11 relative for tanh.approx, 2 ulp for ex2.approx) is below the bf16
// rounding (2^-9) applied to everything this feeds.
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float act_fast(float z, float w) {
  const float t = tanh_fast(z);
  const float elu = z > 0.f ? z : ex2_fast(z * 1.4426950408889634f) - 1.f;
  return fmaf(w, elu - t, t);
}
__device__ __forceinline__ float act_grad_fast(float z, float w, float* diff) {
  const float t = tanh_fast(z);
  const float e = ex2_fast(z * 1.4426950408889634f);
  const float elu = z > 0.f ? z : e - 1.f;
  const float delu = z > 0.f ? 1.f : e;
  *diff = elu - t;
  const float dt = fmaf(-t, t, 1.f);
  return fmaf(w, delu - dt, dt);
}
// ---- bf16x3 (f32-parity tensor-core) mode ----------------------------------------------------
// Activation with f32-class accuracy at MUFU cost: one ex2 and one rcp per element.
//   ea = e^-|z| (ex2.approx: 2 ulp), q = ea^2, tanh|z| = 1 - 2q/(1+q) (rcp.approx: 1 ulp),
//   elu = z > 0 ? z : ea - 1, elu' = z > 0 ? 1 : ea.
// Absolute error <= ~3e-7 on h, act' and diff (|h| is O(1); the parity bar is 1e-5 of the output
// scale), no overflow for any finite z (ea <= 1).
__device__ __forceinline__ float rcp_fast(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// returns act'(z); *diff = elu(z) - tanh(z); *h = act(z)
__device__ __forceinline__ float act_grad_x3(float z, float w, float* diff, float* h) {
  const float ea = ex2_fast(-fabsf(z) * 1.4426950408889634f);
  const float q = ea * ea;
  const float t = copysignf(fmaf(-2.f * q, rcp_fast(1.f + q), 1.f), z);
  const bool pos = z > 0.f;
  const float elu = pos ? z : ea - 1.f;
  const float delu = pos ? 1.f : ea;
  const float d = elu - t;
  *diff = d;
  *h = fmaf(w, d, t);
  const float dt = fmaf(-t, t, 1.f);
  return fmaf(w, delu - dt, dt);
}
// Two f32 values -> their three bf16 planes, packed (lo = a, hi = b) per plane:
// p0 = bf16(v), p1 = bf16(v - p0), p2 = bf16(v - p0 - p1); the subtractions are exact in f32.
__device__ __forceinline__ void split3_pair(float a, float b, uint32_t* p0, uint32_t* p1, uint32_t* p2) {
  __nv_bfloat162 t0 = __floats2bfloat162_rn(a, b);
  const uint32_t u0 = *reinterpret_cast<uint32_t*>(&t0);
  const float ra = a - __uint_as_float(u0 << 16), rb = b - __uint_as_float(u0 & 0xffff0000u);
  __nv_bfloat162 t1 = __floats2bfloat162_rn(ra, rb);
  const uint32_t u1 = *reinterpret_cast<uint32_t*>(&t1);
  const float sa = ra - __uint_as_float(u1 << 16), sb = rb - __uint_as_float(u1 & 0xffff0000u);
  __nv_bfloat162 t2 = __floats2bfloat162_rn(sa, sb);
  *p0 = u0; *p1 = u1; *p2 = *reinterpret_cast<uint32_t*>(&t2);
}
// two planes only (the backpropagated dU of the bf16x3 mode): p0 = bf16(v), p1 = bf16(v - p0)
__device__ __forceinline__ void split2_pair(float a, float b, uint32_t* p0, uint32_t* p1) {
  __nv_bfloat162 t0 = __floats2bfloat162_rn(a, b);
  const uint32_t u0 = *reinterpret_cast<uint32_t*>(&t0);
  __nv_bfloat162 t1 = __floats2bfloat162_rn(a - __uint_as_float(u0 << 16), b - __uint_as_float(u0 & 0xffff0000u));
  *p0 = u0; *p1 = *reinterpret_cast<uint32_t*>(&t1);
}
__device__ __forceinline__ void split3_one(float a, __nv_bfloat16* p0, __nv_bfloat16* p1, __nv_bfloat16* p2) {
  const __nv_bfloat16 h0 = __float2bfloat16_rn(a);
  const float r = a - __bfloat162float(h0);
  const __nv_bfloat16 h1 = __float2bfloat16_rn(r);
  *p0 = h0; *p1 = h1; *p2 = __float2bfloat16_rn(r - __bfloat162float(h1));
}

// ---- packed f32x2 arithmetic (sm_100: FFMA2 on a 64-bit register pair) ---------------------
// The tcgen05 epilogues are bound by the fma pipe (a 3-register FFMA has a reciprocal throughput
// of 2 cycles per SMSP) and by issue slots; FFMA2 does two lanes' worth of work per issue, so the
// element-wise math of those epilogues runs on pairs of adjacent accumulator columns.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 f2_pack(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ f32x2 f2_pack(uint32_t lo, uint32_t hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ f32x2 f2_dup(float x) { return f2_pack(x, x); }
__device__ __forceinline__ float f2_lo(f32x2 v) {
  float lo;
  [[maybe_unused]] float hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
  return lo;
}
__device__ __forceinline__ float f2_hi(f32x2 v) {
  [[maybe_unused]] float lo;
  float hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
  return hi;
}
__device__ __forceinline__ f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 f2_mul(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 f2_add(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint32_t f2_to_bf16x2(f32x2 v) {   // lo -> bits 0..15, hi -> bits 16..31
  __nv_bfloat162 t = __floats2bfloat162_rn(f2_lo(v), f2_hi(v));
  return *reinterpret_cast<uint32_t*>(&t);
}
// Per-kernel constants of the packed activation math (models.py:255-258).
struct ActConst2 {
  f32x2 w, omw, n_omw, log2e, n_one;
  __device__ __forceinline__ explicit ActConst2(float wa)
      : w(f2_dup(wa)), omw(f2_dup(1.f - wa)), n_omw(f2_dup(wa - 1.f)), log2e(f2_dup(1.4426950408889634f)),
        n_one(f2_dup(-1.f)) {}
};
// Two activations at once; same element math as act_fast / act_grad_fast, written without
// selects (a select of two packed halves costs two moves each): with m = min(z,0), r = max(z,0)
//   e = 2^(m*log2e) (exactly 1 for z > 0), elu = r + (e-1), elu' = e, t = tanh(z),
//   diff = elu - t, h = w*diff + t, act'(z) = w*e + (1-w)*(1-t^2)
__device__ __forceinline__ f32x2 act_fast2(f32x2 z, const ActConst2& k) {
  const float z0 = f2_lo(z), z1 = f2_hi(z);
  const f32x2 t = f2_pack(tanh_fast(z0), tanh_fast(z1));
  const f32x2 zl = f2_mul(f2_pack(fminf(z0, 0.f), fminf(z1, 0.f)), k.log2e);
  const f32x2 em1 = f2_add(f2_pack(ex2_fast(f2_lo(zl)), ex2_fast(f2_hi(zl))), k.n_one);
  const f32x2 elu = f2_add(f2_pack(fmaxf(z0, 0.f), fmaxf(z1, 0.f)), em1);
  return f2_fma(k.w, f2_fma(t, k.n_one, elu), t);
}
__device__ __forceinline__ f32x2 act_grad_fast2(f32x2 z, const ActConst2& k, f32x2* diff, f32x2* h = nullptr) {
  const float z0 = f2_lo(z), z1 = f2_hi(z);
  const f32x2 t = f2_pack(tanh_fast(z0), tanh_fast(z1));
  const f32x2 zl = f2_mul(f2_pack(fminf(z0, 0.f), fminf(z1, 0.f)), k.log2e);
  const f32x2 e = f2_pack(ex2_fast(f2_lo(zl)), ex2_fast(f2_hi(zl)));
  const f32x2 elu = f2_add(f2_pack(fmaxf(z0, 0.f), fmaxf(z1, 0.f)), f2_add(e, k.n_one));
  const f32x2 d = f2_fma(t, k.n_one, elu);
  *diff = d;
  if (h) *h = f2_fma(k.w, d, t);
  const f32x2 dtw = f2_fma(f2_mul(t, t), k.n_omw, k.omw);      // (1-w)*(1-t^2)
  return f2_fma(k.w, e, dtw);
}

// Column sums over a warp's 32 rows: lane r holds row r's 32 values v[0..31]; a 5-step
// transpose-reduce (31 shuffles) leaves sum_r v_r[L] in v[0] of lane L.
__device__ __forceinline__ void warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
  for (int hh = 16; hh >= 1; hh >>= 1) {
    const bool up = (lane & hh) != 0;
#pragma unroll
    for (int i = 0; i < hh; ++i) {
      const float send = up ? v[i] : v[i + hh];
      const float keep = up ? v[i + hh] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, hh);
    }
  }
}

// 16-column variant: lane r holds 16 values of row r; after 4 halving steps and one pair add,
// lanes 2c and 2c+1 both hold sum_r v_r[c] in v[0]  (column = lane >> 1).
__device__ __forceinline__ void warp_transpose_sum16(float (&v)[16], int lane) {
#pragma unroll
  for (int hh = 8; hh >= 1; hh >>= 1) {
    const int bit = hh << 1;                       // lane bit that picks the half kept
    const bool up = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < hh; ++i) {
      const float send = up ? v[i] : v[i + hh];
      const float keep = up ? v[i + hh] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// sin/cos for the bf16 path with arguments up to ~1e5 rad (seasonal features): one exact
// f32 range reduction to [-pi, pi] (two-constant Cody-Waite) and the MUFU approximations.
// Absolute error ~1e-6 + |x|*6e-8 (the latter is the rounding of the f32 argument itself,
// which the reference shares), far below the bf16 rounding of the stored feature.
__device__ __forceinline__ void sincos_reduced(float x, float* sn, float* cs) {
  const float k = rintf(x * 0.15915494309189535f);          // x / 2pi
  float r = fmaf(-k, 6.28318548202514648f, x);              // hi part of 2pi (f32)
  r = fmaf(-k, -1.74845553e-7f, r);                         // lo part: 2pi - f32(2pi)
  __sincosf(r, sn, cs);
}
template <bool FAST> __device__ __forceinline__ float act_sel(float z, float w) {
  return FAST ? act_fast(z, w) : act_f(z, w);
}
template <bool FAST> __device__ __forceinline__ float act_grad_sel(float z, float w, float* diff) {
  return FAST ? act_grad_fast(z, w, diff) : act_grad_f(z, w, diff);
}
template <typename T> struct FastMath { static constexpr bool value = false; };
template <> struct FastMath<__nv_bfloat16> { static constexpr bool value = true; };

__device__ __forceinline__ float digamma_f(float x) {
  // psi(x) for x > 0: upward recurrence to x >= 6 then the asymptotic series.
  float acc = 0.f;
  while (x < 6.f) { acc -= 1.f / x; x += 1.f; }
  float inv = 1.f / x, inv2 = inv * inv;
  return acc + logf(x) - 0.5f * inv
         - inv2 * (1.f / 12.f - inv2 * (1.f / 120.f - inv2 * (1.f / 252.f)));
}

// ---- per-row log-likelihood of the head (models.py:157-191) ------------
// o = network output of the row, yv = observation, dv = the network's derived scalars.
// Returns log p(y | o); *rr = d logp / d o; g[0..2] receive this row's addends of the raw
// gradient sums w.r.t. sigma (NORMAL: multiplied by exp(log_noise_scale) later), shape
// (NB/ZINB: by sigmoid(shape_raw)) and pi (ZINB: by pi(1-pi)).  Same expressions as head_kernel.
__device__ __forceinline__ float head_row_loglik(int likelihood, const float* __restrict__ dv, float o, float yv,
                                                 float* rr, float* g) {
  float logp;
  if (likelihood == BNF_NORMAL) {
    const float sg = dv[kDvSigma];
    const float d = yv / sg - o / sg;
    logp = -0.5f * d * d - (0.9189385332046727f + logf(sg));
    *rr = d / sg;
    g[0] += (d * d - 1.f) / sg;
  } else {
    const float mean = softplus_f(o);
    const float shp = dv[kDvShape];
    const float rc = 1.f / shp;
    const float lg = -logf(shp) - logf(mean);
    const float sig_l = sigmoid_f(lg);
    const float nb = rc * log_sigmoid_f(-lg) + yv * log_sigmoid_f(lg)
                     - (lgammaf(1.f + yv) + lgammaf(rc) - lgammaf(1.f + yv + rc)) - logf(rc + yv);
    const float dnb_dl = yv * (1.f - sig_l) - rc * sig_l;
    const float dnb_dr = log_sigmoid_f(-lg) - digamma_f(rc) + digamma_f(1.f + yv + rc) - 1.f / (rc + yv);
    float wnb = 1.f;
    logp = nb;
    if (likelihood == BNF_ZINB) {
      const float pi = dv[kDvPi];
      if (yv == 0.f) {
        const float A = (1.f - pi) * expf(nb), tot = A + pi;
        logp = logf(tot);
        wnb = A / tot;
        g[2] += (1.f - expf(nb)) / tot;
      } else {
        logp = log1pf(-pi) + nb;
        g[2] += -1.f / (1.f - pi);
      }
    }
    *rr = wnb * dnb_dl * (-sigmoid_f(o) / mean);
    g[1] += wnb * (dnb_dl * (-1.f / shp) + dnb_dr * (-1.f / (shp * shp)));
  }
  return logp;
}

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) {
  return __bfloat162float(v);
}
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

// ---- counter-based RNG: Philox4x32-10 (Salmon et al., SC'11) --------------------------------
// One call maps a 128-bit counter + 64-bit key to four independent 32-bit words; every consumer
// below gives each output element its OWN counter block, so draws never share words.
__host__ __device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c.x, p1 = (uint64_t)0xCD9E8D57u * c.z;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
// ---- keyed permutation of [0, n) without a sort (permute_dataset, inference.py:35-39) ---------
// A balanced Feistel network on 2*half_bits bits (2^(2*half_bits) >= n) is a bijection of
// [0, 2^(2*half_bits)); walking its cycle until the value falls below n restricts it to a bijection
// of [0, n) (format-preserving encryption by cycle walking; cf. Mitchell et al., "Bandwidth-optimal
// random shuffling for GPUs", 2021).  Every output index is computed independently in O(1): a
// training step only evaluates the B entries of its batch window, no index array is materialised
// for the epoch.  Round keys: two Philox blocks keyed by (seed; member, epoch).
constexpr int kPermRounds = 8;
struct PermKeys { uint32_t k[kPermRounds]; };
__host__ __device__ __forceinline__ PermKeys perm_keys(uint64_t seed, uint64_t member, uint32_t epoch) {
  PermKeys pk;
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  const uint4 a = philox4x32_10(make_uint4(epoch, 0x50524d30u, (uint32_t)member, (uint32_t)(member >> 32)), key);
  const uint4 b = philox4x32_10(make_uint4(epoch, 0x50524d31u, (uint32_t)member, (uint32_t)(member >> 32)), key);
  pk.k[0] = a.x; pk.k[1] = a.y; pk.k[2] = a.z; pk.k[3] = a.w;
  pk.k[4] = b.x; pk.k[5] = b.y; pk.k[6] = b.z; pk.k[7] = b.w;
  return pk;
}
__host__ __device__ __forceinline__ uint32_t perm_half_bits(uint32_t n) {
  uint32_t bits = 1;
  while (bits < 16 && (1ull << (2 * bits)) < (unsigned long long)n) ++bits;
  return bits;                       // n <= 2^31 (row counts are int32)
}
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {      // murmur3 finaliser
  x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ uint32_t perm_index(uint32_t i, uint32_t n, uint32_t half_bits, const PermKeys& pk) {
  const uint32_t mask = (1u << half_bits) - 1u;
  uint32_t v = i;
  do {
    uint32_t l = v >> half_bits, r = v & mask;
#pragma unroll
    for (int q = 0; q < kPermRounds; ++q) {
      const uint32_t f = mix32(r ^ pk.k[q]) & mask;
      const uint32_t nr = l ^ f;
      l = r;
      r = nr;
    }
    v = (l << half_bits) | r;
  } while (v >= n);
  return v;
}

__device__ __forceinline__ float u32_to_unit(uint32_t x) {   // (0, 1]
  return fmaf((float)x, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
}
// four standard normals from one Philox block (two Box-Muller pairs)
__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint64_t stream, uint64_t block) {
  const uint4 r = philox4x32_10(make_uint4((uint32_t)block, (uint32_t)(block >> 32), (uint32_t)stream, (uint32_t)(stream >> 32)),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float r0 = sqrtf(-2.f * logf(u32_to_unit(r.x))), r1 = sqrtf(-2.f * logf(u32_to_unit(r.z)));
  float s0, c0, s1, c1;
  sincosf(6.283185307179586f * u32_to_unit(r.y), &s0, &c0);
  sincosf(6.283185307179586f * u32_to_unit(r.w), &s1, &c1);
  return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- encode work items: (row, unit) ----------------------------------------
struct UnitInfo { int kind, a, b; };  // kind 0:x 1:fourier(dim a, degree b) 2:seasonal(k=a) 3:inter(j=a)

__device__ __forceinline__ int num_units(const DevModel& m) {
  int u = m.D;
  for (int i = 0; i < m.D; ++i) u += m.fourier_deg[i] > 0 ? m.fourier_deg[i] : 0;
  return u + m.n_seasonal + m.n_inter;
}
__device__ __forceinline__ UnitInfo decode_unit(const DevModel& m, int u) {
  if (u < m.D) return {0, u, 0};
  u -= m.D;
  for (int i = 0; i < m.D; ++i) {
    int deg = m.fourier_deg[i] > 0 ? m.fourier_deg[i] : 0;
    if (u < deg) return {1, i, u};
    u -= deg;
  }
  if (u < m.n_seasonal) return {2, u, 0};
  return {3, u - m.n_seasonal, 0};
}

__device__ __forceinline__ const float* row_ptr(const float* x, const int32_t* idx, int64_t idx_stride,
                                                int net, int b, int D) {
  int64_t r = idx ? (int64_t)idx[(int64_t)net * idx_stride + b] : (int64_t)b;
  return x + r * D;
}


// ---- feature encode for one row (models.py:216-252) -------------------------
// xr: the raw input row (D floats); dv: this network's derived scalars.
// emit(col, value) is called once per feature column in [0, F).
template <typename Emit>
__device__ __forceinline__ void encode_row(const DevModel& m, const float* __restrict__ dv,
                                           const float* xr, Emit emit) {
  float sx[kMaxD];
#pragma unroll 4
  for (int i = 0; i < m.D; ++i) sx[i] = xr[i] / dv[kDvDenom + i];
  {
    const float s = dv[kDvSX];
    for (int i = 0; i < m.D; ++i) emit(m.col_x + i, sx[i] * s);
  }
  const float two_pi = 6.283185307179586f;  // f32(2*pi), models.py:85
  for (int i = 0; i < m.D; ++i) {
    const int deg = m.fourier_deg[i];
    if (deg <= 0) continue;
    const float s = dv[kDvSFourier + i];
    const int c0 = m.fourier_col[i];
    float c = two_pi;                        // 2*pi*2^d is exact scaling in f32
    for (int d = 0; d < deg; ++d, c *= 2.f) {
      float sn, cs;
      sincosf(c * sx[i], &sn, &cs);
      const float den = (float)(d + 1);
      emit(c0 + d, (cs / den) * s);
      emit(c0 + deg + d, (sn / den) * s);
    }
  }
  if (m.n_seasonal > 0) {
    const float s = dv[kDvSSeas];
    const float t = xr[0];                   // RAW time, models.py:223
    for (int k = 0; k < m.n_seasonal; ++k) {
      float sn, cs;
      sincosf(m.seasonal_w[k] * t, &sn, &cs);
      emit(m.col_seasonal + k, (cs / m.seasonal_h[k]) * s);
      emit(m.col_seasonal + m.n_seasonal + k, (sn / m.seasonal_h[k]) * s);
    }
  }
  if (m.n_inter > 0) {
    const float s = dv[kDvSInter];
    for (int j = 0; j < m.n_inter; ++j)
      emit(m.col_inter + j, (sx[m.inter_a[j]] * sx[m.inter_b[j]]) * s);
  }
}

}  // namespace bnf
