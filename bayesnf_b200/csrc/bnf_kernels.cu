// SIMT kernels of the BayesNF hot path: feature encode, fp32 GEMMs (parity
// mode), head / likelihood, activation backward, encode backward, prior + Adam,
// VI sampling / gradient assembly, mixture quantiles.  The bf16 tcgen05 GEMMs
// live in bnf_tc.cu and share every non-GEMM kernel in this file.
#include <atomic>
#include <cfloat>
#include <cstdio>
#include <cstdlib>

#include "bnf_device.cuh"
#include "bnf_kernels.h"
#include "bnf_prof.h"

namespace bnf {

// =============================================================================
// prep: per-network derived scalars
// =============================================================================
// One network's derived scalars, computed by the first 32 threads t = 0..31 of the caller.
// Loads bypass L1 (__ldcg): map_update_kernel calls this right after other blocks of the same
// grid wrote the parameters.
__device__ __forceinline__ void prep_one(const DevModel& m, const float* p, float* dv, int t) {
  if (t == 0) {
    dv[kDvActW] = sigmoid_f(__ldcg(p + m.off_actw));
    dv[kDvSOut] = softplus_f(__ldcg(p + m.off_out_scale));
    dv[kDvSigma] = 0.01f + expf(__ldcg(p));
    dv[kDvShape] = softplus_f(__ldcg(p + 1));
    dv[kDvPi] = 1.f / (1.f + expf(-__ldcg(p + 2)));  // literal models.py:184
    dv[kDvSX] = softplus_f(__ldcg(p + m.off_scale_x));
    dv[kDvSSeas] = m.off_scale_seasonal >= 0 ? softplus_f(__ldcg(p + m.off_scale_seasonal)) : 0.f;
    dv[kDvSInter] = m.off_scale_inter >= 0 ? softplus_f(__ldcg(p + m.off_scale_inter)) : 0.f;
  }
  if (t < m.L) dv[kDvSLayer + t] = softplus_f(__ldcg(p + m.off_layer_scale[t]));
  if (t < m.D) {
    dv[kDvDenom + t] = m.input_scales[t] * expf(__ldcg(p + m.off_lsa + t));
    dv[kDvSFourier + t] = m.fourier_scale_off[t] >= 0 ? softplus_f(__ldcg(p + m.fourier_scale_off[t])) : 0.f;
  }
}

// zero_acc (optional): [ll | prior] accumulators of n_net floats each at zero_acc / zero_acc2,
// zero_cursors (optional): two int32 cursors; loss_slot (optional): where map_update_kernel finds
// the loss buffer of the current call -- the prologue of bnf_map_steps in one launch.
__global__ void __launch_bounds__(256)
prep_kernel(const __grid_constant__ DevModel m, const float* params, float* __restrict__ derived, int n_net,
            float* zero_acc, float* zero_acc2, int32_t* zero_cursors, float** loss_slot, float* out_loss,
            float* __restrict__ zero_rows) {
  pdl_enter(params, derived, zero_acc, zero_acc2, zero_cursors, zero_rows);
  const int net = blockIdx.y;
  if (net >= n_net) return;
  // zero_rows (optional): the [n_net, P] gradient accumulator, cleared here instead of by a separate
  // memset node (one launch less in front of every call's first step)
  if (zero_rows) {
    float* zr = zero_rows + (size_t)net * m.P;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m.P; i += gridDim.x * blockDim.x) zr[i] = 0.f;
  }
  if (blockIdx.x != 0) return;
  if (threadIdx.x == 0) {
    if (zero_acc) zero_acc[net] = 0.f;
    if (zero_acc2) zero_acc2[net] = 0.f;
    if (net == 0 && zero_cursors) { zero_cursors[0] = 0; zero_cursors[1] = 0; }
    // the loss buffer of THIS call: map_update_kernel reads the pointer from device memory, so a
    // cached CUDA graph of the step does not bake the caller's output buffer into its nodes
    if (net == 0 && loss_slot) *loss_slot = out_loss;
  }
  if (threadIdx.x < 32) prep_one(m, params + (size_t)net * m.P, derived + (size_t)net * kDerivedStride, threadIdx.x);
}

// =============================================================================
// encode: (x rows) -> feat [n_net, B, Fp]  (models.py:216-252)
// work item = (row, unit); unit = one x column, one (dim,degree) sin/cos pair,
// one seasonal sin/cos pair, or one interaction column.
// =============================================================================
constexpr int kEncRows = 32;

// X3 (T = bf16): accurate trig, every feature stored as its three bf16 planes, row layout
// [plane 0: Fp | plane 1: Fp | plane 2: Fp] (the split A operand of the bf16x3 Dense_0 GEMM).
template <typename T, bool X3 = false>
__global__ void encode_kernel(const __grid_constant__ DevModel m, const float* __restrict__ derived,
                              const float* __restrict__ x, const int32_t* __restrict__ idx,
                              int64_t idx_stride, int B, T* __restrict__ feat) {
  constexpr bool FAST = FastMath<T>::value && !X3;
  extern __shared__ float tile[];  // [kEncRows][Fp+1]
  pdl_enter(derived, x, idx, feat);
  const int net = blockIdx.y;
  const int row0 = blockIdx.x * kEncRows;
  const float* dv = derived + (size_t)net * kDerivedStride;
  const int ld = m.Fp + 1;
  for (int e = threadIdx.x; e < kEncRows * ld; e += blockDim.x) tile[e] = 0.f;
  __syncthreads();
  const int U = num_units(m);
  const float two_pi = 6.283185307179586f;
  for (int w = threadIdx.x; w < kEncRows * U; w += blockDim.x) {
    const int u = w / kEncRows, r = w % kEncRows;
    const int b = row0 + r;
    if (b >= B) continue;
    const UnitInfo ui = decode_unit(m, u);
    const float* xr = row_ptr(x, idx, idx_stride, net, b, m.D);
    float* trow = tile + r * ld;
    if (ui.kind == 0) {
      float sx = xr[ui.a] / dv[kDvDenom + ui.a];
      trow[m.col_x + ui.a] = sx * dv[kDvSX];
    } else if (ui.kind == 1) {
      const int i = ui.a, d = ui.b, deg = m.fourier_deg[i];
      float sx = xr[i] / dv[kDvDenom + i];
      float c = two_pi * (float)(1 << d);
      float sn, cs;
      if (FAST) sincos_reduced(c * sx, &sn, &cs); else sincosf(c * sx, &sn, &cs);
      const float den = (float)(d + 1), s = dv[kDvSFourier + i];
      trow[m.fourier_col[i] + d] = (cs / den) * s;
      trow[m.fourier_col[i] + deg + d] = (sn / den) * s;
    } else if (ui.kind == 2) {
      const int k = ui.a;
      float sn, cs;
      if (FAST) sincos_reduced(m.seasonal_w[k] * xr[0], &sn, &cs); else sincosf(m.seasonal_w[k] * xr[0], &sn, &cs);
      const float s = dv[kDvSSeas], hk = m.seasonal_h[k];
      trow[m.col_seasonal + k] = (cs / hk) * s;
      trow[m.col_seasonal + m.n_seasonal + k] = (sn / hk) * s;
    } else {
      const int j = ui.a;
      float sa = xr[m.inter_a[j]] / dv[kDvDenom + m.inter_a[j]];
      float sb = xr[m.inter_b[j]] / dv[kDvDenom + m.inter_b[j]];
      trow[m.col_inter + j] = (sa * sb) * dv[kDvSInter];
    }
  }
  // constant-one feature in the first pad column (ignored by the forward GEMM: the staged kernel rows
  // >= F are zero): row F of the Dense_0 wgrad GEMM then is the Dense_0 bias gradient (bnf_tc.cu)
  if (m.F < m.Fp) for (int r = threadIdx.x; r < kEncRows; r += blockDim.x) tile[r * ld + m.F] = 1.f;
  __syncthreads();
  const int rows = min(kEncRows, B - row0);
  if constexpr (X3) {
    __nv_bfloat16* out = feat + ((size_t)net * B + row0) * 3 * m.Fp;
    for (int e = threadIdx.x; e < rows * m.Fp; e += blockDim.x) {
      int r = e / m.Fp, c = e % m.Fp;
      __nv_bfloat16* o = out + (size_t)r * 3 * m.Fp + c;
      split3_one(tile[r * ld + c], o, o + m.Fp, o + 2 * m.Fp);
    }
  } else {
    T* out = feat + ((size_t)net * B + row0) * m.Fp;
    for (int e = threadIdx.x; e < rows * m.Fp; e += blockDim.x) {
      int r = e / m.Fp, c = e % m.Fp;
      out[e] = from_f<T>(tile[r * ld + c]);
    }
  }
}

// bf16 path: the same features with the per-item bookkeeping hoisted out of the inner loop.
// ncu on the generic kernel above (chickenpox shape): 236 instructions per (row, unit) item,
// 84 % issue-slot utilisation -- it is instruction-bound on unit decoding, IEEE divisions and
// the element-wise copy-out.  Here a block first builds a per-unit constant table (column
// indices, argument multiplier, output scales with the 1/(d+1), 1/h divisions folded in) and
// the scaled inputs of its 64 rows in shared memory; an item is then table lookup -> one
// sincos -> two bf16 stores into a padded tile that leaves as 16-byte vectors.
struct EncUnit { float mult, k0, k1; int kind, dim, dim2, c0, c1; };
constexpr int kEncFastRows = 64;

// X3: the bf16x3 variant -- accurate sincosf (the f32 arguments reach 1e3..1e5 rad), the tile is kept in
// f32 and leaves as three bf16 planes [plane 0: Fp | plane 1: Fp | plane 2: Fp] per row.
template <bool X3>
__global__ void __launch_bounds__(256)
encode_fast_kernel(const __grid_constant__ DevModel m, const float* __restrict__ derived,
                   const float* __restrict__ x, const int32_t* __restrict__ idx, int64_t idx_stride,
                   int B, __nv_bfloat16* __restrict__ feat, int U) {
  extern __shared__ __align__(16) uint8_t esm[];
  constexpr int R = kEncFastRows;
  constexpr int ES = X3 ? 4 : 2;                         // bytes per tile element
  pdl_enter(derived, x, idx, feat);
  const int ldt = m.Fp * ES + 16;                        // bytes per tile row (16-byte aligned, 4-way banks)
  uint8_t* tile = esm;                                   // [R][ldt] features
  float* sxs = reinterpret_cast<float*>(esm + R * ldt);  // [R][kMaxD+1] x/denom, slot D = raw time
  EncUnit* tab = reinterpret_cast<EncUnit*>(sxs + R * (kMaxD + 1));
  const int net = blockIdx.y, row0 = blockIdx.x * R, tid = threadIdx.x;
  const float* dv = derived + (size_t)net * kDerivedStride;
  for (int e = tid; e < R * ldt / 16; e += 256) reinterpret_cast<uint4*>(tile)[e] = make_uint4(0, 0, 0, 0);
  for (int u = tid; u < U; u += 256) {
    const UnitInfo ui = decode_unit(m, u);
    EncUnit t;
    t.kind = ui.kind; t.dim = 0; t.dim2 = 0; t.c0 = 0; t.c1 = 0; t.mult = 0.f; t.k0 = 0.f; t.k1 = 0.f;
    if (ui.kind == 0) {
      t.dim = ui.a; t.c0 = m.col_x + ui.a; t.k0 = dv[kDvSX];
    } else if (ui.kind == 1) {
      const int i = ui.a, d = ui.b;
      t.dim = i; t.mult = 6.283185307179586f * (float)(1 << d);
      t.c0 = m.fourier_col[i] + d; t.c1 = t.c0 + m.fourier_deg[i];
      // X3 keeps the reference's operation order (feature / (d+1)) * scale, models.py:87,250
      t.k0 = X3 ? (float)(d + 1) : dv[kDvSFourier + i] / (float)(d + 1);
      t.k1 = dv[kDvSFourier + i];
    } else if (ui.kind == 2) {
      const int k = ui.a;
      t.dim = m.D; t.mult = m.seasonal_w[k];
      t.c0 = m.col_seasonal + k; t.c1 = t.c0 + m.n_seasonal;
      t.k0 = X3 ? m.seasonal_h[k] : dv[kDvSSeas] / m.seasonal_h[k];
      t.k1 = dv[kDvSSeas];
    } else {
      t.dim = m.inter_a[ui.a]; t.dim2 = m.inter_b[ui.a]; t.c0 = m.col_inter + ui.a; t.k0 = dv[kDvSInter];
    }
    tab[u] = t;
  }
  for (int e = tid; e < R * m.D; e += 256) {
    const int r = e / m.D, i = e - r * m.D;
    const int b = min(row0 + r, B - 1);
    const float xv = row_ptr(x, idx, idx_stride, net, b, m.D)[i];
    sxs[r * (kMaxD + 1) + i] = xv / dv[kDvDenom + i];
    if (i == 0) sxs[r * (kMaxD + 1) + m.D] = xv;
  }
  __syncthreads();
  auto put = [&](int r, int c, float v) {
    if constexpr (X3) reinterpret_cast<float*>(tile + r * ldt)[c] = v;
    else reinterpret_cast<__nv_bfloat16*>(tile + r * ldt)[c] = __float2bfloat16_rn(v);
  };
  for (int w = tid; w < R * U; w += 256) {
    const int u = w / R, r = w % R;                      // a warp shares one unit: uniform table reads
    const EncUnit t = tab[u];
    const float* sr = sxs + r * (kMaxD + 1);
    if (t.kind == 0) {
      put(r, t.c0, sr[t.dim] * t.k0);
    } else if (t.kind == 3) {
      put(r, t.c0, (sr[t.dim] * sr[t.dim2]) * t.k0);
    } else {
      float sn, cs;
      if constexpr (X3) {
        sincosf(t.mult * sr[t.dim], &sn, &cs);
        put(r, t.c0, (cs / t.k0) * t.k1);
        put(r, t.c1, (sn / t.k0) * t.k1);
      } else {
        sincos_reduced(t.mult * sr[t.dim], &sn, &cs);
        put(r, t.c0, cs * t.k0);
        put(r, t.c1, sn * t.k0);
      }
    }
  }
  if (m.F < m.Fp) for (int r = tid; r < R; r += 256) put(r, m.F, 1.f);     // constant-one feature (see encode_kernel)
  __syncthreads();
  const int rows = min(R, B - row0), chunks = m.Fp / 8;
  if constexpr (X3) {
    // 8 features of a row -> 16 bytes in each of the three planes
    for (int e = tid; e < rows * chunks; e += 256) {
      const int r = e / chunks, j = e - r * chunks;
      const float4 f0 = *reinterpret_cast<const float4*>(tile + r * ldt + j * 32);
      const float4 f1 = *reinterpret_cast<const float4*>(tile + r * ldt + j * 32 + 16);
      alignas(16) uint32_t pk[3][4];
      split3_pair(f0.x, f0.y, &pk[0][0], &pk[1][0], &pk[2][0]);
      split3_pair(f0.z, f0.w, &pk[0][1], &pk[1][1], &pk[2][1]);
      split3_pair(f1.x, f1.y, &pk[0][2], &pk[1][2], &pk[2][2]);
      split3_pair(f1.z, f1.w, &pk[0][3], &pk[1][3], &pk[2][3]);
      __nv_bfloat16* o = feat + ((size_t)net * B + row0 + r) * 3 * m.Fp + j * 8;
#pragma unroll
      for (int pl = 0; pl < 3; ++pl) *reinterpret_cast<uint4*>(o + pl * m.Fp) = *reinterpret_cast<const uint4*>(pk[pl]);
    }
  } else {
    uint4* out = reinterpret_cast<uint4*>(feat + ((size_t)net * B + row0) * m.Fp);
    for (int e = tid; e < rows * chunks; e += 256) {
      const int r = e / chunks, j = e - r * chunks;
      out[e] = *reinterpret_cast<const uint4*>(tile + r * ldt + j * 16);
    }
  }
}

// encode backward: dfeat [n_net,B,Fp] f32 -> grads of
// feature_inv_sp_scale{g} and log_scale_adjustment, accumulated into grad[n_net,P].
// A warp owns whole units (one x column / one sin-cos pair / one interaction column): its
// lanes stride over the block's rows and keep the three partial sums in registers, so there is
// one warp reduction per (unit, block) instead of one per 32 rows; rows per block are chosen at
// launch so the grid is a whole number of waves.
constexpr int kEncBwdMaxRows = 1024;
template <bool FAST>
__global__ void __launch_bounds__(256)
encode_bwd_kernel(const __grid_constant__ DevModel m, const float* __restrict__ params,
                  const float* __restrict__ derived, const float* __restrict__ x,
                  const int32_t* __restrict__ idx, int64_t idx_stride, int B, int R /* rows per block */,
                  const float* __restrict__ dfeat, int64_t g_row, int64_t g_col /* element strides */,
                  float* __restrict__ grad) {
  __shared__ float acc[kMaxD + kMaxD + 3];  // [0,D): lsa ; D + {0:x,1:seasonal,2:inter, 3+i: fourier_i}
  pdl_enter(params, derived, x, idx, dfeat, grad);
  const int net = blockIdx.y;
  const int row0 = blockIdx.x * R, row1 = min(B, row0 + R);
  const float* dv = derived + (size_t)net * kDerivedStride;
  const int nacc = m.D + 3 + m.D;
  for (int e = threadIdx.x; e < nacc; e += blockDim.x) acc[e] = 0.f;
  __syncthreads();
  const int U = num_units(m);
  const float two_pi = 6.283185307179586f;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int u = warp; u < U; u += 8) {
    const UnitInfo ui = decode_unit(m, u);
    int slot = 0, dim_a = 0, dim_b = -1;
    if (ui.kind == 0) { slot = 0; dim_a = ui.a; }
    else if (ui.kind == 1) { slot = 3 + ui.a; dim_a = ui.a; }
    else if (ui.kind == 2) { slot = 1; dim_a = -1; }
    else { slot = 2; dim_a = m.inter_a[ui.a]; dim_b = m.inter_b[ui.a]; }
    float gs = 0.f, gl_a = 0.f, gl_b = 0.f;
#pragma unroll 2
    for (int b = row0 + lane; b < row1; b += 32) {
      const float* xr = row_ptr(x, idx, idx_stride, net, b, m.D);
      const float* g = dfeat + (size_t)net * B * m.Fp + (size_t)b * g_row;
      if (ui.kind == 0) {
        float sx = xr[ui.a] / dv[kDvDenom + ui.a];
        float G = g[(m.col_x + ui.a) * g_col];
        gs = fmaf(G, sx, gs);
        gl_a = fmaf(dv[kDvSX] * G, -sx, gl_a);
      } else if (ui.kind == 1) {
        const int i = ui.a, d = ui.b, deg = m.fourier_deg[i];
        float sx = xr[i] / dv[kDvDenom + i];
        float c = two_pi * (float)(1 << d);
        float sn, cs;
        if (FAST) sincos_reduced(c * sx, &sn, &cs); else sincosf(c * sx, &sn, &cs);
        const float den = (float)(d + 1);
        float Gc = g[(m.fourier_col[i] + d) * g_col], Gs = g[(m.fourier_col[i] + deg + d) * g_col];
        gs += Gc * (cs / den) + Gs * (sn / den);
        float dsx = dv[kDvSFourier + i] * (c / den) * (-sn * Gc + cs * Gs);
        gl_a = fmaf(dsx, -sx, gl_a);
      } else if (ui.kind == 2) {
        const int k = ui.a;
        float sn, cs;
        if (FAST) sincos_reduced(m.seasonal_w[k] * xr[0], &sn, &cs); else sincosf(m.seasonal_w[k] * xr[0], &sn, &cs);
        float hk = m.seasonal_h[k];
        gs += g[(m.col_seasonal + k) * g_col] * (cs / hk) + g[(m.col_seasonal + m.n_seasonal + k) * g_col] * (sn / hk);
      } else {
        const int j = ui.a;
        float sa = xr[dim_a] / dv[kDvDenom + dim_a];
        float sb = xr[dim_b] / dv[kDvDenom + dim_b];
        float G = g[(m.col_inter + j) * g_col];
        gs = fmaf(G, sa * sb, gs);
        // d(sa*sb)/d lsa_a = -sa*sb, same for b
        gl_a = fmaf(dv[kDvSInter] * G, -sa * sb, gl_a);
      }
    }
    if (ui.kind == 3) gl_b = gl_a;
    gs = warp_sum(gs);
    gl_a = warp_sum(gl_a);
    gl_b = warp_sum(gl_b);
    if (lane == 0) {
      atomicAdd(&acc[m.D + slot], gs);
      if (dim_a >= 0) atomicAdd(&acc[dim_a], gl_a);
      if (dim_b >= 0) atomicAdd(&acc[dim_b], gl_b);
    }
  }
  __syncthreads();
  const float* p = params + (size_t)net * m.P;
  float* gr = grad + (size_t)net * m.P;
  for (int e = threadIdx.x; e < nacc; e += blockDim.x) {
    float v = acc[e];
    if (e < m.D) { atomicAdd(&gr[m.off_lsa + e], v); continue; }
    int slot = e - m.D, off;
    if (slot == 0) off = m.off_scale_x;
    else if (slot == 1) off = m.off_scale_seasonal;
    else if (slot == 2) off = m.off_scale_inter;
    else off = m.fourier_scale_off[slot - 3];
    if (off >= 0) atomicAdd(&gr[off], v * sigmoid_f(p[off]));  // d softplus = sigmoid
  }
}

// =============================================================================
// fp32 SIMT GEMM, 64x64x16 tiles, 256 threads, 4x4 per thread, fused epilogue.
// Operand element (m,k): A_KC ? A[m*lda+k] : A[k*lda+m]; (k,n): B_KC ? B[n*ldb+k] : B[k*ldb+n]
// =============================================================================
template <typename TA, typename TB, bool A_KC, bool B_KC, typename Epi>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const TA* __restrict__ A, size_t a_batch, int lda, const TB* __restrict__ Bm,
                 size_t b_batch, int ldb, int M, int N, int K, Epi epi) {
  __shared__ float As[16][68];
  __shared__ float Bs[16][68];
  pdl_enter(A, Bm);   // the epilogue functor's pointers are not __restrict__: ordinary loads
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64, net = blockIdx.z;
  A += (size_t)net * a_batch;
  Bm += (size_t)net * b_batch;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = tid + i * 256;
      int mm, kk;
      if (A_KC) { mm = e >> 4; kk = e & 15; } else { kk = e >> 6; mm = e & 63; }
      int gm = m0 + mm, gk = k0 + kk;
      float v = 0.f;
      if (gm < M && gk < K) v = to_f<TA>(A_KC ? A[(size_t)gm * lda + gk] : A[(size_t)gk * lda + gm]);
      As[kk][mm] = v;
      int nn;
      if (B_KC) { nn = e >> 4; kk = e & 15; } else { kk = e >> 6; nn = e & 63; }
      int gn = n0 + nn;
      gk = k0 + kk;
      v = 0.f;
      if (gn < N && gk < K) v = to_f<TB>(B_KC ? Bm[(size_t)gn * ldb + gk] : Bm[(size_t)gk * ldb + gn]);
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gm = m0 + ty * 4 + i, gn = n0 + tx * 4 + j;
      if (gm < M && gn < N) epi(net, gm, gn, acc[i][j]);
    }
}

// forward layer epilogue (models.py:264-268): z = s*(acc/sqrt(fan) + b); h = act(z)
template <typename T>
struct EpiFwd {
  const float* params; const float* derived; int P, off_bias, layer; float isf;
  T* z; T* h; size_t batch; int ld;
  __device__ void operator()(int net, int m, int n, float acc) const {
    const float* dv = derived + (size_t)net * kDerivedStride;
    float u = acc * isf + params[(size_t)net * P + off_bias + n];
    float zz = dv[kDvSLayer + layer] * u;
    size_t o = (size_t)net * batch + (size_t)m * ld + n;
    if (z) z[o] = from_f<T>(zz);
    h[o] = from_f<T>(act_sel<FastMath<T>::value>(zz, dv[kDvActW]));
  }
};
template <typename T>
struct EpiStoreScaled {  // dgrad: out = acc * isf
  T* out; size_t batch; int ld; float isf;
  __device__ void operator()(int net, int m, int n, float acc) const {
    out[(size_t)net * batch + (size_t)m * ld + n] = from_f<T>(acc * isf);
  }
};
struct EpiWgrad {  // grad[net*P + off + m*N + n] += acc*isf
  float* grad; int P, off, N; float isf;
  __device__ void operator()(int net, int m, int n, float acc) const {
    grad[(size_t)net * P + off + (size_t)m * N + n] += acc * isf;
  }
};

template <typename T>
void launch_fwd_layer_simt(const DevModel& m, int layer, const float* params, const float* derived,
                           const T* a_in, int K, int lda, T* z, T* h, int n_net, int B,
                           cudaStream_t st) {
  EpiFwd<T> epi{params, derived, m.P, m.off_bias[layer], layer,
                layer == 0 ? m.inv_sqrt_F : m.inv_sqrt_W, z, h, (size_t)B * m.W, m.W};
  dim3 grid((m.W + 63) / 64, (B + 63) / 64, n_net);
  BNF_PROF("gemm_simt_fwd", st);
  launch_k(gemm_simt_kernel<T, float, true, false, EpiFwd<T>>, grid, dim3(256), 0, st,
           a_in, (size_t)B * lda, lda, params + m.off_kernel[layer], (size_t)m.P, m.W, B, m.W, K, epi);
}
template <typename T, typename TO>
void launch_dgrad_simt(const DevModel& m, int layer, const float* params, const T* dU, TO* out,
                       int Kout, int ld_out, int n_net, int B, cudaStream_t st) {
  // out[b,k] = isf * sum_n dU[b,n] * K[k,n]
  EpiStoreScaled<TO> epi{out, (size_t)B * ld_out, ld_out, layer == 0 ? m.inv_sqrt_F : m.inv_sqrt_W};
  dim3 grid((Kout + 63) / 64, (B + 63) / 64, n_net);
  BNF_PROF("gemm_simt_dgrad", st);
  launch_k(gemm_simt_kernel<T, float, true, true, EpiStoreScaled<TO>>, grid, dim3(256), 0, st,
           dU, (size_t)B * m.W, m.W, params + m.off_kernel[layer], (size_t)m.P, m.W, B, Kout, m.W, epi);
}
template <typename T>
void launch_wgrad_simt(const DevModel& m, int layer, const T* a_in, int Kin, int lda, const T* dU,
                       float* grad, int n_net, int B, cudaStream_t st) {
  EpiWgrad epi{grad, m.P, m.off_kernel[layer], m.W, layer == 0 ? m.inv_sqrt_F : m.inv_sqrt_W};
  dim3 grid((m.W + 63) / 64, (Kin + 63) / 64, n_net);
  BNF_PROF("gemm_simt_wgrad", st);
  launch_k(gemm_simt_kernel<T, T, false, false, EpiWgrad>, grid, dim3(256), 0, st,
           a_in, (size_t)B * lda, lda, dU, (size_t)B * m.W, m.W, Kin, m.W, B, epi);
}

// =============================================================================
// head: o = s_out*(h.Ko/sqrt(W) + bo) ; likelihood ; r = dlogp/do  (models.py:269-273,157-191)
// one warp per row.  Accumulates loglik and the scalar-head gradients.
// =============================================================================
constexpr int kHeadRows = 256;
// NP = 3 (T = bf16): h rows hold three bf16 planes [W | W | W] whose sum is the f32 value (bf16x3 mode)
template <typename T, int NP = 1>
__global__ void __launch_bounds__(256)
head_kernel(const __grid_constant__ DevModel m, const float* __restrict__ params,
            const float* __restrict__ derived, const T* __restrict__ h, const float* __restrict__ y_all,
            const int32_t* __restrict__ idx, int64_t idx_stride, int B, float* __restrict__ out_loc,
            float* __restrict__ opre_out, float* __restrict__ r_out, float* __restrict__ ll,
            float* __restrict__ grad) {
  pdl_enter(params, derived, h, y_all, idx, out_loc, opre_out, r_out, ll, grad);
  const int net = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* p = params + (size_t)net * m.P;
  const float* dv = derived + (size_t)net * kDerivedStride;
  const float* Ko = p + m.off_kernel[m.L];
  const float bo = p[m.off_bias[m.L]];
  const float s_out = dv[kDvSOut];
  float a_ll = 0.f, a_g0 = 0.f, a_g1 = 0.f, a_g2 = 0.f, a_gs = 0.f, a_gb = 0.f;
  // Each warp takes 32 rows at a time: the row dot products are computed cooperatively
  // (coalesced 16-byte loads), lane j keeps row j's result, then all 32 lanes run the
  // per-row likelihood math in parallel and write r / o_pre coalesced.
  const int blk0 = blockIdx.x * kHeadRows;
  for (int base = blk0 + warp * 32; base < min(B, blk0 + kHeadRows); base += 8 * 32) {
    float mydot = 0.f;
    constexpr int VEC = 16 / sizeof(T);
    for (int j0 = 0; j0 < 32; j0 += 4) {       // 4 rows in flight: independent loads, then reduce
      if (base + j0 >= B) break;                // warp-uniform
      float dot[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int b = base + j0 + u;
        if (b < B) {
          const T* hr = h + ((size_t)net * B + b) * NP * m.W;
          if (m.W % VEC == 0) {
            for (int n = lane * VEC; n < m.W; n += 32 * VEC) {
              float hf[VEC];
#pragma unroll
              for (int pl = 0; pl < NP; ++pl) {
                alignas(16) T hv[VEC];
                *reinterpret_cast<uint4*>(hv) = *reinterpret_cast<const uint4*>(hr + pl * m.W + n);
#pragma unroll
                for (int k = 0; k < VEC; ++k) hf[k] = pl == 0 ? to_f<T>(hv[k]) : hf[k] + to_f<T>(hv[k]);
              }
#pragma unroll
              for (int k = 0; k < VEC; ++k) dot[u] = fmaf(hf[k], Ko[n + k], dot[u]);
            }
          } else {
            for (int n = lane; n < m.W; n += 32) {
              float hf = to_f<T>(hr[n]);
              for (int pl = 1; pl < NP; ++pl) hf += to_f<T>(hr[pl * m.W + n]);
              dot[u] = fmaf(hf, Ko[n], dot[u]);
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float d = warp_sum(dot[u]);
        if (lane == j0 + u) mydot = d;
      }
    }
    const int b = base + lane;
    if (b < B) {
      const float opre = mydot * m.inv_sqrt_W + bo;
      const float o = s_out * opre;
      if (out_loc) out_loc[(size_t)net * B + b] = o;
      if (r_out) {
        int64_t row = idx ? (int64_t)idx[(int64_t)net * idx_stride + b] : (int64_t)b;
        const float yv = y_all[row];
        float logp, r;
        if (m.likelihood == BNF_NORMAL) {
          const float sg = dv[kDvSigma];
          const float d = yv / sg - o / sg;              // TFP Normal._log_prob form
          logp = -0.5f * d * d - (0.9189385332046727f + logf(sg));
          r = d / sg;
          a_g0 += (d * d - 1.f) / sg;                     // d/dsigma, chain to lns at the end
        } else {
          const float mean = softplus_f(o);
          const float shp = dv[kDvShape];
          const float rc = 1.f / shp;                     // total_count
          const float lg = -logf(shp) - logf(mean);       // logits, models.py:173-175
          const float sig_l = sigmoid_f(lg);
          float nb = rc * log_sigmoid_f(-lg) + yv * log_sigmoid_f(lg)
                     - (lgammaf(1.f + yv) + lgammaf(rc) - lgammaf(1.f + yv + rc)) - logf(rc + yv);
          float dnb_dl = yv * (1.f - sig_l) - rc * sig_l;
          float dnb_dr = log_sigmoid_f(-lg) - digamma_f(rc) + digamma_f(1.f + yv + rc) - 1.f / (rc + yv);
          float wnb = 1.f;                                // d logp / d nb
          logp = nb;
          if (m.likelihood == BNF_ZINB) {
            const float pi = dv[kDvPi];
            if (yv == 0.f) {
              const float A = (1.f - pi) * expf(nb), tot = A + pi;
              logp = logf(tot);
              wnb = A / tot;
              a_g2 += (1.f - expf(nb)) / tot;             // d/dpi
            } else {
              logp = log1pf(-pi) + nb;
              a_g2 += -1.f / (1.f - pi);
            }
          }
          r = wnb * dnb_dl * (-sigmoid_f(o) / mean);
          a_g1 += wnb * (dnb_dl * (-1.f / shp) + dnb_dr * (-1.f / (shp * shp)));  // d/dshape
        }
        a_ll += logp;
        a_gs += r * opre;
        a_gb += r * s_out;
        r_out[(size_t)net * B + b] = r;
        opre_out[(size_t)net * B + b] = opre;
      }
    }
  }
  a_ll = warp_sum(a_ll); a_g0 = warp_sum(a_g0); a_g1 = warp_sum(a_g1);
  a_g2 = warp_sum(a_g2); a_gs = warp_sum(a_gs); a_gb = warp_sum(a_gb);
  // lane 0 of each warp holds its partial sums: combine the 8 warps in shared memory so the
  // block issues one atomic per quantity (the targets are 6 addresses per network)
  __shared__ float hred[6][8];
  if (lane == 0) {
    hred[0][warp] = a_ll; hred[1][warp] = a_g0; hred[2][warp] = a_g1;
    hred[3][warp] = a_g2; hred[4][warp] = a_gs; hred[5][warp] = a_gb;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    a_ll = a_g0 = a_g1 = a_g2 = a_gs = a_gb = 0.f;
    for (int i = 0; i < 8; ++i) {
      a_ll += hred[0][i]; a_g0 += hred[1][i]; a_g1 += hred[2][i];
      a_g2 += hred[3][i]; a_gs += hred[4][i]; a_gb += hred[5][i];
    }
  }
  if (r_out && threadIdx.x == 0) atomicAdd(&ll[net], a_ll);
  if (r_out && grad && threadIdx.x == 0) {
    float* g = grad + (size_t)net * m.P;
    if (m.likelihood == BNF_NORMAL) {
      atomicAdd(&g[0], a_g0 * expf(p[0]));                // dsigma/dlns = exp(lns)
    } else {
      atomicAdd(&g[1], a_g1 * sigmoid_f(p[1]));           // dshape/dp1 = sigmoid
      if (m.likelihood == BNF_ZINB) {
        const float pi = dv[kDvPi];
        atomicAdd(&g[2], a_g2 * pi * (1.f - pi));
      }
    }
    atomicAdd(&g[m.off_out_scale], a_gs * sigmoid_f(p[m.off_out_scale]));
    atomicAdd(&g[m.off_bias[m.L]], a_gb);
  }
}

// =============================================================================
// activation backward + column reductions for hidden layer `layer`.
//   IS_HEAD : dh[b,n] = r[b]*s_out*Ko[n]/sqrt(W) (rank-1), also dKo
//   else    : dh read from `dh_in` (output of the dgrad GEMM), overwritten by dU
// dz = dh*act'(z); dU = s_l*dz; g_actw += dh*(elu-tanh); g_ls += dz*u; g_b[n] += dU
// =============================================================================
constexpr int kActRows = 64;
template <typename T, bool IS_HEAD>
__global__ void __launch_bounds__(128)
act_bwd_kernel(const __grid_constant__ DevModel m, int layer, const float* __restrict__ params,
               const float* __restrict__ derived, const T* __restrict__ z, const T* __restrict__ h,
               const float* __restrict__ r, T* __restrict__ dU /* in: dh (unless head), out: dU */,
               int B, float* __restrict__ grad) {
  __shared__ float red[2][4];
  pdl_enter(params, derived, z, h, r, dU, grad);
  const int net = blockIdx.z;
  const int n = blockIdx.x * 128 + threadIdx.x;
  const int b0 = blockIdx.y * kActRows, b1 = min(B, b0 + kActRows);
  const float* p = params + (size_t)net * m.P;
  const float* dv = derived + (size_t)net * kDerivedStride;
  const float w = dv[kDvActW], s_l = dv[kDvSLayer + layer];
  float g_w = 0.f, g_s = 0.f, g_b = 0.f, g_ko = 0.f;
  if (n < m.W) {
    float head_c = 0.f;
    if (IS_HEAD) head_c = dv[kDvSOut] * m.inv_sqrt_W * p[m.off_kernel[m.L] + n];
    for (int b = b0; b < b1; ++b) {
      const size_t o = ((size_t)net * B + b) * m.W + n;
      const float zz = to_f<T>(z[o]);
      float dh;
      if (IS_HEAD) {
        const float rb = r[(size_t)net * B + b];
        dh = rb * head_c;
        g_ko += rb * to_f<T>(h[o]);
      } else {
        dh = to_f<T>(dU[o]);
      }
      float diff;
      const float da = act_grad_sel<FastMath<T>::value>(zz, w, &diff);
      const float dz = dh * da;
      g_w += dh * diff;
      g_s += dz * zz;                       // = dz*u*s_l ; divided by s_l below
      const float du = dz * s_l;
      g_b += du;
      dU[o] = from_f<T>(du);
    }
    float* g = grad + (size_t)net * m.P;
    atomicAdd(&g[m.off_bias[layer] + n], g_b);
    if (IS_HEAD) atomicAdd(&g[m.off_kernel[m.L] + n], g_ko * dv[kDvSOut] * m.inv_sqrt_W);
  }
  g_w = warp_sum(g_w);
  g_s = warp_sum(g_s);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = g_w; red[1][warp] = g_s; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float tw = red[0][0] + red[0][1] + red[0][2] + red[0][3];
    float ts = red[1][0] + red[1][1] + red[1][2] + red[1][3];
    float* g = grad + (size_t)net * m.P;
    atomicAdd(&g[m.off_actw], tw * w * (1.f - w));
    atomicAdd(&g[m.off_layer_scale[layer]], (ts / s_l) * sigmoid_f(p[m.off_layer_scale[layer]]));
  }
}

// Vectorised variant (16-byte loads/stores, 8 bf16 or 4 f32 columns per thread),
// used when W is a multiple of the vector width and 256 % (W/VEC) == 0 so that a
// thread keeps the same column group for every row it visits (column sums stay in
// registers).  Same math as act_bwd_kernel.
constexpr int kActVecRows = 256;
template <typename T, bool IS_HEAD>
__global__ void __launch_bounds__(256)
act_bwd_vec_kernel(const __grid_constant__ DevModel m, int layer, const float* __restrict__ params,
                   const float* __restrict__ derived, const T* __restrict__ z, const T* __restrict__ h,
                   const float* __restrict__ r, T* __restrict__ dU, int B, float* __restrict__ grad) {
  constexpr int VEC = 16 / sizeof(T);
  constexpr bool FAST = FastMath<T>::value;
  __shared__ float red[2][8];
  extern __shared__ float colsum[];            // [W] bias grads (+ [W] Dense_L kernel grads at the head)
  pdl_enter(params, derived, z, h, r, dU, grad);
  const int net = blockIdx.y;
  for (int i = threadIdx.x; i < (IS_HEAD ? 2 : 1) * m.W; i += blockDim.x) colsum[i] = 0.f;
  __syncthreads();
  const int G = m.W / VEC;                     // column groups per row
  const int cg = threadIdx.x % G;
  const int rstep = 256 / G;                   // rows covered per pass
  const int b0 = blockIdx.x * kActVecRows, b1 = min(B, b0 + kActVecRows);
  const float* p = params + (size_t)net * m.P;
  const float* dv = derived + (size_t)net * kDerivedStride;
  const float w = dv[kDvActW], s_l = dv[kDvSLayer + layer];
  float g_w = 0.f, g_s = 0.f, g_b[VEC], g_ko[VEC], head_c[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    g_b[k] = 0.f; g_ko[k] = 0.f;
    head_c[k] = IS_HEAD ? dv[kDvSOut] * m.inv_sqrt_W * p[m.off_kernel[m.L] + cg * VEC + k] : 0.f;
  }
#pragma unroll 4
  for (int b = b0 + threadIdx.x / G; b < b1; b += rstep) {
    const size_t o = ((size_t)net * B + b) * m.W + (size_t)cg * VEC;
    alignas(16) T zv[VEC];
    alignas(16) T dv_in[VEC];
    alignas(16) T hv[VEC];
    alignas(16) T out[VEC];
    *reinterpret_cast<uint4*>(zv) = *reinterpret_cast<const uint4*>(z + o);
    float rb = 0.f;
    if (IS_HEAD) {
      rb = r[(size_t)net * B + b];
      *reinterpret_cast<uint4*>(hv) = *reinterpret_cast<const uint4*>(h + o);
    } else {
      *reinterpret_cast<uint4*>(dv_in) = *reinterpret_cast<const uint4*>(dU + o);
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const float zz = to_f<T>(zv[k]);
      float dh;
      if (IS_HEAD) { dh = rb * head_c[k]; g_ko[k] += rb * to_f<T>(hv[k]); }
      else dh = to_f<T>(dv_in[k]);
      float diff;
      const float da = act_grad_sel<FAST>(zz, w, &diff);
      const float dz = dh * da;
      g_w += dh * diff;
      g_s += dz * zz;
      const float du = dz * s_l;
      g_b[k] += du;
      out[k] = from_f<T>(du);
    }
    *reinterpret_cast<uint4*>(dU + o) = *reinterpret_cast<const uint4*>(out);
  }
  float* g = grad + (size_t)net * m.P;
  // block-level column sums in shared memory first: one global atomic per column per block
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    atomicAdd(&colsum[cg * VEC + k], g_b[k]);
    if (IS_HEAD) atomicAdd(&colsum[m.W + cg * VEC + k], g_ko[k]);
  }
  g_w = warp_sum(g_w);
  g_s = warp_sum(g_s);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = g_w; red[1][warp] = g_s; }
  __syncthreads();
  for (int i = threadIdx.x; i < m.W; i += blockDim.x) {
    atomicAdd(&g[m.off_bias[layer] + i], colsum[i]);
    if (IS_HEAD) atomicAdd(&g[m.off_kernel[m.L] + i], colsum[m.W + i] * dv[kDvSOut] * m.inv_sqrt_W);
  }
  if (threadIdx.x == 0) {
    float tw = 0.f, ts = 0.f;
    for (int i = 0; i < 8; ++i) { tw += red[0][i]; ts += red[1][i]; }
    atomicAdd(&g[m.off_actw], tw * w * (1.f - w));
    atomicAdd(&g[m.off_layer_scale[layer]], (ts / s_l) * sigmoid_f(p[m.off_layer_scale[layer]]));
  }
}

// =============================================================================
// head + activation backward of the last hidden layer in ONE kernel (training path).
// A block owns 256 rows: (A) each warp forms the h.Ko dots of 32 rows (coalesced 16-byte
// loads, 4 rows in flight), all lanes evaluate their row's likelihood and r = dlogp/do
// into shared memory; (B) the block runs the elementwise backward of the same rows
// (h re-read while it is still L2/L1-resident, z from HBM), so r never leaves the SM and
// the bias / Dense_L column sums are flushed once per block; the rows per block are picked at
// launch so that the grid is a whole number of waves of resident blocks (no ragged tail).
// Same math as head_kernel + act_bwd_vec_kernel<.., true>.
// =============================================================================
constexpr int kHeadFusedMaxRows = 512;   // rows per block are chosen at launch so the grid is whole waves
// X3 (T = bf16, bf16x3 mode): h rows hold three bf16 planes [W | W | W], dU rows two, z is f32
// (`z` then points to floats), accurate activation math.
template <typename T, bool X3 = false>
__global__ void __launch_bounds__(256)
head_fused_kernel(const __grid_constant__ DevModel m, const float* __restrict__ params,
                  const float* __restrict__ derived, const T* __restrict__ h, const void* __restrict__ z_any,
                  const float* __restrict__ y_all, const int32_t* __restrict__ idx, int64_t idx_stride,
                  int B, int R /* rows per block */, T* __restrict__ dU, float* __restrict__ ll,
                  float* __restrict__ grad) {
  constexpr int VEC = 16 / sizeof(T);
  constexpr bool FAST = FastMath<T>::value && !X3;
  constexpr int NP = X3 ? 3 : 1;
  const T* z = static_cast<const T*>(z_any);
  const float* zf = static_cast<const float*>(z_any);
  extern __shared__ float fsm[];
  float* rs = fsm;                              // [R] r = dlogp/do of this block's rows
  float* colsum = fsm + kHeadFusedMaxRows;      // [2W] bias / Dense_L kernel column sums
  float* kos = colsum + 2 * m.W;                // [W] Dense_L kernel
  __shared__ float hred[8][8];
  pdl_enter(params, derived, h, z, zf, y_all, idx, dU, ll, grad);
  const int net = blockIdx.y;
  const int b0 = blockIdx.x * R, b1 = min(B, b0 + R);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* p = params + (size_t)net * m.P;
  const float* dv = derived + (size_t)net * kDerivedStride;
  const float* Ko = p + m.off_kernel[m.L];
  const float bo = p[m.off_bias[m.L]], s_out = dv[kDvSOut];
  for (int i = threadIdx.x; i < 2 * m.W; i += 256) colsum[i] = 0.f;
  for (int i = threadIdx.x; i < m.W; i += 256) kos[i] = Ko[i];
  __syncthreads();
  // ---- (A) row dots + likelihood: groups of 32 rows round-robin over the 8 warps
  float a_ll = 0.f, a_g0 = 0.f, a_g1 = 0.f, a_g2 = 0.f, a_gs = 0.f, a_gb = 0.f;
  for (int base = b0 + warp * 32; base < b1; base += 256) {
    float mydot = 0.f;
    for (int j0 = 0; j0 < 32; j0 += 4) {
      if (base + j0 >= b1) break;               // warp-uniform
      float dot[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int b = base + j0 + u;
        if (b < b1) {
          const T* hr = h + ((size_t)net * B + b) * NP * m.W;
          for (int n = lane * VEC; n < m.W; n += 32 * VEC) {
            float hf[VEC];
            alignas(16) float kk[VEC];
#pragma unroll
            for (int pl = 0; pl < NP; ++pl) {
              alignas(16) T hv[VEC];
              *reinterpret_cast<uint4*>(hv) = *reinterpret_cast<const uint4*>(hr + pl * m.W + n);
#pragma unroll
              for (int k = 0; k < VEC; ++k) hf[k] = pl == 0 ? to_f<T>(hv[k]) : hf[k] + to_f<T>(hv[k]);
            }
#pragma unroll
            for (int k = 0; k < VEC; k += 4) *reinterpret_cast<float4*>(kk + k) = *reinterpret_cast<const float4*>(kos + n + k);
#pragma unroll
            for (int k = 0; k < VEC; ++k) dot[u] = fmaf(hf[k], kk[k], dot[u]);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float d = warp_sum(dot[u]);
        if (lane == j0 + u) mydot = d;
      }
    }
    const int b = base + lane;
    if (b < b1) {
      const float opre = mydot * m.inv_sqrt_W + bo;
      const float o = s_out * opre;
      int64_t row = idx ? (int64_t)idx[(int64_t)net * idx_stride + b] : (int64_t)b;
      const float yv = y_all[row];
      float logp, rr;
      if (m.likelihood == BNF_NORMAL) {
        const float sg = dv[kDvSigma];
        const float d = yv / sg - o / sg;
        logp = -0.5f * d * d - (0.9189385332046727f + logf(sg));
        rr = d / sg;
        a_g0 += (d * d - 1.f) / sg;
      } else {
        const float mean = softplus_f(o);
        const float shp = dv[kDvShape];
        const float rc = 1.f / shp;
        const float lg = -logf(shp) - logf(mean);
        const float sig_l = sigmoid_f(lg);
        float nb = rc * log_sigmoid_f(-lg) + yv * log_sigmoid_f(lg)
                   - (lgammaf(1.f + yv) + lgammaf(rc) - lgammaf(1.f + yv + rc)) - logf(rc + yv);
        float dnb_dl = yv * (1.f - sig_l) - rc * sig_l;
        float dnb_dr = log_sigmoid_f(-lg) - digamma_f(rc) + digamma_f(1.f + yv + rc) - 1.f / (rc + yv);
        float wnb = 1.f;
        logp = nb;
        if (m.likelihood == BNF_ZINB) {
          const float pi = dv[kDvPi];
          if (yv == 0.f) {
            const float A = (1.f - pi) * expf(nb), tot = A + pi;
            logp = logf(tot);
            wnb = A / tot;
            a_g2 += (1.f - expf(nb)) / tot;
          } else {
            logp = log1pf(-pi) + nb;
            a_g2 += -1.f / (1.f - pi);
          }
        }
        rr = wnb * dnb_dl * (-sigmoid_f(o) / mean);
        a_g1 += wnb * (dnb_dl * (-1.f / shp) + dnb_dr * (-1.f / (shp * shp)));
      }
      a_ll += logp;
      a_gs += rr * opre;
      a_gb += rr * s_out;
      rs[b - b0] = rr;
    }
  }
  __syncthreads();
  // ---- (B) activation backward of layer L-1 for these rows (thread keeps one column group)
  const int layer = m.L - 1;
  const float w = dv[kDvActW], s_l = dv[kDvSLayer + layer];
  float g_w = 0.f, g_s = 0.f;
  {
    const int G = m.W / VEC;
    const int cg = threadIdx.x % G, rstep = 256 / G;
    float g_b[VEC], g_ko[VEC], head_c[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      g_b[k] = 0.f; g_ko[k] = 0.f;
      head_c[k] = s_out * m.inv_sqrt_W * Ko[cg * VEC + k];
    }
    // 4 independent row iterations in flight: the loop is latency-bound otherwise
#pragma unroll 4
    for (int b = b0 + threadIdx.x / G; b < b1; b += rstep) {
      const size_t o = ((size_t)net * B + b) * m.W + (size_t)cg * VEC;
      const size_t o3 = ((size_t)net * B + b) * NP * m.W + (size_t)cg * VEC;    // X3: plane 0 of the row
      float zr[VEC], hr[VEC], du_r[VEC];
      if constexpr (X3) {
#pragma unroll
        for (int k = 0; k < VEC; k += 4) *reinterpret_cast<float4*>(zr + k) = *reinterpret_cast<const float4*>(zf + o + k);
      } else {
        alignas(16) T zv[VEC];
        *reinterpret_cast<uint4*>(zv) = *reinterpret_cast<const uint4*>(z + o);
#pragma unroll
        for (int k = 0; k < VEC; ++k) zr[k] = to_f<T>(zv[k]);
      }
#pragma unroll
      for (int pl = 0; pl < NP; ++pl) {
        alignas(16) T hv[VEC];
        *reinterpret_cast<uint4*>(hv) = *reinterpret_cast<const uint4*>(h + o3 + pl * m.W);
#pragma unroll
        for (int k = 0; k < VEC; ++k) hr[k] = pl == 0 ? to_f<T>(hv[k]) : hr[k] + to_f<T>(hv[k]);
      }
      const float rb = rs[b - b0];
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const float zz = zr[k];
        const float dh = rb * head_c[k];
        g_ko[k] += rb * hr[k];
        float diff;
        float da;
        if constexpr (X3) {
          // the activation math of the bf16x3 tensor-core epilogues (one ex2 + one rcp, <= 3e-7 abs.)
          float hh;
          da = act_grad_x3(zz, w, &diff, &hh);
        } else {
          da = act_grad_sel<FAST>(zz, w, &diff);
        }
        const float dz = dh * da;
        g_w += dh * diff;
        g_s += dz * zz;
        const float du = dz * s_l;
        g_b[k] += du;
        du_r[k] = du;
      }
      if constexpr (X3) {
        // dU carries two planes (bnf_tc.cu: kDuPlanes), row layout [plane 0: W | plane 1: W]
        alignas(16) uint32_t pk[2][VEC / 2];
#pragma unroll
        for (int k = 0; k < VEC; k += 2) split2_pair(du_r[k], du_r[k + 1], &pk[0][k / 2], &pk[1][k / 2]);
        const size_t o2 = ((size_t)net * B + b) * 2 * m.W + (size_t)cg * VEC;
#pragma unroll
        for (int pl = 0; pl < 2; ++pl)
          *reinterpret_cast<uint4*>(dU + o2 + pl * m.W) = *reinterpret_cast<const uint4*>(pk[pl]);
      } else {
        alignas(16) T out[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) out[k] = from_f<T>(du_r[k]);
        *reinterpret_cast<uint4*>(dU + o) = *reinterpret_cast<const uint4*>(out);
      }
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      atomicAdd(&colsum[cg * VEC + k], g_b[k]);
      atomicAdd(&colsum[m.W + cg * VEC + k], g_ko[k]);
    }
  }
  // ---- block reductions -> global atomics
  a_ll = warp_sum(a_ll); a_g0 = warp_sum(a_g0); a_g1 = warp_sum(a_g1); a_g2 = warp_sum(a_g2);
  a_gs = warp_sum(a_gs); a_gb = warp_sum(a_gb); g_w = warp_sum(g_w); g_s = warp_sum(g_s);
  if (lane == 0) {
    hred[0][warp] = a_ll; hred[1][warp] = a_g0; hred[2][warp] = a_g1; hred[3][warp] = a_g2;
    hred[4][warp] = a_gs; hred[5][warp] = a_gb; hred[6][warp] = g_w; hred[7][warp] = g_s;
  }
  __syncthreads();
  float* g = grad + (size_t)net * m.P;
  for (int i = threadIdx.x; i < m.W; i += 256) {
    atomicAdd(&g[m.off_bias[layer] + i], colsum[i]);
    atomicAdd(&g[m.off_kernel[m.L] + i], colsum[m.W + i] * s_out * m.inv_sqrt_W);
  }
  if (threadIdx.x == 0) {
    float t[8];
    for (int k = 0; k < 8; ++k) { t[k] = 0.f; for (int i = 0; i < 8; ++i) t[k] += hred[k][i]; }
    atomicAdd(&ll[net], t[0]);
    if (m.likelihood == BNF_NORMAL) {
      atomicAdd(&g[0], t[1] * expf(p[0]));
    } else {
      atomicAdd(&g[1], t[2] * sigmoid_f(p[1]));
      if (m.likelihood == BNF_ZINB) { const float pi = dv[kDvPi]; atomicAdd(&g[2], t[3] * pi * (1.f - pi)); }
    }
    atomicAdd(&g[m.off_out_scale], t[4] * sigmoid_f(p[m.off_out_scale]));
    atomicAdd(&g[m.off_bias[m.L]], t[5]);
    atomicAdd(&g[m.off_actw], t[6] * w * (1.f - w));
    atomicAdd(&g[m.off_layer_scale[layer]], (t[7] / s_l) * sigmoid_f(p[m.off_layer_scale[layer]]));
  }
}

// =============================================================================
// head + activation backward of the last hidden layer with ONE WARP PER ROW (bf16 / bf16x3 at
// W = 512 and 1024): a row's h stays in registers between the h.Ko dot and the Dense_L kernel
// gradient r*h, so h, z and dU cross HBM exactly once (head_fused_kernel reads h twice, 8 instead
// of 6 bytes per element: profiles/ncu_wind_tc_gemm_r2_summary.csv).  Lane l owns the columns
// (j*32 + l)*8 + k (j < NC/8, k < 8) of every row, so its 2*NC column sums (bias of layer L-1,
// Dense_L kernel) live in registers for the whole block and are flushed once.  RIF rows are in
// flight per warp; lanes 0..RIF-1 evaluate the likelihood of one row each.
// Same math as head_fused_kernel.
// =============================================================================
template <int NC, bool X3>
__global__ void __launch_bounds__(256)
head_rows_kernel(const __grid_constant__ DevModel m, const float* __restrict__ params,
                 const float* __restrict__ derived, const __nv_bfloat16* __restrict__ h,
                 const void* __restrict__ z_any, const float* __restrict__ y_all,
                 const int32_t* __restrict__ idx, int64_t idx_stride, int B, int R /* rows per block */,
                 __nv_bfloat16* __restrict__ dU, float* __restrict__ ll, float* __restrict__ grad) {
  using T = __nv_bfloat16;
  constexpr int W = NC * 32;
  constexpr int NV = NC / 8;                          // 16-byte vectors per lane and row
  constexpr int RIF = X3 ? 64 / NC : 128 / NC;        // rows in flight per warp (64 registers of h)
  constexpr int NP = X3 ? 3 : 1;
  const T* z = static_cast<const T*>(z_any);
  const float* zf = static_cast<const float*>(z_any);
  extern __shared__ float fsm[];
  float* colsum = fsm;                                // [2W] bias / Dense_L kernel column sums
  float* kos = fsm + 2 * W;                           // [W] Dense_L kernel
  __shared__ float hred[8][8];
  pdl_enter(params, derived, h, z, zf, y_all, idx, dU, ll, grad);
  const int net = blockIdx.y;
  const int b0 = blockIdx.x * R, b1 = min(B, b0 + R);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* p = params + (size_t)net * m.P;
  const float* dv = derived + (size_t)net * kDerivedStride;
  const float* Ko = p + m.off_kernel[m.L];
  const float bo = p[m.off_bias[m.L]], s_out = dv[kDvSOut];
  const int layer = m.L - 1;
  const float w = dv[kDvActW], s_l = dv[kDvSLayer + layer];
  const float c_head = s_out * m.inv_sqrt_W;
  for (int i = threadIdx.x; i < 2 * W; i += 256) colsum[i] = 0.f;
  for (int i = threadIdx.x; i < W; i += 256) kos[i] = Ko[i];
  __syncthreads();
  float g_b[NC], g_ko[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) { g_b[c] = 0.f; g_ko[c] = 0.f; }
  float a_ll = 0.f, a_gs = 0.f, a_gb = 0.f, g_w = 0.f, g_s = 0.f;
  float gl3[3] = {0.f, 0.f, 0.f};
  for (int base = b0 + warp * RIF; base < b1; base += 8 * RIF) {
    // ---- the rows' h into registers (bf16: packed as loaded; bf16x3: the f32 sum of the planes)
    uint4 hv[X3 ? 1 : RIF][X3 ? 1 : NV];
    float hf[X3 ? RIF : 1][X3 ? NC : 1];
#pragma unroll
    for (int u = 0; u < RIF; ++u) {
      const int b = min(base + u, b1 - 1);            // a ragged tail re-reads the last row (ignored below)
      const T* hr = h + ((size_t)net * B + b) * NP * W;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int c0 = (j * 32 + lane) * 8;
        if constexpr (X3) {
#pragma unroll
          for (int pl = 0; pl < 3; ++pl) {
            alignas(16) T t8[8];
            *reinterpret_cast<uint4*>(t8) = *reinterpret_cast<const uint4*>(hr + pl * W + c0);
#pragma unroll
            for (int k = 0; k < 8; ++k) hf[u][j * 8 + k] = pl == 0 ? to_f<T>(t8[k]) : hf[u][j * 8 + k] + to_f<T>(t8[k]);
          }
        } else {
          hv[u][j] = *reinterpret_cast<const uint4*>(hr + c0);
        }
      }
    }
    // ---- h.Ko of every row in flight
    float dot[RIF];
#pragma unroll
    for (int u = 0; u < RIF; ++u) {
      float d = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int c0 = (j * 32 + lane) * 8;
        alignas(16) float kk[8];
        *reinterpret_cast<float4*>(kk) = *reinterpret_cast<const float4*>(kos + c0);
        *reinterpret_cast<float4*>(kk + 4) = *reinterpret_cast<const float4*>(kos + c0 + 4);
        if constexpr (X3) {
#pragma unroll
          for (int k = 0; k < 8; ++k) d = fmaf(hf[u][j * 8 + k], kk[k], d);
        } else {
          const T* t8 = reinterpret_cast<const T*>(&hv[u][j]);
#pragma unroll
          for (int k = 0; k < 8; ++k) d = fmaf(to_f<T>(t8[k]), kk[k], d);
        }
      }
      dot[u] = warp_sum(d);
    }
    // ---- likelihood: lane u owns row base + u
    float mydot = 0.f;
#pragma unroll
    for (int u = 0; u < RIF; ++u) if (lane == u) mydot = dot[u];
    float rr = 0.f;
    if (lane < RIF && base + lane < b1) {
      const int b = base + lane;
      const float opre = mydot * m.inv_sqrt_W + bo;
      const int64_t row = idx ? (int64_t)idx[(int64_t)net * idx_stride + b] : (int64_t)b;
      a_ll += head_row_loglik(m.likelihood, dv, s_out * opre, y_all[row], &rr, gl3);
      a_gs += rr * opre;
      a_gb += rr * s_out;
    }
    // ---- activation backward of layer L-1, row by row, h from registers
#pragma unroll
    for (int u = 0; u < RIF; ++u) {
      const int b = base + u;
      const float rb = __shfl_sync(0xffffffffu, rr, u);
      if (b < b1) {                                   // warp-uniform
        const float rbs = rb * c_head;
        const size_t orow = (size_t)net * B + b;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          const int c0 = (j * 32 + lane) * 8;
          float zr[8], du_r[8];
          alignas(16) float kk[8];
          if constexpr (X3) {
            *reinterpret_cast<float4*>(zr) = *reinterpret_cast<const float4*>(zf + orow * W + c0);
            *reinterpret_cast<float4*>(zr + 4) = *reinterpret_cast<const float4*>(zf + orow * W + c0 + 4);
          } else {
            alignas(16) T zv[8];
            *reinterpret_cast<uint4*>(zv) = *reinterpret_cast<const uint4*>(z + orow * W + c0);
#pragma unroll
            for (int k = 0; k < 8; ++k) zr[k] = to_f<T>(zv[k]);
          }
          *reinterpret_cast<float4*>(kk) = *reinterpret_cast<const float4*>(kos + c0);
          *reinterpret_cast<float4*>(kk + 4) = *reinterpret_cast<const float4*>(kos + c0 + 4);
          const T* t8 = reinterpret_cast<const T*>(&hv[X3 ? 0 : u][X3 ? 0 : j]);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float hh = X3 ? hf[X3 ? u : 0][X3 ? j * 8 + k : 0] : to_f<T>(t8[k]);
            const float zz = zr[k];
            const float dh = rbs * kk[k];
            g_ko[j * 8 + k] = fmaf(rb, hh, g_ko[j * 8 + k]);
            float diff, da;
            if constexpr (X3) {
              float unused;
              da = act_grad_x3(zz, w, &diff, &unused);
            } else {
              da = act_grad_sel<true>(zz, w, &diff);
            }
            const float dz = dh * da;
            g_w = fmaf(dh, diff, g_w);
            g_s = fmaf(dz, zz, g_s);
            const float du = dz * s_l;
            g_b[j * 8 + k] += du;
            du_r[k] = du;
          }
          if constexpr (X3) {
            alignas(16) uint32_t pk[2][4];
#pragma unroll
            for (int k = 0; k < 8; k += 2) split2_pair(du_r[k], du_r[k + 1], &pk[0][k / 2], &pk[1][k / 2]);
#pragma unroll
            for (int pl = 0; pl < 2; ++pl)
              *reinterpret_cast<uint4*>(dU + orow * 2 * W + pl * W + c0) = *reinterpret_cast<const uint4*>(pk[pl]);
          } else {
            alignas(16) T out[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) out[k] = from_f<T>(du_r[k]);
            *reinterpret_cast<uint4*>(dU + orow * W + c0) = *reinterpret_cast<const uint4*>(out);
          }
        }
      }
    }
  }
  // ---- column sums: registers -> shared (8 warps) -> one global atomic per column and block
#pragma unroll
  for (int j = 0; j < NV; ++j) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = (j * 32 + lane) * 8 + k;
      atomicAdd(&colsum[c], g_b[j * 8 + k]);
      atomicAdd(&colsum[W + c], g_ko[j * 8 + k]);
    }
  }
  a_ll = warp_sum(a_ll); gl3[0] = warp_sum(gl3[0]); gl3[1] = warp_sum(gl3[1]); gl3[2] = warp_sum(gl3[2]);
  a_gs = warp_sum(a_gs); a_gb = warp_sum(a_gb); g_w = warp_sum(g_w); g_s = warp_sum(g_s);
  if (lane == 0) {
    hred[0][warp] = a_ll; hred[1][warp] = gl3[0]; hred[2][warp] = gl3[1]; hred[3][warp] = gl3[2];
    hred[4][warp] = a_gs; hred[5][warp] = a_gb; hred[6][warp] = g_w; hred[7][warp] = g_s;
  }
  __syncthreads();
  float* g = grad + (size_t)net * m.P;
  for (int i = threadIdx.x; i < W; i += 256) {
    atomicAdd(&g[m.off_bias[layer] + i], colsum[i]);
    atomicAdd(&g[m.off_kernel[m.L] + i], colsum[W + i] * c_head);
  }
  if (threadIdx.x == 0) {
    float t[8];
    for (int k = 0; k < 8; ++k) { t[k] = 0.f; for (int i = 0; i < 8; ++i) t[k] += hred[k][i]; }
    atomicAdd(&ll[net], t[0]);
    if (m.likelihood == BNF_NORMAL) {
      atomicAdd(&g[0], t[1] * expf(p[0]));
    } else {
      atomicAdd(&g[1], t[2] * sigmoid_f(p[1]));
      if (m.likelihood == BNF_ZINB) { const float pi = dv[kDvPi]; atomicAdd(&g[2], t[3] * pi * (1.f - pi)); }
    }
    atomicAdd(&g[m.off_out_scale], t[4] * sigmoid_f(p[m.off_out_scale]));
    atomicAdd(&g[m.off_bias[m.L]], t[5]);
    atomicAdd(&g[m.off_actw], t[6] * w * (1.f - w));
    atomicAdd(&g[m.off_layer_scale[layer]], (t[7] / s_l) * sigmoid_f(p[m.off_layer_scale[layer]]));
  }
}

// =============================================================================
// minibatch row windows drawn on the device
//   MAP/MLE (inference.py:583-597): every member draws a fresh permutation of the n_total rows per
//   epoch and walks it in windows of B rows (the ragged tail is dropped); VI (:704-709): one
//   shared window per step = the first B entries of a fresh permutation.
// The kernel reads the number of COMPLETED steps g from device memory (so its launch arguments
// never change and the whole step replays as a CUDA graph): epoch = g / steps_per_epoch, window
// = g % steps_per_epoch, and evaluates the keyed permutation (perm_index) for the B slots of each
// index row: idx_win[row*B + i] = perm_{member(row), epoch}(window*B + i).
// =============================================================================
__global__ void __launch_bounds__(256)
batch_window_kernel(int n_total, int B, int steps_per_epoch, uint64_t seed, int64_t first_member,
                    const int32_t* step_count, int32_t* __restrict__ idx_win) {
  __shared__ PermKeys pk;
  pdl_enter(step_count, idx_win);
  const int g = __ldcg(step_count);
  const uint32_t epoch = (uint32_t)(g / steps_per_epoch), window = (uint32_t)(g % steps_per_epoch);
  if (threadIdx.x == 0) pk = perm_keys(seed, (uint64_t)(first_member + blockIdx.y), epoch);
  __syncthreads();
  const uint32_t hb = perm_half_bits((uint32_t)n_total);
  int32_t* out = idx_win + (size_t)blockIdx.y * B;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x)
    out[i] = (int32_t)perm_index(window * (uint32_t)B + (uint32_t)i, (uint32_t)n_total, hb, pk);
}

// =============================================================================
// prior + Adam (models.py:94-103; inference.py:558-569,580,605-606)
// g_loss = -(c_ll*g_ll + pw*dlogprior); optax.adam; also sum of log-prior.
// =============================================================================
// prologue of a multi-step call: loss-row cursor = 0, address of the caller's loss buffer -> workspace
__global__ void arm_loss_kernel(float** loss_slot, float* out_loss, int32_t* cursor) {
  *loss_slot = out_loss;
  *cursor = 0;
}
void launch_arm_loss(float** loss_slot, float* out_loss, int32_t* cursor, cudaStream_t st) {
  BNF_PROF("arm_loss", st);
  arm_loss_kernel<<<1, 1, 0, st>>>(loss_slot, out_loss, cursor);
}

__global__ void tick_kernel(int32_t* step_count, int32_t* slot) {
  *step_count += 1;
  if (slot) *slot += 1;
}

__global__ void __launch_bounds__(256)
map_adam_kernel(int P, float* __restrict__ params, float* __restrict__ am, float* __restrict__ av,
                const float* __restrict__ g_ll, const int32_t* __restrict__ step_count, float c_ll,
                float prior_weight, float lr, float* __restrict__ prior_out) {
  const int net = blockIdx.y;
  const int t = *step_count;  // already incremented for this step
  const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
  const float bc1 = 1.f - powf(b1, (float)t), bc2 = 1.f - powf(b2, (float)t);
  float lp = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
    const size_t o = (size_t)net * P + i;
    const float th = params[o];
    float g = -(c_ll * g_ll[o]);
    if (prior_weight != 0.f) {
      const float zz = th - (i == 1 ? -1.5f : 0.f);
      lp += -zz - 2.f * softplus_f(-zz);
      g = -(c_ll * g_ll[o] + prior_weight * (-tanhf(0.5f * zz)));
    }
    const float mm = (1.f - b1) * g + b1 * am[o];
    const float vv = (1.f - b2) * (g * g) + b2 * av[o];
    am[o] = mm;
    av[o] = vv;
    params[o] = th + (-lr) * ((mm / bc1) / (sqrtf(vv / bc2) + eps));
  }
  if (prior_weight != 0.f) {
    __shared__ float pred[8];
    lp = warp_sum(lp);
    if ((threadIdx.x & 31) == 0) pred[threadIdx.x >> 5] = lp;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
      for (int i = 0; i < 8; ++i) tot += pred[i];
      atomicAdd(&prior_out[net], tot);       // one atomic per block (8 addresses in total)
    }
  }
}

// out row = (*slot - 1): the device-side step cursor keeps the launch arguments
// identical for every step, so a whole step can be replayed as one CUDA graph.
__global__ void map_loss_kernel(int n_net, const float* ll, const float* prior, float c_ll,
                                float prior_weight, float* out, const int32_t* slot) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot) out += (size_t)(*slot - 1) * n_net;
  if (j < n_net) {
    out[j] = prior_weight == 0.f ? -(ll[j] * c_ll) : -(ll[j] * c_ll + prior[j] * prior_weight);
  }
}

// -----------------------------------------------------------------------------
// Fused MAP update = the whole tail of a training step in one launch:
//   prior gradient + optax.adam (as map_adam_kernel), gradient buffer zeroed for the next
//   step, bf16 restaging of the hidden kernels (natural (in,out) layout = the same element
//   order as the f32 master, so the store is coalesced), and -- by the LAST block to finish
//   (ticket counter) -- the step's loss row, the reset of the loglik / prior accumulators, the
//   Adam step count / loss-row cursor ticks and the derived scalars of the NEXT step.
// A training step is then encode -> GEMMs -> head -> GEMMs -> encode_bwd -> this kernel, with
// no memset / prep / cast / loss nodes in between (inference.py:599-608 per step).
// step_count holds the number of COMPLETED steps on entry; slot the loss row to write.
// -----------------------------------------------------------------------------
// FAST (tensor-core mode): MUFU exp/log/rcp/sqrt approximations (~1e-6 relative on the update of
// f32 master weights whose bf16 copies feed the GEMMs); the fp32 parity mode keeps optax's exact
// division / sqrt sequence.
template <bool FAST>
__global__ void __launch_bounds__(256)
map_update_kernel(const __grid_constant__ DevModel m, float* params, float* __restrict__ am,
                  float* __restrict__ av, float* __restrict__ grad, int32_t* step_count, float c_ll,
                  float prior_weight, float lr, float* prior, float* ll, float* const* loss_slot,
                  int32_t* slot, unsigned int* counter, float* __restrict__ derived,
                  __nv_bfloat16* __restrict__ wn, size_t w_per_net, int wn_planes, int n_net) {
  __shared__ float pred[8];
  __shared__ int s_last;
  pdl_enter(params, am, av, grad, step_count, prior, ll, loss_slot, slot, counter, derived, wn);
  const int net = blockIdx.y, P = m.P;
  const int t = __ldcg(step_count) + 1;
  const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
  // the two powf() bias corrections cost more than a whole element update: once per block
  __shared__ float s_bc[2];
  if (threadIdx.x == 0) { s_bc[0] = 1.f - powf(b1, (float)t); s_bc[1] = 1.f - powf(b2, (float)t); }
  __syncthreads();
  const float bc1 = s_bc[0], bc2 = s_bc[1];
  const float rbc1 = 1.f / bc1, rbc2 = 1.f / bc2;
  float lp = 0.f;
  // four independent elements per thread and iteration: 16 loads in flight before the first use
  constexpr int U = 4;
  const int stride = gridDim.x * blockDim.x;
  for (int i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < P; i0 += U * stride) {
    float th[U], gl[U], m0[U], v0[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * stride;
      if (i < P) {
        const size_t o = (size_t)net * P + i;
        th[u] = params[o]; gl[u] = grad[o]; m0[u] = am[o]; v0[u] = av[o];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * stride;
      if (i >= P) break;
      const size_t o = (size_t)net * P + i;
      float g = -(c_ll * gl[u]);
      if (prior_weight != 0.f) {
        const float zz = th[u] - (i == 1 ? -1.5f : 0.f);
        if (FAST) {
          // one exponential serves both: softplus(-z) = max(-z,0) + log(1+e), tanh(z/2) = sgn(z)(1-e)/(1+e), e = exp(-|z|)
          const float e = ex2_fast(-fabsf(zz) * 1.4426950408889634f);
          const float r1 = __fdividef(1.f, 1.f + e);
          lp += -zz - 2.f * (fmaxf(-zz, 0.f) + __logf(1.f + e));
          g = -(c_ll * gl[u] - prior_weight * copysignf((1.f - e) * r1, zz));
        } else {
          lp += -zz - 2.f * softplus_f(-zz);
          g = -(c_ll * gl[u] + prior_weight * (-tanhf(0.5f * zz)));
        }
      }
      const float mm = (1.f - b1) * g + b1 * m0[u];
      const float vv = (1.f - b2) * (g * g) + b2 * v0[u];
      am[o] = mm;
      av[o] = vv;
      float th_new;
      if (FAST) {
        float sq;
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq) : "f"(vv * rbc2));
        th_new = th[u] - lr * __fdividef(mm * rbc1, sq + eps);
      } else {
        th_new = th[u] + (-lr) * ((mm / bc1) / (sqrtf(vv / bc2) + eps));
      }
      params[o] = th_new;
      // The zero for the next step's accumulation is stored AFTER the values that depend on the
      // loaded gradient: a store issued right behind the load of the same address stalls the
      // SM's in-order LSU until the line arrives (measured: 4x slower kernel).
      grad[o] = 0.f;
      if (wn) {
        // hidden-layer kernel leaf?  wn = [layer][Kp][W] per network, rows >= fan_in stay zero
        for (int l = 0; l < m.L; ++l) {
          const int rel = i - m.off_kernel[l];
          const int cnt = (l == 0 ? m.F : m.W) * m.W;
          if (rel >= 0 && rel < cnt) {
            const size_t lo = l == 0 ? 0 : (size_t)m.Fp * m.W + (size_t)(l - 1) * m.W * m.W;
            if (wn_planes == 3) {
              // bf16x3 staging: [Kp][3*W], plane p of (k, n) at k*3W + p*W + n
              const int k = rel / m.W, n = rel - k * m.W;
              __nv_bfloat16* d = wn + 3 * ((size_t)net * w_per_net + lo) + (size_t)k * 3 * m.W + n;
              split3_one(th_new, d, d + m.W, d + 2 * m.W);
            } else {
              wn[(size_t)net * w_per_net + lo + rel] = __float2bfloat16_rn(th_new);
            }
            break;
          }
        }
      }
    }
  }
  if (prior_weight != 0.f) {
    lp = warp_sum(lp);
    if ((threadIdx.x & 31) == 0) pred[threadIdx.x >> 5] = lp;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
      for (int i = 0; i < 8; ++i) tot += pred[i];
      atomicAdd(&prior[net], tot);
    }
  }
  // ---- last block of the grid: loss row, accumulator reset, ticks, next step's derived ----
  // (bar.sync, then ONE gpu-scope fence by the signalling thread: fences are cumulative over
  // the block barrier, the same pattern as a cooperative-groups grid sync)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int ticket = atomicAdd(counter, 1u);
    s_last = ticket == gridDim.x * gridDim.y - 1u;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int row = __ldcg(slot);
  float* out_loss = reinterpret_cast<float*>(__ldcg(reinterpret_cast<const unsigned long long*>(loss_slot)));
  for (int j = threadIdx.x; j < n_net; j += blockDim.x) {
    const float l = __ldcg(ll + j), pr = __ldcg(prior + j);
    out_loss[(size_t)row * n_net + j] = prior_weight == 0.f ? -(l * c_ll) : -(l * c_ll + pr * prior_weight);
    ll[j] = 0.f;
    prior[j] = 0.f;
  }
  if (threadIdx.x == 0) {
    *slot = row + 1;
    *step_count = t;
    *counter = 0u;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = warp; j < n_net; j += 8)
    prep_one(m, params + (size_t)j * P, derived + (size_t)j * kDerivedStride, lane);
}

// =============================================================================
// VI (inference.py:687-739)
// =============================================================================
// z[s,e,p] = mu[e,p] + sigma[e,p]*eps[s,e,p] ; eps drawn here when eps_in == NULL: a thread owns
// four consecutive elements = the four normals of ONE Philox block (counter = element index / 4),
// so no two elements share a random word (mean-field draws must be independent).
__global__ void __launch_bounds__(256)
vi_sample_kernel(int P, int E, int S, const float* __restrict__ mu, const float* __restrict__ rho,
                 const float* __restrict__ eps_in, float* __restrict__ eps_out, uint64_t seed,
                 uint64_t stream_id, const int32_t* __restrict__ step_ptr, float* __restrict__ z) {
  // step_ptr (optional): the device-side step counter -- every optimisation step draws from its
  // own Philox stream without the host passing a new seed (the step sequence replays as a graph)
  if (step_ptr) stream_id ^= (uint64_t)(uint32_t)__ldcg(step_ptr) << 20;
  const size_t total = (size_t)S * E * P, EP = (size_t)E * P;
  const size_t groups = (total + 3) / 4;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (size_t)gridDim.x * blockDim.x) {
    float e4[4] = {0.f, 0.f, 0.f, 0.f};
    if (!eps_in) {
      const float4 n4 = philox_normal4(seed, stream_id, g);
      e4[0] = n4.x; e4[1] = n4.y; e4[2] = n4.z; e4[3] = n4.w;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const size_t i = 4 * g + k;
      if (i >= total) break;
      const size_t ep = i % EP;
      const float e = eps_in ? eps_in[i] : e4[k];
      if (eps_out) eps_out[i] = e;
      const float sg = 1e-4f + softplus_f(rho[ep]);
      z[i] = mu[ep] + sg * e;
    }
  }
}

// Gradient assembly + Adam on (mu, rho) + loss terms.
//   T(z) = logprior(z) + c*loglik(z), c = (N/B)/kl ;  gT = dlogprior(z) + c*g_ll
//   dL/dmu = -mean_s gT ; dL/drho = sigmoid(rho)*(-1/sigma - mean_s gT*eps)
//   loss_e = mean_s [ logq(z_s) - logprior(z_s) ] - c*mean_s loglik_s   (second part added later)
__global__ void __launch_bounds__(256)
vi_adam_kernel(int P, int E, int S, float* __restrict__ mu, float* __restrict__ rho,
               float* __restrict__ am, float* __restrict__ av, const float* __restrict__ z,
               const float* __restrict__ eps, const float* __restrict__ g_ll,
               const int32_t* __restrict__ step_count, float c, float lr, float* __restrict__ loss_acc) {
  const int e = blockIdx.y;
  const int t = *step_count;
  const float b1 = 0.9f, b2 = 0.999f, aeps = 1e-8f;
  // the two powf() bias corrections cost more than an element update: once per block
  __shared__ float s_bc[2];
  if (threadIdx.x == 0) { s_bc[0] = 1.f - powf(b1, (float)t); s_bc[1] = 1.f - powf(b2, (float)t); }
  __syncthreads();
  const float bc1 = s_bc[0], bc2 = s_bc[1];
  const float invS = 1.f / (float)S;
  float lacc = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
    const size_t o = (size_t)e * P + i;
    const float rh = rho[o], m0 = mu[o];
    const float sg = 1e-4f + softplus_f(rh);
    const float loc = (i == 1 ? -1.5f : 0.f);
    float s_gt = 0.f, s_gte = 0.f;
    for (int s = 0; s < S; ++s) {
      const size_t os = ((size_t)s * E + e) * P + i;
      const float zz = z[os] - loc, ee = eps[os];
      const float gt = -tanhf(0.5f * zz) + c * g_ll[os];
      s_gt += gt;
      s_gte += gt * ee;
      // log q - log prior for this coordinate
      lacc += (-0.5f * ee * ee - 0.9189385332046727f - logf(sg)) - (-zz - 2.f * softplus_f(-zz));
    }
    const float gmu = -s_gt * invS;
    const float grho = sigmoid_f(rh) * (-1.f / sg - s_gte * invS);
    // Adam state layout: [E, 2, P]  (mu block then rho block per member)
    const size_t om = ((size_t)e * 2 + 0) * P + i, orr = ((size_t)e * 2 + 1) * P + i;
    float mm = (1.f - b1) * gmu + b1 * am[om], vv = (1.f - b2) * (gmu * gmu) + b2 * av[om];
    am[om] = mm; av[om] = vv;
    mu[o] = m0 + (-lr) * ((mm / bc1) / (sqrtf(vv / bc2) + aeps));
    mm = (1.f - b1) * grho + b1 * am[orr]; vv = (1.f - b2) * (grho * grho) + b2 * av[orr];
    am[orr] = mm; av[orr] = vv;
    rho[o] = rh + (-lr) * ((mm / bc1) / (sqrtf(vv / bc2) + aeps));
  }
  lacc = warp_sum(lacc);
  if ((threadIdx.x & 31) == 0) atomicAdd(&loss_acc[e], lacc * invS);
}

// out row = (*slot - 1) of the buffer whose address sits in *out_slot when those are given (multi-
// step calls: device cursors keep the launch arguments constant), else `out` directly.
__global__ void vi_loss_kernel(int E, int S, const float* loss_acc, const float* ll, float c, float* out,
                               float* const* out_slot, const int32_t* slot) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (out_slot) out = *out_slot + (size_t)(*slot - 1) * E;
  if (e < E) {
    float sll = 0.f;
    for (int s = 0; s < S; ++s) sll += ll[(size_t)s * E + e];
    out[e] = loss_acc[e] - c * sll / (float)S;
  }
}

// =============================================================================
// init (inference.py:399-427, :203-231): TruncatedNormal(0,1,[-2,2]) kernels
// =============================================================================
__global__ void __launch_bounds__(256)
init_params_kernel(const __grid_constant__ DevModel m, float lns_init, uint64_t seed,
                   int64_t first_member, float* __restrict__ params) {
  const int net = blockIdx.y;
  float* p = params + (size_t)net * m.P;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m.P; i += gridDim.x * blockDim.x) {
    bool is_kernel = false;
    for (int l = 0; l <= m.L; ++l) {
      int fan = l == 0 ? m.F : m.W, out = l == m.L ? 1 : m.W;
      if (i >= m.off_kernel[l] && i < m.off_kernel[l] + fan * out) is_kernel = true;
    }
    float v = 0.f;
    if (i == 0) v = lns_init;
    else if (is_kernel) {
      // rejection sampling on the Philox blocks of this (member, parameter): stream = global
      // member id, block = (parameter, attempt group)
      bool found = false;
      for (int blk = 0; blk < 4 && !found; ++blk) {
        const float4 n4 = philox_normal4(seed, (uint64_t)(first_member + net), ((uint64_t)i << 2) | (uint64_t)blk);
        const float c[4] = {n4.x, n4.y, n4.z, n4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (!found && fabsf(c[k]) <= 2.f) { v = c[k]; found = true; }
      }
      // all 16 candidates outside [-2,2] has probability 0.0455^16: v stays 0
    }
    p[i] = v;
  }
}

// =============================================================================
// mixture quantiles (inference.py:42-84)
// =============================================================================
__global__ void minmax_kernel(const float* __restrict__ v, size_t n, float* __restrict__ out /*[2]: min,max as ordered ints*/) {
  float lo = FLT_MAX, hi = -FLT_MAX;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float x = v[i];
    lo = fminf(lo, x);
    hi = fmaxf(hi, x);
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    // float atomic min/max through the order-preserving int mapping
    auto enc = [](float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; };
    atomicMin((int*)&out[0], enc(lo));
    atomicMax((int*)&out[1], enc(hi));
  }
}
__global__ void minmax_init_kernel(float* mm) {
  ((int*)mm)[0] = 0x7fffffff; ((int*)mm)[1] = (int)0x80000000;
  ((int*)mm)[2] = 0x7fffffff; ((int*)mm)[3] = (int)0x80000000;
}
__device__ __forceinline__ float dec_ordered(float enc) {
  int i = __float_as_int(enc);
  return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff);
}

__device__ __forceinline__ float mix_cdf(const float* __restrict__ means, const float* __restrict__ scales,
                                         int M, int N, int n, float xq) {
  float acc = 0.f;
  for (int c = 0; c < M; ++c) {
    float zz = (xq - means[(size_t)c * N + n]) / scales[c];
    acc += 0.5f * erfcf(-zz * 0.7071067811865476f);
  }
  return acc / (float)M;
}

__global__ void __launch_bounds__(128)
quantile_root_kernel(const float* __restrict__ means, const float* __restrict__ scales, int M, int N,
                     const float* __restrict__ mm /* enc min/max of means [0,1], scales [2,3] */,
                     float q, float* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float smax = dec_ordered(mm[3]);
  float b = dec_ordered(mm[0]) - 5.f * smax;   // low
  float a = dec_ordered(mm[1]) + 5.f * smax;   // high
  float fa = mix_cdf(means, scales, M, N, n, a) - q;
  float fb = mix_cdf(means, scales, M, N, n, b) - q;
  float c = a, fc = fa, t = 0.5f;
  float best = fabsf(fa) < fabsf(fb) ? a : b, fbest = fminf(fabsf(fa), fabsf(fb));
  for (int it = 0; it < 60 && fbest > 1e-5f; ++it) {
    const float xt = a + t * (b - a);
    const float ft = mix_cdf(means, scales, M, N, n, xt) - q;
    const bool same = (ft > 0.f) == (fa > 0.f) && (ft < 0.f) == (fa < 0.f);
    if (same) { c = a; fc = fa; } else { c = b; fc = fb; b = a; fb = fa; }
    a = xt; fa = ft;
    if (fabsf(ft) < fbest) { fbest = fabsf(ft); best = xt; }
    const float xi = (a - b) / (c - b), phi = (fa - fb) / (fc - fb);
    if (phi * phi < xi && (1.f - phi) * (1.f - phi) < 1.f - xi) {
      t = fa / (fb - fa) * fc / (fb - fc) + (c - a) / (b - a) * fa / (fc - fa) * fb / (fc - fb);
    } else {
      t = 0.5f;
    }
    const float tl = 1e-8f / fmaxf(fabsf(b - a), 1e-30f);
    t = fminf(fmaxf(t, tl), 1.f - tl);
    if (!(t == t) || isinf(t)) t = 0.5f;
    if (fabsf(b - a) <= 2.f * FLT_EPSILON * fabsf(a) + 1e-30f) break;
  }
  out[n] = best;
}

__global__ void __launch_bounds__(128)
quantile_approx_kernel(const float* __restrict__ means, const float* __restrict__ scales, int M, int N,
                       float ndtri_q, float* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s1 = 0.f, s2 = 0.f;
  for (int c = 0; c < M; ++c) {
    float mu = means[(size_t)c * N + n], sg = scales[c];
    s1 += mu;
    s2 += sg * sg + mu * mu;
  }
  const float mean = s1 / (float)M;
  const float sd = sqrtf(s2 / (float)M - mean * mean);
  out[n] = mean + sd * ndtri_q;
}

// =============================================================================
// NB / ZINB predictive mean and mixture quantiles (inference.py:271-333)
// CDF of NegativeBinomial(total_count=r, logits=l) at integer k: I_{sigmoid(-l)}(r, 1+k)
// (TFP NegativeBinomial._cdf), evaluated in double (continued fraction, modified Lentz).
// =============================================================================
__device__ double betacf_d(double a, double b, double x) {
  const double FPMIN = 1e-300, EPS = 1e-12;
  const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
  double c = 1.0, d = 1.0 - qab * x / qap;
  if (fabs(d) < FPMIN) d = FPMIN;
  d = 1.0 / d;
  double h = d;
  for (int m = 1; m <= 2000; ++m) {
    const double m2 = 2.0 * m;
    double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
    d = 1.0 + aa * d; if (fabs(d) < FPMIN) d = FPMIN;
    c = 1.0 + aa / c; if (fabs(c) < FPMIN) c = FPMIN;
    d = 1.0 / d;
    h *= d * c;
    aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
    d = 1.0 + aa * d; if (fabs(d) < FPMIN) d = FPMIN;
    c = 1.0 + aa / c; if (fabs(c) < FPMIN) c = FPMIN;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (fabs(del - 1.0) < EPS) break;
  }
  return h;
}
// I_x(a, b) with x given as (log x, log(1-x)) to stay accurate when x -> 0 or 1
__device__ double betainc_d(double a, double b, double logx, double log1mx) {
  const double x = exp(logx), y = exp(log1mx);
  if (x <= 0.0) return 0.0;
  if (y <= 0.0) return 1.0;
  const double bt = exp(lgamma(a + b) - lgamma(a) - lgamma(b) + a * logx + b * log1mx);
  if (x < (a + 1.0) / (a + b + 2.0)) return bt * betacf_d(a, b, x) / a;
  return 1.0 - bt * betacf_d(b, a, y) / b;
}
__device__ __forceinline__ double log_sigmoid_d(double v) { return v < 0 ? v - log1p(exp(v)) : -log1p(exp(-v)); }

struct NbComp { double r, logits, pi; };
__device__ __forceinline__ NbComp nb_component(float loc, float shape_raw, const float* pi_logit, int c) {
  // models.py:166-191 literally: mean=softplus(loc), shape=softplus(p1), r=1/shape,
  // logits=-log(shape)-log(mean), pi=sigmoid(p2)
  const double a = log1p(exp(-fabs((double)shape_raw))) + fmax((double)shape_raw, 0.0);
  const double mean_net = log1p(exp(-fabs((double)loc))) + fmax((double)loc, 0.0);
  NbComp o;
  o.r = 1.0 / a;
  o.logits = -log(a) - log(mean_net);
  o.pi = pi_logit ? 1.0 / (1.0 + exp(-(double)pi_logit[c])) : 0.0;
  return o;
}

// means[c,n] = distribution mean; ws[0] / ws[1] = global max of mean / stddev (ordered-int floats)
__global__ void __launch_bounds__(256)
nb_stats_kernel(const float* __restrict__ loc, const float* __restrict__ shape_raw,
                const float* __restrict__ pi_logit, int M, int N, float* __restrict__ means,
                float* __restrict__ ws) {
  float mx_mean = -FLT_MAX, mx_sd = -FLT_MAX;
  const size_t total = (size_t)M * N;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i / N);
    const NbComp k = nb_component(loc[i], shape_raw[c], pi_logit, c);
    const double nb_mean = k.r * exp(k.logits);
    const double nb_var = nb_mean / exp(log_sigmoid_d(-k.logits));
    const double mean = (1.0 - k.pi) * nb_mean;
    const double var = (1.0 - k.pi) * (nb_var + nb_mean * nb_mean) - mean * mean;
    means[i] = (float)mean;
    mx_mean = fmaxf(mx_mean, (float)mean);
    mx_sd = fmaxf(mx_sd, (float)sqrt(fmax(var, 0.0)));
  }
  for (int o = 16; o > 0; o >>= 1) {
    mx_mean = fmaxf(mx_mean, __shfl_xor_sync(0xffffffffu, mx_mean, o));
    mx_sd = fmaxf(mx_sd, __shfl_xor_sync(0xffffffffu, mx_sd, o));
  }
  if ((threadIdx.x & 31) == 0) {
    auto enc = [](float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; };
    atomicMax((int*)&ws[0], enc(mx_mean));
    atomicMax((int*)&ws[1], enc(mx_sd));
  }
}
__global__ void nb_stats_init_kernel(float* ws) { ((int*)ws)[0] = (int)0x80000000; ((int*)ws)[1] = (int)0x80000000; }

__device__ double nb_mix_cdf(const float* __restrict__ loc, const float* __restrict__ shape_raw,
                             const float* __restrict__ pi_logit, int M, int N, int n, double kk) {
  double acc = 0.0;
  for (int c = 0; c < M; ++c) {
    const NbComp k = nb_component(loc[(size_t)c * N + n], shape_raw[c], pi_logit, c);
    const double cdf = betainc_d(k.r, 1.0 + kk, log_sigmoid_d(-k.logits), log_sigmoid_d(k.logits));
    acc += k.pi + (1.0 - k.pi) * cdf;
  }
  return acc / (double)M;
}

// exact discrete quantile min{k>=0 : meanCDF(k) >= q} on [0, ceil(high)], high as inference.py:319-324
__global__ void __launch_bounds__(128)
nb_quantile_kernel(const float* __restrict__ loc, const float* __restrict__ shape_raw,
                   const float* __restrict__ pi_logit, int M, int N, const float* __restrict__ ws,
                   float q, float* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float high = dec_ordered(ws[0]) + 1.1f * rsqrtf(1.f - q) * dec_ordered(ws[1]);
  double lo = -1.0, hi = ceil((double)high);
  if (!(hi >= 0.0)) hi = 0.0;
  while (hi - lo > 1.0) {
    const double mid = floor(0.5 * (lo + hi));
    if (nb_mix_cdf(loc, shape_raw, pi_logit, M, N, n, mid) >= (double)q) hi = mid; else lo = mid;
  }
  out[n] = (float)hi;
}

// =============================================================================
// host-side launch wrappers
// =============================================================================
void launch_prep(const DevModel& m, const float* params, float* derived, int n_net, float* zero_acc,
                 float* zero_acc2, int32_t* zero_cursors, cudaStream_t st, float** loss_slot, float* out_loss,
                 float* zero_rows) {
  int nb = 1;
  if (zero_rows) {
    nb = (m.P + 4095) / 4096;                 // ~16 floats per thread
    const int cap = (148 * 8 + n_net - 1) / n_net;
    if (nb > cap) nb = cap;
    if (nb < 1) nb = 1;
  }
  BNF_PROF("prep", st);
  launch_k(prep_kernel, dim3(nb, n_net), dim3(zero_rows ? 256 : 32), 0, st, m, params, derived, n_net, zero_acc,
           zero_acc2, zero_cursors, loss_slot, out_loss, zero_rows);
}

static int sm_count_cached() {     // per device: a process may move between GPUs
  static int sms[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!sms[dev]) {
    cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
    if (sms[dev] <= 0) sms[dev] = 148;
  }
  return sms[dev];
}
// rows per block such that n_net * ceil(B / R) blocks fill a whole number of waves of
// (SM count * blocks_per_sm) resident blocks, with min_rows <= R <= max_rows
static int balanced_rows(int B, int n_net, int blocks_per_sm, int min_rows, int max_rows) {
  const int sms = sm_count_cached();
  const long long slots = (long long)sms * blocks_per_sm;
  for (int w = 1; w <= 1024; ++w) {
    const int nb = (int)(slots * w / n_net);     // blocks per network in w waves
    if (nb < 1) continue;
    const int r = (B + nb - 1) / nb;
    if (r <= max_rows) return r < min_rows ? min_rows : r;
  }
  return max_rows;
}

// rows per block of head_fused_kernel: experiment hook BNF_HEAD_FUSED_MAX_ROWS (default 512)
static int head_fused_max_rows() {
  const char* e = getenv("BNF_HEAD_FUSED_MAX_ROWS");
  const int v = e ? atoi(e) : 0;
  return (v >= 32 && v <= kHeadFusedMaxRows) ? v : kHeadFusedMaxRows;
}

template <typename T>
void launch_encode(const DevModel& m, const float* derived, const float* x, const int32_t* idx,
                   int64_t idx_stride, int B, T* feat, int n_net, cudaStream_t st) {
  if constexpr (FastMath<T>::value) {
    const char* e = getenv("BNF_ENCODE_GENERIC");
    if (!(e && e[0] == '1')) {
      int U = m.D + m.n_seasonal + m.n_inter;
      for (int i = 0; i < m.D; ++i) U += m.fourier_deg[i] > 0 ? m.fourier_deg[i] : 0;
      const size_t smem_f = (size_t)kEncFastRows * (m.Fp * 2 + 16) + (size_t)kEncFastRows * (kMaxD + 1) * 4 +
                            (size_t)U * sizeof(EncUnit);
      if (smem_f <= 48 * 1024) {
        dim3 grid_f((B + kEncFastRows - 1) / kEncFastRows, n_net);
        BNF_PROF("encode", st);
        launch_k(encode_fast_kernel<false>, grid_f, dim3(256), smem_f, st, m, derived, x, idx, idx_stride, B, feat, U);
        return;
      }
    }
  }
  dim3 grid((B + kEncRows - 1) / kEncRows, n_net);
  size_t smem = (size_t)kEncRows * (m.Fp + 1) * sizeof(float);
  BNF_PROF("encode", st);
  launch_k(encode_kernel<T>, grid, dim3(256), smem, st, m, derived, x, idx, idx_stride, B, feat);
}
void launch_encode_x3(const DevModel& m, const float* derived, const float* x, const int32_t* idx,
                      int64_t idx_stride, int B, __nv_bfloat16* feat3, int n_net, cudaStream_t st) {
  {
    const char* e = getenv("BNF_ENCODE_GENERIC");
    int U = m.D + m.n_seasonal + m.n_inter;
    for (int i = 0; i < m.D; ++i) U += m.fourier_deg[i] > 0 ? m.fourier_deg[i] : 0;
    const size_t smem_f = (size_t)kEncFastRows * (m.Fp * 4 + 16) + (size_t)kEncFastRows * (kMaxD + 1) * 4 +
                          (size_t)U * sizeof(EncUnit);
    if (!(e && e[0] == '1') && smem_f <= 48 * 1024) {
      dim3 grid_f((B + kEncFastRows - 1) / kEncFastRows, n_net);
      BNF_PROF("encode", st);
      launch_k(encode_fast_kernel<true>, grid_f, dim3(256), smem_f, st, m, derived, x, idx, idx_stride, B, feat3, U);
      return;
    }
  }
  dim3 grid((B + kEncRows - 1) / kEncRows, n_net);
  size_t smem = (size_t)kEncRows * (m.Fp + 1) * sizeof(float);
  BNF_PROF("encode", st);
  launch_k(encode_kernel<__nv_bfloat16, true>, grid, dim3(256), smem, st, m, derived, x, idx, idx_stride, B, feat3);
}
template void launch_encode<float>(const DevModel&, const float*, const float*, const int32_t*, int64_t, int, float*, int, cudaStream_t);
template void launch_encode<__nv_bfloat16>(const DevModel&, const float*, const float*, const int32_t*, int64_t, int, __nv_bfloat16*, int, cudaStream_t);

void launch_encode_bwd(const DevModel& m, const float* params, const float* derived, const float* x,
                       const int32_t* idx, int64_t idx_stride, int B, const float* dfeat, float* grad,
                       int n_net, bool fast_trig, bool col_major, cudaStream_t st) {
  // dfeat element (net, b, c) sits at net*B*Fp + b*g_row + c*g_col: row-major from the SIMT
  // dgrad, column-major from the tensor-core dgrad epilogue
  const int64_t g_row = col_major ? 1 : m.Fp, g_col = col_major ? B : 1;
  static int occ[2] = {0, 0};
  int& oc = occ[fast_trig ? 1 : 0];
  if (!oc) {
    if (fast_trig) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&oc, encode_bwd_kernel<true>, 256, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&oc, encode_bwd_kernel<false>, 256, 0);
    if (oc < 1) oc = 1;
  }
  const int R = balanced_rows(B, n_net, oc, 64, kEncBwdMaxRows);
  dim3 grid((B + R - 1) / R, n_net);
  BNF_PROF("encode_bwd", st);
  if (fast_trig)
    launch_k(encode_bwd_kernel<true>, grid, dim3(256), 0, st, m, params, derived, x, idx, idx_stride, B, R, dfeat, g_row, g_col, grad);
  else
    launch_k(encode_bwd_kernel<false>, grid, dim3(256), 0, st, m, params, derived, x, idx, idx_stride, B, R, dfeat, g_row, g_col, grad);
}

template <typename T>
void launch_head(const DevModel& m, const float* params, const float* derived, const T* h,
                 const float* y, const int32_t* idx, int64_t idx_stride, int B, float* out_loc,
                 float* opre, float* r, float* ll, float* grad, int n_net, cudaStream_t st) {
  dim3 grid((B + kHeadRows - 1) / kHeadRows, n_net);
  BNF_PROF("head", st);
  launch_k(head_kernel<T>, grid, dim3(256), 0, st, m, params, derived, h, y, idx, idx_stride, B, out_loc, opre, r, ll, grad);
}
void launch_head_x3(const DevModel& m, const float* params, const float* derived, const __nv_bfloat16* h3,
                    const float* y, const int32_t* idx, int64_t idx_stride, int B, float* out_loc,
                    float* opre, float* r, float* ll, float* grad, int n_net, cudaStream_t st) {
  dim3 grid((B + kHeadRows - 1) / kHeadRows, n_net);
  BNF_PROF("head", st);
  launch_k(head_kernel<__nv_bfloat16, 3>, grid, dim3(256), 0, st, m, params, derived, h3, y, idx, idx_stride, B, out_loc, opre, r, ll, grad);
}
template void launch_head<float>(const DevModel&, const float*, const float*, const float*, const float*, const int32_t*, int64_t, int, float*, float*, float*, float*, float*, int, cudaStream_t);
template void launch_head<__nv_bfloat16>(const DevModel&, const float*, const float*, const __nv_bfloat16*, const float*, const int32_t*, int64_t, int, float*, float*, float*, float*, float*, int, cudaStream_t);

// returns false when the shape does not fit the fused kernel (caller uses head + act_bwd)
// one-warp-per-row variant (W = 512 / 1024, bf16 and bf16x3), opt-in with BNF_HEAD_ROWS=1: it moves
// 6 instead of 8 bytes per element but is SLOWER than head_fused_kernel (air-quality 0.341 vs
// 0.329 ms, ZINB 0.507 vs 0.394 ms, wind 2.98 vs 2.44 ms; profiles/experiments r2v) -- the kernel
// is bound by issued instructions, not by the second read of h, and eight warps per SM at ~250
// registers hide less latency.  Kept as the measured answer to "remove the re-read", tested in
// tests/test_gpu_tc.py::test_head_rows_variant_agrees.
constexpr int kHeadRowsMaxRows = 512;
static bool head_rows_wanted(const DevModel& m) {
  if (m.W != 512 && m.W != 1024) return false;
  const char* e = getenv("BNF_HEAD_ROWS");
  return e && e[0] == '1';
}
template <int NC, bool X3>
static void launch_head_rows_t(const DevModel& m, const float* params, const float* derived, const __nv_bfloat16* h,
                               const void* z, const float* y, const int32_t* idx, int64_t idx_stride, int B,
                               __nv_bfloat16* dU, float* ll, float* grad, int n_net, cudaStream_t st) {
  constexpr int RIF = X3 ? 64 / NC : 128 / NC;
  const size_t smem = (size_t)3 * NC * 32 * sizeof(float);
  static int occ = 0, dev_of = -1;
  int dev_now = 0;
  cudaGetDevice(&dev_now);
  if (!occ || dev_of != dev_now) {
    dev_of = dev_now;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, head_rows_kernel<NC, X3>, 256, smem);
    if (occ < 1) occ = 1;
  }
  const char* e = getenv("BNF_HEAD_ROWS_MAX");
  const int mx = e && atoi(e) >= 8 * RIF ? atoi(e) : kHeadRowsMaxRows;
  const int R = balanced_rows(B, n_net, occ, 8 * RIF, mx);
  dim3 grid((B + R - 1) / R, n_net);
  BNF_PROF("head_fused", st);
  launch_k(head_rows_kernel<NC, X3>, grid, dim3(256), smem, st, m, params, derived, h, z, y, idx, idx_stride, B, R, dU, ll, grad);
}

template <typename T>
bool launch_head_fused(const DevModel& m, const float* params, const float* derived, const T* h, const T* z,
                       const float* y, const int32_t* idx, int64_t idx_stride, int B, T* dU, float* ll,
                       float* grad, int n_net, cudaStream_t st) {
  constexpr int VEC = 16 / sizeof(T);
  if (m.W % VEC != 0) return false;
  if constexpr (sizeof(T) == 2) {
    if (head_rows_wanted(m)) {
      if (m.W == 512) launch_head_rows_t<16, false>(m, params, derived, h, (const void*)z, y, idx, idx_stride, B, dU, ll, grad, n_net, st);
      else launch_head_rows_t<32, false>(m, params, derived, h, (const void*)z, y, idx, idx_stride, B, dU, ll, grad, n_net, st);
      return true;
    }
  }
  const int G = m.W / VEC;
  if (G > 256 || 256 % G != 0) return false;
  const size_t smem = (size_t)(kHeadFusedMaxRows + 3 * m.W) * sizeof(float);
  // rows per block: the smallest whole number of waves of resident blocks that keeps R <= max
  static int occ_cache[2] = {0, 0}, w_cache[2] = {0, 0}, dev_cache[2] = {-1, -1};
  int& occ = occ_cache[sizeof(T) == 4 ? 0 : 1];
  int& w_of = w_cache[sizeof(T) == 4 ? 0 : 1];
  int& dev_of = dev_cache[sizeof(T) == 4 ? 0 : 1];
  int dev_now = 0;
  cudaGetDevice(&dev_now);
  if (!occ || w_of != m.W || dev_of != dev_now) {   // the smem attribute is per device
    w_of = m.W;
    dev_of = dev_now;
    if (smem > 48 * 1024) cudaFuncSetAttribute(head_fused_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, head_fused_kernel<T>, 256, smem);
    if (occ < 1) occ = 1;
  }
  const int R = balanced_rows(B, n_net, occ, 32, head_fused_max_rows());
  dim3 grid((B + R - 1) / R, n_net);
  BNF_PROF("head_fused", st);
  launch_k(head_fused_kernel<T>, grid, dim3(256), smem, st, m, params, derived, h, (const void*)z, y, idx, idx_stride, B, R, dU, ll, grad);
  return true;
}
bool head_fused_x3_supported(const DevModel& m) {
  const int G = m.W / 8;
  return m.W % 8 == 0 && G <= 256 && 256 % G == 0;
}
bool launch_head_fused_x3(const DevModel& m, const float* params, const float* derived, const __nv_bfloat16* h3,
                          const float* z, const float* y, const int32_t* idx, int64_t idx_stride, int B,
                          __nv_bfloat16* dU3, float* ll, float* grad, int n_net, cudaStream_t st) {
  if (!head_fused_x3_supported(m)) return false;
  if (head_rows_wanted(m)) {
    if (m.W == 512) launch_head_rows_t<16, true>(m, params, derived, h3, (const void*)z, y, idx, idx_stride, B, dU3, ll, grad, n_net, st);
    else launch_head_rows_t<32, true>(m, params, derived, h3, (const void*)z, y, idx, idx_stride, B, dU3, ll, grad, n_net, st);
    return true;
  }
  const size_t smem = (size_t)(kHeadFusedMaxRows + 3 * m.W) * sizeof(float);
  static int occ = 0, w_of = 0, dev_of = -1;
  int dev_now = 0;
  cudaGetDevice(&dev_now);
  if (!occ || w_of != m.W || dev_of != dev_now) {
    w_of = m.W;
    dev_of = dev_now;
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(head_fused_kernel<__nv_bfloat16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, head_fused_kernel<__nv_bfloat16, true>, 256, smem);
    if (occ < 1) occ = 1;
  }
  const int R = balanced_rows(B, n_net, occ, 32, head_fused_max_rows());
  dim3 grid((B + R - 1) / R, n_net);
  BNF_PROF("head_fused", st);
  launch_k(head_fused_kernel<__nv_bfloat16, true>, grid, dim3(256), smem, st, m, params, derived, h3, (const void*)z, y, idx,
           idx_stride, B, R, dU3, ll, grad);
  return true;
}
template bool launch_head_fused<float>(const DevModel&, const float*, const float*, const float*, const float*, const float*, const int32_t*, int64_t, int, float*, float*, float*, int, cudaStream_t);
template bool launch_head_fused<__nv_bfloat16>(const DevModel&, const float*, const float*, const __nv_bfloat16*, const __nv_bfloat16*, const float*, const int32_t*, int64_t, int, __nv_bfloat16*, float*, float*, int, cudaStream_t);

template <typename T>
void launch_act_bwd(const DevModel& m, int layer, bool is_head, const float* params,
                    const float* derived, const T* z, const T* h, const float* r, T* dU, int B,
                    float* grad, int n_net, cudaStream_t st) {
  constexpr int VEC = 16 / sizeof(T);
  const int G = m.W / VEC;
  if (m.W % VEC == 0 && G <= 256 && 256 % G == 0) {
    dim3 grid((B + kActVecRows - 1) / kActVecRows, n_net);
    if (is_head) {
      BNF_PROF("act_bwd", st);
      launch_k(act_bwd_vec_kernel<T, true>, grid, dim3(256), 2 * m.W * sizeof(float), st, m, layer, params, derived, z, h, r, dU, B, grad);
    } else {
      BNF_PROF("act_bwd", st);
      launch_k(act_bwd_vec_kernel<T, false>, grid, dim3(256), m.W * sizeof(float), st, m, layer, params, derived, z, h, r, dU, B, grad);
    }
    return;
  }
  dim3 grid((m.W + 127) / 128, (B + kActRows - 1) / kActRows, n_net);
  if (is_head) {
    BNF_PROF("act_bwd", st);
    launch_k(act_bwd_kernel<T, true>, grid, dim3(128), 0, st, m, layer, params, derived, z, h, r, dU, B, grad);
  } else {
    BNF_PROF("act_bwd", st);
    launch_k(act_bwd_kernel<T, false>, grid, dim3(128), 0, st, m, layer, params, derived, z, h, r, dU, B, grad);
  }
}
template void launch_act_bwd<float>(const DevModel&, int, bool, const float*, const float*, const float*, const float*, const float*, float*, int, float*, int, cudaStream_t);
template void launch_act_bwd<__nv_bfloat16>(const DevModel&, int, bool, const float*, const float*, const __nv_bfloat16*, const __nv_bfloat16*, const float*, __nv_bfloat16*, int, float*, int, cudaStream_t);

template <typename T>
void launch_fwd_layer_simt_t(const DevModel& m, int layer, const float* params, const float* derived,
                             const T* a_in, int K, int lda, T* z, T* h, int n_net, int B, cudaStream_t st) {
  launch_fwd_layer_simt<T>(m, layer, params, derived, a_in, K, lda, z, h, n_net, B, st);
}
template <typename T, typename TO>
void launch_dgrad_simt_t(const DevModel& m, int layer, const float* params, const T* dU, TO* out,
                         int Kout, int ld_out, int n_net, int B, cudaStream_t st) {
  launch_dgrad_simt<T, TO>(m, layer, params, dU, out, Kout, ld_out, n_net, B, st);
}
template <typename T>
void launch_wgrad_simt_t(const DevModel& m, int layer, const T* a_in, int Kin, int lda, const T* dU,
                         float* grad, int n_net, int B, cudaStream_t st) {
  launch_wgrad_simt<T>(m, layer, a_in, Kin, lda, dU, grad, n_net, B, st);
}
#define BNF_INST(T)                                                                                   \
  template void launch_fwd_layer_simt_t<T>(const DevModel&, int, const float*, const float*, const T*, \
                                           int, int, T*, T*, int, int, cudaStream_t);                 \
  template void launch_dgrad_simt_t<T, T>(const DevModel&, int, const float*, const T*, T*, int, int,  \
                                          int, int, cudaStream_t);                                    \
  template void launch_wgrad_simt_t<T>(const DevModel&, int, const T*, int, int, const T*, float*, int, \
                                       int, cudaStream_t);
BNF_INST(float)
BNF_INST(__nv_bfloat16)
template void launch_dgrad_simt_t<__nv_bfloat16, float>(const DevModel&, int, const float*,
                                                        const __nv_bfloat16*, float*, int, int, int,
                                                        int, cudaStream_t);
#undef BNF_INST

void launch_tick(int32_t* step_count, int32_t* slot, cudaStream_t st) {
  BNF_PROF("tick", st);
  tick_kernel<<<1, 1, 0, st>>>(step_count, slot);
}

void launch_map_adam(int P, float* params, float* am, float* av, const float* g_ll,
                     const int32_t* step_count, float c_ll, float prior_weight, float lr,
                     float* prior_out, int n_net, cudaStream_t st) {
  int bx = (P + 255) / 256;
  if (bx > 1024) bx = 1024;
  BNF_PROF("map_adam", st);
  map_adam_kernel<<<dim3(bx, n_net), 256, 0, st>>>(P, params, am, av, g_ll, step_count, c_ll,
                                                   prior_weight, lr, prior_out);
}
void launch_map_update(const DevModel& m, float* params, float* am, float* av, float* grad,
                       int32_t* step_count, float c_ll, float prior_weight, float lr, float* prior,
                       float* ll, float* const* loss_slot, int32_t* slot, unsigned int* counter, float* derived,
                       __nv_bfloat16* wn, size_t w_per_net, int wn_planes, int n_net, cudaStream_t st) {
  // about one wave of resident blocks in total: every block pays one fence + one ticket atomic
  const int sms = sm_count_cached();
  // blocks per SM: a thread's fixed costs (index set-up, bias-correction hand-off, ticket) are ~250
  // instructions against ~100 per element, so small problems take fewer, longer threads (measured,
  // chickenpox: 8 -> 18.9 us, 4 -> 17.3 us, 2 -> 19.0 us); large ones need the parallelism to keep
  // HBM busy (wind: 8 -> 0.98 ms, 2 -> 1.45 ms).  BNF_UPDATE_BLOCKS_PER_SM overrides.
  static int per_sm_env = -1;
  if (per_sm_env < 0) {
    const char* e = getenv("BNF_UPDATE_BLOCKS_PER_SM");
    per_sm_env = e ? atoi(e) : 0;
    if (per_sm_env < 0 || per_sm_env > 16) per_sm_env = 0;
  }
  const long long total_params = (long long)m.P * n_net;
  const int per_sm = per_sm_env ? per_sm_env : (total_params < 4LL * sms * 8 * 256 ? 4 : 8);
  int bx = (sms * per_sm + n_net - 1) / n_net;
  const int bmax = (m.P + 255) / 256;
  if (bx > bmax) bx = bmax;
  if (bx < 1) bx = 1;
  BNF_PROF("map_update", st);
  // fast (MUFU) update math only in the single-pass bf16 mode; fp32 and bf16x3 keep optax's exact sequence
  if (wn && wn_planes == 1)
    launch_k(map_update_kernel<true>, dim3(bx, n_net), dim3(256), 0, st, m, params, am, av, grad, step_count, c_ll,
             prior_weight, lr, prior, ll, loss_slot, slot, counter, derived, wn, w_per_net, wn_planes, n_net);
  else
    launch_k(map_update_kernel<false>, dim3(bx, n_net), dim3(256), 0, st, m, params, am, av, grad, step_count, c_ll,
             prior_weight, lr, prior, ll, loss_slot, slot, counter, derived, wn, w_per_net, wn_planes, n_net);
}
void launch_map_loss(int n_net, const float* ll, const float* prior, float c_ll, float prior_weight,
                     float* out, const int32_t* slot, cudaStream_t st) {
  BNF_PROF("map_loss", st);
  map_loss_kernel<<<(n_net + 127) / 128, 128, 0, st>>>(n_net, ll, prior, c_ll, prior_weight, out, slot);
}

void launch_vi_sample(int P, int E, int S, const float* mu, const float* rho, const float* eps_in,
                      float* eps_out, uint64_t seed, uint64_t stream_id, const int32_t* step_ptr, float* z, cudaStream_t st) {
  size_t total = ((size_t)S * E * P + 3) / 4;     // four elements per thread
  int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  BNF_PROF("vi_sample", st);
  vi_sample_kernel<<<blocks, 256, 0, st>>>(P, E, S, mu, rho, eps_in, eps_out, seed, stream_id, step_ptr, z);
}
void launch_vi_adam(int P, int E, int S, float* mu, float* rho, float* am, float* av, const float* z,
                    const float* eps, const float* g_ll, const int32_t* step_count, float c, float lr,
                    float* loss_acc, cudaStream_t st) {
  int bx = (P + 255) / 256;
  if (bx > 1024) bx = 1024;
  BNF_PROF("vi_adam", st);
  vi_adam_kernel<<<dim3(bx, E), 256, 0, st>>>(P, E, S, mu, rho, am, av, z, eps, g_ll, step_count, c, lr, loss_acc);
}
void launch_vi_loss(int E, int S, const float* loss_acc, const float* ll, float c, float* out,
                    float* const* out_slot, const int32_t* slot, cudaStream_t st) {
  BNF_PROF("vi_loss", st);
  vi_loss_kernel<<<(E + 127) / 128, 128, 0, st>>>(E, S, loss_acc, ll, c, out, out_slot, slot);
}
void launch_batch_window(int n_total, int B, int steps_per_epoch, int n_rows, uint64_t seed, int64_t first_member,
                         const int32_t* step_count, int32_t* idx_win, cudaStream_t st) {
  int bx = (B + 255) / 256;
  if (bx > 64) bx = 64;
  BNF_PROF("batch_window", st);
  launch_k(batch_window_kernel, dim3(bx, n_rows), dim3(256), 0, st, n_total, B, steps_per_epoch, seed, first_member,
           step_count, idx_win);
}

void launch_init_params(const DevModel& m, float lns_init, uint64_t seed, int64_t first_member,
                        int n_net, float* params, cudaStream_t st) {
  int bx = (m.P + 255) / 256;
  if (bx > 2048) bx = 2048;
  BNF_PROF("init_params", st);
  init_params_kernel<<<dim3(bx, n_net), 256, 0, st>>>(m, lns_init, seed, first_member, params);
}

void launch_quantiles(const float* means, const float* scales, int M, int N, const double* q, int nq,
                      bool approximate, const float* ndtri_q, float* out, float* mm, cudaStream_t st) {
  if (!approximate) {
    BNF_PROF("minmax_init", st);
    minmax_init_kernel<<<1, 1, 0, st>>>(mm);
    size_t n = (size_t)M * N;
    int blocks = (int)((n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024);
    BNF_PROF("minmax", st);
    minmax_kernel<<<blocks, 256, 0, st>>>(means, n, mm);
    BNF_PROF("minmax", st);
    minmax_kernel<<<1, 256, 0, st>>>(scales, (size_t)M, mm + 2);
  }
  for (int i = 0; i < nq; ++i) {
    if (approximate) {
      BNF_PROF("quantile_approx", st);
      quantile_approx_kernel<<<(N + 127) / 128, 128, 0, st>>>(means, scales, M, N, ndtri_q[i], out + (size_t)i * N);
    } else {
      BNF_PROF("quantile_root", st);
      quantile_root_kernel<<<(N + 127) / 128, 128, 0, st>>>(means, scales, M, N, mm, (float)q[i], out + (size_t)i * N);
    }
  }
}

void launch_nb_quantiles(const float* loc, const float* shape_raw, const float* pi_logit, int M, int N,
                         const double* q, int nq, float* means, float* out, float* ws, cudaStream_t st) {
  {
    BNF_PROF("nb_stats_init", st);
    nb_stats_init_kernel<<<1, 1, 0, st>>>(ws);
  }
  {
    size_t total = (size_t)M * N;
    int blocks = (int)((total + 255) / 256 < 2048 ? (total + 255) / 256 : 2048);
    BNF_PROF("nb_stats", st);
    nb_stats_kernel<<<blocks, 256, 0, st>>>(loc, shape_raw, pi_logit, M, N, means, ws);
  }
  for (int i = 0; i < nq; ++i) {
    BNF_PROF("nb_quantile", st);
    nb_quantile_kernel<<<(N + 127) / 128, 128, 0, st>>>(loc, shape_raw, pi_logit, M, N, ws, (float)q[i], out + (size_t)i * N);
  }
}

}  // namespace bnf
