// tcgen05 GEMMs of the dense stack (bf16 operands, f32 accumulation in TMEM) -- single-pass bf16
// and the split-operand bf16x3 mode (f32-class accuracy: three / two bf16 planes per operand, six /
// five products per GEMM into one accumulator; see kX3PlaneA and DESIGN.md section 3.3).
//
// One persistent, warp-specialised kernel template serves the GEMMs of a hidden layer
// (models.py:263-268 forward and its jax.value_and_grad backward), one instantiation per
// epilogue MODE:
//   fwd   : Z[b,n]  = sum_k A[b,k]  * K[k,n]       A K-major, K MN-major   TC_FWD / TC_FWD_HEAD
//   dgrad : dH[b,k] = sum_n dU[b,n] * K[k,n]       dU, K K-major           TC_DGRAD_ACT / TC_DGRAD_ENC
//   wgrad : dK[k,n] = sum_b A[b,k]  * dU[b,n]      both MN-major           TC_WGRAD
// Per CTA: warps 0..kEpi-1 = epilogue (tcgen05.ld -> registers -> fused math -> swizzled smem
// staging -> TMA store), then the TMA producer warp and the TMEM allocator + tcgen05.mma issuer
// warp (highest warp ids = highest issue priority).  smem ring of kStages {A 128x64, B BLOCK_Nx64}
// bf16 tiles in the 128B-swizzle canonical layout written by TMA; two TMEM accumulator stages so
// the epilogue of tile i overlaps the MMAs of tile i+1.  DESIGN.md section 3.1 lists what each
// epilogue fuses and the measurements behind its structure.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "bnf_device.cuh"
#include "bnf_prof.h"
#include "bnf_tc.h"

namespace bnf {

typedef __nv_bfloat16 bf16;

static thread_local char g_tc_err[256] = "";
const char* tc_last_error() { return g_tc_err; }
static int tc_fail(int code, const char* msg) {
  snprintf(g_tc_err, sizeof(g_tc_err), "%s", msg);
  return code;
}

const char* tc_unsupported_reason(const DevModel& m) {
  if (m.W % 64 != 0) return "width must be a multiple of 64";
  if (m.W < 64) return "width must be >= 64";
  return nullptr;
}

static inline int kp_of(const DevModel& m, int layer) { return layer == 0 ? m.Fp : m.W; }
static inline size_t layer_off(const DevModel& m, int layer) {
  return layer == 0 ? 0 : (size_t)m.Fp * m.W + (size_t)(layer - 1) * m.W * m.W;
}
size_t tc_weight_elems(const DevModel& m) { return layer_off(m, m.L); }

// -----------------------------------------------------------------------------
// weight staging: f32 master (in,out) -> bf16 natural [Kp][W] and transposed [W][Kp]
// -----------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cast_weights_kernel(const __grid_constant__ DevModel m, const float* __restrict__ params,
                    bf16* __restrict__ wt, bf16* __restrict__ wn, size_t per_net) {
  __shared__ float tile[32][33];
  pdl_enter(params, wt, wn);
  const int net = blockIdx.z;    // (the layers are looped over here: gridDim.z <= 65535 networks)
  const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int layer = 0; layer < m.L; ++layer) {
    const int Kin = layer == 0 ? m.F : m.W, Kp = layer == 0 ? m.Fp : m.W;
    if (k0 >= Kp) continue;      // block-uniform
    const float* src = params + (size_t)net * m.P + m.off_kernel[layer];
    const size_t base = (size_t)net * per_net + (layer == 0 ? 0 : (size_t)m.Fp * m.W + (size_t)(layer - 1) * m.W * m.W);
    for (int r = ty; r < 32; r += 8) {
      int k = k0 + r, n = n0 + tx;
      float v = (k < Kin && n < m.W) ? src[(size_t)k * m.W + n] : 0.f;
      tile[r][tx] = v;
      if (k < Kp && n < m.W) wn[base + (size_t)k * m.W + n] = __float2bfloat16_rn(v);
    }
    if (wt == nullptr) continue;   // forward reads wn MN-major: no transposed copy
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
      int n = n0 + r, k = k0 + tx;
      if (n < m.W && k < Kp) wt[base + (size_t)n * Kp + k] = __float2bfloat16_rn(tile[tx][r]);
    }
    __syncthreads();
  }
}

// bf16x3 staging: wn3 = [layer][Kp][3*W] per network, plane p of element (k, n) at k*3W + p*W + n
// (natural (in,out) order inside a plane, so both the forward (MN-major) and the dgrad (K-major)
// B operand read it); rows >= fan_in are zero.
__global__ void __launch_bounds__(256)
cast_weights_x3_kernel(const __grid_constant__ DevModel m, const float* __restrict__ params,
                       bf16* __restrict__ wn3, size_t per_net) {
  pdl_enter(params, wn3);
  const int net = blockIdx.z;
  for (int layer = 0; layer < m.L; ++layer) {
    const int Kin = layer == 0 ? m.F : m.W, Kp = layer == 0 ? m.Fp : m.W;
    const float* src = params + (size_t)net * m.P + m.off_kernel[layer];
    bf16* dst = wn3 + 3 * ((size_t)net * per_net + (layer == 0 ? 0 : (size_t)m.Fp * m.W + (size_t)(layer - 1) * m.W * m.W));
    const int total = Kp * m.W;
    for (int e = (blockIdx.x * 256 + threadIdx.x); e < total; e += gridDim.x * 256) {
      const int k = e / m.W, n = e - k * m.W;
      const float v = k < Kin ? src[e] : 0.f;
      bf16* d = dst + (size_t)k * 3 * m.W + n;
      split3_one(v, d, d + m.W, d + 2 * m.W);
    }
  }
}
void tc_cast_weights_x3(const DevModel& m, const float* params, bf16* wn3, int n_net, cudaStream_t st) {
  const int kmax = m.Fp > m.W ? m.Fp : m.W;
  int bx = (kmax * m.W + 255) / 256;
  if (bx > 64) bx = 64;
  BNF_PROF("cast_weights", st);
  launch_k(cast_weights_x3_kernel, dim3(bx, 1, n_net), dim3(256), 0, st, m, params, wn3, tc_weight_elems(m));
}

// wt may be NULL (only the natural-layout copy is needed)
void tc_cast_weights(const DevModel& m, const float* params, bf16* wt, bf16* wn, int n_net, cudaStream_t st) {
  int kmax = m.Fp > m.W ? m.Fp : m.W;
  dim3 grid((m.W + 31) / 32, (kmax + 31) / 32, n_net);
  BNF_PROF("cast_weights", st);
  launch_k(cast_weights_kernel, grid, dim3(256), 0, st, m, params, wt, wn, tc_weight_elems(m));
}

// -----------------------------------------------------------------------------
// PTX wrappers
// -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  long long t0 = 0;
  for (uint32_t spin = 0;; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) return;
    if (spin == 1024) t0 = clock64();
    // a protocol bug must surface as a launch failure, never as a hung GPU
    if (spin > 1024 && (spin & 1023) == 0 && clock64() - t0 > 8000000000LL) __trap();
  }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// ---- CTA-pair (cta_group::2) helpers ------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same smem offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}
// 2-SM TMA load: completion bytes are credited to the LEADER CTA's barrier (peer bit cleared)
__device__ __forceinline__ void tma_load_3d_2sm(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {   // arrives on both CTAs' barrier
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// one lane of a converged warp (PTX elect.sync): keeps the surrounding loop warp-uniform so
// the uniform-datapath tcgen05 / TMA instructions need no per-instruction election
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, %1;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred) : "r"(0xffffffffu));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// registers -> TMEM (32 lanes x 32 columns); the caller issues tmem_st_wait() before re-reading
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
        "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
        "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
// 16-column variants (BNF_EPI16 epilogues: half the live registers per step)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): 128B swizzle, version 1
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address, bits [0,14)
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;     // leading byte offset
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;     // stride byte offset
  d |= (uint64_t)1 << 46;                               // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                               // SWIZZLE_128B
  return d;
}

// -----------------------------------------------------------------------------
// the GEMM kernel
// -----------------------------------------------------------------------------
enum { TC_FWD = 0, TC_DGRAD_BF16 = 1, TC_DGRAD_F32 = 2, TC_WGRAD = 3, TC_PLAIN_F32 = 4, TC_DGRAD_ACT = 5,
       TC_DGRAD_ENC = 6 /* layer-0 dgrad fused with the feature-encode backward */,
       TC_FWD_HEAD = 7 /* last hidden layer fwd + head + log-likelihood + its own activation backward */ };

struct TcArgs {
  int mode, n_net;
  int m_tiles, n_tiles, k_splits, k_blocks;   // k_blocks: 64-wide reduction blocks in total
  int m_valid, n_valid;                       // rows / cols of the output that exist
  // epilogue
  const float* params; const float* derived; int P, off_bias, layer; float isf;
  bf16* out0; bf16* out1; float* outf;
  long long out_batch; int ld_out; int grad_off;
  // A_MODE 2 (fused encode + Dense_0): raw inputs instead of a feature matrix
  const float* x; const int32_t* idx; long long idx_stride; int write_feat;
  // TC_DGRAD_ACT (dgrad fused with the activation backward of the previous layer)
  const bf16* zin; float* gradp; int off_bias_prev, off_ls_prev, off_actw, layer_prev;
  int out_cm; // TC_DGRAD_F32: write outf column-major [net][col][row]
  const float* y; float* ll;   // TC_FWD_HEAD: observations, per-network log-likelihood accumulators
  // split-operand (bf16x3) mode: every operand tensor holds three bf16 planes side by side along
  // its contiguous dimension (plane p at columns [p*pstride, (p+1)*pstride)); the reduction runs over
  // six segments of kseg k-blocks, one per plane pair (see kX3PlaneA / kX3PlaneB)
  int x3, kseg, a_pstride, b_pstride;
  uint32_t x3_pa, x3_pb;     // nibble s = plane of A / B multiplied in segment s (k_blocks = segments * kseg)
  int x3_nseg;               // > 0: segments INTERLEAVED per k-block (k-block kb = segment kb % nseg of reduction block kb / nseg)
  // Dense_0 bias gradient through the wgrad GEMM: the encoders write a constant-one feature into
  // the first pad column F (< Fp) of `feat`, so row F of the Dense_0 wgrad accumulator is
  // isf * sum_b dU_0[b, :] = the bias gradient.  TC_WGRAD (layer 0): bias_row = F, that row goes
  // (rescaled by 1/isf) to grad + bias_off instead of the kernel leaf.  TC_DGRAD_ACT producing
  // dU_0: skip_bias = 1, its epilogue drops the 31-shuffle column sums.
  int bias_row, bias_off, skip_bias; float bias_rescale;
  int dbg;   // BNF_TC_DBG ablation mask (only read when compiled with -DBNF_TC_EXPERIMENT)
  long long* tl;   // BNF_TC_TL timeline buffer [cta][16 tiles][16 events] of clock64 (experiment builds)
};
// Epilogue ablation hooks for scripts/epi_experiment.py: compiled out unless -DBNF_TC_EXPERIMENT.
// bits: 1 skip z loads (dgrad), 2 skip bias column sums (dgrad), 4 skip TMA stores, 8 skip the
// activation math.  Results are wrong with a non-zero mask - timing only.
#ifdef BNF_TC_EXPERIMENT
#define DBG(bit) (a.dbg & (bit))
// timeline stamp: event `ev` of this CTA's `ti`-th tile (one lane of one warp per event id)
#define TL(ti, ev) do { if (a.tl && (ti) < 16) a.tl[((size_t)blockIdx.x * 16 + (ti)) * 16 + (ev)] = clock64(); } while (0)
#else
#define DBG(bit) 0
#define TL(ti, ev) do { } while (0)
#endif

// Epilogue warps per mode (ids 0..kEpi-1; kEpi/4 warps share one TMEM lane quarter and interleave
// its 32-column chunks).  The epilogue warps run at ~0.2 IPC each (dependent MUFU/FMA chains, TMEM /
// mbarrier / proxy-fence latencies), so warps per SM sub-partition are what hides latency; the
// register file caps the count (12 warps <-> 146 registers/thread, 16 <-> 112 = spills).
// Measured (profiles/experiments/README.md): TC_DGRAD_ACT with 12 warps: wind dgrad -8 %,
// chickenpox -4 % (one ring stage less); TC_FWD / TC_FWD_HEAD: no change (HBM-write / MUFU bound).
// TC_FWD_HEAD pass 2: the chunk loop is a real loop and the per-lane column sums live in shared memory
// (default; measured r2y: chickenpox bf16 0.1613 -> 0.1569 ms/step, bf16x3 0.3851 -> 0.3771).  With statically
// indexed register sums the loop had to be unrolled, a third of the kernel's code, and ncu attributed 13 %
// (bf16) / 28 % (bf16x3) of the kernel's warp samples to instruction-fetch stalls.  -DBNF_HEAD_UNROLLED
// restores the unrolled form (the sixteen-warp experiment build keeps it as well).
#if !defined(BNF_HEAD_UNROLLED) && !defined(BNF_EPI16)
#define BNF_HEAD_ROLLED 1
#endif
#ifdef BNF_EPI16
// EXPERIMENT (not validated on hardware yet): sixteen epilogue warps
// (four per SM sub-partition) for the three activation epilogues.  To fit 576 threads in the
// register file (<= 112 registers) every 32-column chunk is processed as two 16-column halves
// (tcgen05.ld/st .x16, 16-value transposes); the staging tiles, TMA stores and the z ring keep
// their 32x32 shape.
constexpr bool kEpi16 = true;
__host__ __device__ constexpr int epi_warps_of(int mode, int a_mode) {
  return ((mode == 5 /*TC_DGRAD_ACT*/ || mode == 7 /*TC_FWD_HEAD*/ || mode == 0 /*TC_FWD*/) && a_mode != 2) ? 16 : 8;
}
#elif defined(BNF_FWD_EPI12)   // experiment: three warps per quarter for the plain forward epilogue as well
constexpr bool kEpi16 = false;
__host__ __device__ constexpr int epi_warps_of(int mode, int a_mode) {
  return (mode == 5 /*TC_DGRAD_ACT*/ || (mode == 0 /*TC_FWD*/ && a_mode != 2)) ? 12 : 8;
}
#else
constexpr bool kEpi16 = false;
__host__ __device__ constexpr int epi_warps_of(int mode, int /*a_mode*/) { return mode == 5 /*TC_DGRAD_ACT*/ ? 12 : 8; }
#endif
// bf16x3 (split-operand, f32-parity) mode.  An f32 value a is carried as a0 + a1 + a2 with
// a0 = bf16(a), a1 = bf16(a - a0), a2 = bf16(a - a0 - a1) (residual <= 2^-27 |a|); the f32 product
// sum is the sum of the six bf16 GEMMs a_i.b_j with i + j <= 2 (dropped terms <= 2^-27).  They run
// as ONE reduction of six segments into one TMEM accumulator, smallest products first: tcgen05
// accumulates with round-toward-zero (measured, profiles/experiments/README.md r2a-2), so the
// accumulator should be small while the many small addends arrive.  Segment s multiplies plane
// kX3PlaneA[s] of A with plane kX3PlaneB[s] of B: (2,0) (0,2) (1,1) (1,0) (0,1) (0,0).
constexpr uint32_t kX3PlaneA = 0x001102u, kX3PlaneB = 0x010120u;   // nibble s = plane of segment s
// The backpropagated dU tensors carry only TWO planes (d0 + d1, residual <= 2^-18 |d|: gradients are
// judged at 1e-4 of a leaf's scale, and a leaf is a sum over the batch of independently rounded
// terms), which drops one of the six products and a third of the dU traffic in every backward GEMM:
//   wgrad  (A = activations, 3 planes; B = dU, 2 planes): (2,0) (1,1) (1,0) (0,1) (0,0)
//   dgrad  (A = dU, 2 planes; B = kernels, 3 planes):     (0,2) (1,1) (1,0) (0,1) (0,0)
constexpr int kX3BwdSegs = 5, kDuPlanes = 2;
constexpr uint32_t kX3WgradA = 0x00112u, kX3WgradB = 0x01010u, kX3DgradA = 0x00110u, kX3DgradB = 0x01012u;
constexpr int kEncWarps = 4;                       // A_MODE 2 only: feature-encoder warps (one thread per tile row)
constexpr int kEncMaxKb = 2;                       // A_MODE 2: Fp <= 128 = at most two 64-wide k-blocks
constexpr int kEncMaxUnits = 128;                  // A_MODE 2: every unit owns >= 1 of the <= 128 feature columns
struct EncUnitT { float mult, k0; int kind, dim, dim2, c0, c1, pad; };        // one feature unit (32 bytes)

constexpr int kBarBytes = kEpi16 ? 1024 : 512;     // mbarriers + TMEM base slot
constexpr int kZRing = 2;                          // TC_DGRAD_ACT: per-warp ring of 32x32 z tiles
constexpr int kAccCols = 1024;                     // TC_DGRAD_ACT: widest layer whose bias sums stay in smem
// A_MODE: 0 = A,B K-major by TMA; 1 = A,B MN-major by TMA; 2 = A generated in two resident smem
// buffers by encoder warps from the raw input rows (fused models.py:216-252 encode + Dense_0), B
// MN-major; 3 = A K-major, B MN-major (forward straight from the natural (in,out) bf16 kernel copy,
// so the transposed staging copy and its cast kernel are not needed).
// CTA2: a pair of CTAs (one TPC) computes a 256 x BLOCK_N tile with tcgen05.mma.cta_group::2:
// each CTA stages its own 128 A rows and HALF of the B tile, so operand traffic per FLOP
// from L2 drops by a third and the ring gets deeper (32 KB stages).
template <int BLOCK_N, int A_MODE = 0, bool CTA2 = false, int MODE = 0, bool X3 = false> struct TcCfg {
  static_assert(!X3 || MODE == TC_FWD || MODE == TC_FWD_HEAD || MODE == TC_DGRAD_ACT || MODE == TC_DGRAD_ENC,
                "x3 epilogues: fwd, fwd+head, dgrad+act, dgrad0+encode");
  static_assert(!X3 || !kEpi16, "the x3 epilogues have no sixteen-warp variant");
#ifdef BNF_HEAD_ROLLED
  static_assert(!kEpi16, "BNF_HEAD_ROLLED is written for the eight-warp TC_FWD_HEAD epilogue");
#endif
  // TC_DGRAD_ENC: the dfeat tile stays on chip -- a 128 x (BLOCK_N+1) f32 tile in shared memory.
  // Its epilogue (the encode backward) is latency-bound, so the kernel is sized for TWO CTAs per
  // SM at Fp = 64: two ring stages, no TMA-store staging tiles (~100 KB, 128 TMEM columns each).
  static constexpr int kGBytes = MODE == TC_DGRAD_ENC ? 128 * (BLOCK_N + 1) * 4 + 2 * 128 * (kMaxD + 1) * 4 : 0;
  // TC_FWD_HEAD: Dense_L kernel [2][256] + per-row partial dots [<= 4][128] in shared memory
  // + the tile's h = act(z) as bf16 in per-warp 64B-swizzled 32x32 tiles (pass 1 -> pass 2)
  static constexpr int kHeadScratch = (MODE == TC_FWD_HEAD && !X3)     // (x3 recomputes instead of stashing h)
      ? epi_warps_of(MODE, A_MODE) * ((BLOCK_N / 32 + epi_warps_of(MODE, A_MODE) / 4 - 1) / (epi_warps_of(MODE, A_MODE) / 4)) * 2048 : 0;
  // BNF_HEAD_ROLLED (default): the per-lane column sums of TC_FWD_HEAD's pass 2 live in shared memory
  // ([warp][chunk][2][32] floats, lane-private words) instead of statically indexed registers
#ifdef BNF_HEAD_ROLLED
  static constexpr int kHeadColBytes = MODE == TC_FWD_HEAD
      ? (X3 ? 8 : epi_warps_of(MODE, A_MODE)) * ((BLOCK_N / 32 + (X3 ? 8 : epi_warps_of(MODE, A_MODE)) / 4 - 1) / ((X3 ? 8 : epi_warps_of(MODE, A_MODE)) / 4)) * 64 * 4 : 0;
#else
  static constexpr int kHeadColBytes = 0;
#endif
  static constexpr int kHeadBytes = MODE == TC_FWD_HEAD ? (2 * 256 + 4 * 128) * 4 + kHeadScratch + kHeadColBytes : 0;
  // per epilogue warp: two 32x32 bf16 tiles (TMA-store staging); TC_DGRAD_ACT: one output tile
  // plus a ring of kZRing z tiles landed by TMA
  // X3: TC_FWD [z f32 4 KB | h planes 3 x 2 KB]; TC_DGRAD_ACT [z ring kZRing x 4 KB | dU planes 3 x 2 KB]
  static constexpr int kStgWarp = X3 ? (MODE == TC_DGRAD_ACT ? 4096 * kZRing + kDuPlanes * 2048 : MODE == TC_FWD_HEAD ? kDuPlanes * 2048 : 4096 + 3 * 2048)
                                     : (MODE == TC_DGRAD_ACT ? 2048 * (1 + kZRing) : 4096);
  static constexpr int kEpiW = X3 ? 8 : epi_warps_of(MODE, A_MODE);
  static constexpr int kStagesBf16 = MODE == TC_DGRAD_ENC ? 2 :
      MODE == TC_FWD_HEAD ? (kEpiW > 12 ? (CTA2 ? 2 : (BLOCK_N == 256 ? 1 : (BLOCK_N == 128 ? 3 : 4)))
                                        : (CTA2 ? 3 : (BLOCK_N == 256 ? 2 : 4))) :
      MODE == TC_DGRAD_ACT ? (kEpiW > 12 ? (CTA2 ? 3 : (BLOCK_N == 256 ? 2 : (BLOCK_N == 128 ? 3 : 5))) :
                              kEpiW > 8 ? (CTA2 ? 4 : (BLOCK_N == 256 ? 3 : (BLOCK_N == 128 ? 4 : 6)))
                                        : (CTA2 ? 5 : (BLOCK_N == 256 ? 3 : (BLOCK_N == 128 ? 5 : 7)))) :
      A_MODE == 2 ? (BLOCK_N == 256 ? 3 : 4) :
      (kEpiW > 12 ? (CTA2 ? 5 : (BLOCK_N == 256 ? 3 : (BLOCK_N == 128 ? 5 : 6))) :
       kEpiW > 8 ? (CTA2 ? 5 : (BLOCK_N == 256 ? 3 : (BLOCK_N == 128 ? 5 : 7)))
                 : (CTA2 ? 6 : (BLOCK_N == 256 ? 4 : (BLOCK_N == 128 ? 6 : 8))));
  static constexpr int kEpi = kEpiW;
  static constexpr int kABytes = 128 * 64 * 2;
  static constexpr int kBBytes = (CTA2 ? BLOCK_N / 2 : BLOCK_N) * 64 * 2;
  // A_MODE 2: the ring holds only B tiles; the generated A tiles live in two resident buffers
  static constexpr int kStageBytes = A_MODE == 2 ? kBBytes : kABytes + kBBytes;
  static constexpr int kBOff = A_MODE == 2 ? 0 : kABytes;       // B tile inside a ring stage
  // X3 (bigger staging tiles): as many ring stages as fit beside them, at most 6
  static constexpr int kX3Fixed = kEpi * kStgWarp + kBarBytes + 2 * 256 * 4 + (MODE == TC_DGRAD_ACT ? (kAccCols + 32) * 4 : 0) +
                                  (MODE == TC_FWD_HEAD ? (2 * 256 + 4 * 128) * 4 + kHeadColBytes : 0);
  static constexpr int kStagesX3 = (232448 - kX3Fixed) / kStageBytes > 6 ? 6 : (232448 - kX3Fixed) / kStageBytes;
  static constexpr int kStages = (X3 && MODE != TC_DGRAD_ENC) ? kStagesX3 : kStagesBf16;
  static_assert(kStages >= 2, "at least two ring stages");
  static constexpr int kThreads = 64 + 32 * kEpi + (A_MODE == 2 ? 32 * kEncWarps : 0);
  // A_MODE 2: [2 m-tile buffers][kEncMaxKb k-blocks][16 KB A tile] + unit table + per-row scaled inputs
  static constexpr int kXBytes = A_MODE == 2 ? 2 * kEncMaxKb * kABytes + kEncMaxUnits * 32 + 128 * (kMaxD + 1) * 4 : 0;
  static constexpr int kStagingBytes = MODE == TC_DGRAD_ENC ? 0 : kEpi * kStgWarp;
  // CTA-wide partial sums of the epilogue's column / scalar gradients (flushed to HBM when the
  // network changes): TC_DGRAD_ACT [kAccCols] bias columns + scalars
  static constexpr int kAccFloats = MODE == TC_DGRAD_ACT ? kAccCols + 32 : 0;
  static constexpr int kMinBlocks = (MODE == TC_DGRAD_ENC && BLOCK_N == 64) ? 2 : 1;
  static constexpr int kSmem = kStages * kStageBytes + kStagingBytes + kBarBytes + 2 * 256 * 4 /*bias*/ + kXBytes + kGBytes + kHeadBytes + kAccFloats * 4;
  static_assert(kSmem <= 232448, "dynamic shared memory exceeds the 227 KB per-CTA limit");
  static constexpr int kTmemCols = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
};

template <int BLOCK_N, int A_MODE, int MODE, bool CTA2, bool X3 = false>
__global__ void __launch_bounds__((TcCfg<BLOCK_N, A_MODE, CTA2, MODE, X3>::kThreads), (TcCfg<BLOCK_N, A_MODE, CTA2, MODE, X3>::kMinBlocks))
tc_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_o0, const __grid_constant__ CUtensorMap map_o1,
               const __grid_constant__ TcArgs a, const __grid_constant__ DevModel dm) {
  using Cfg = TcCfg<BLOCK_N, A_MODE, CTA2, MODE, X3>;
  constexpr bool A_MN = A_MODE == 1;                  // A operand MN-major
  constexpr bool B_MN = A_MODE != 0;                  // B operand MN-major (natural (in,out) kernel copy / dU)
  constexpr bool ENCODE = A_MODE == 2;
  static_assert(!(CTA2 && ENCODE), "the fused encode kernel is single-CTA");
  constexpr int kEpi = Cfg::kEpi;                    // epilogue warps (ids 0..kEpi-1)
  constexpr int kParts = kEpi / 4;                   // warps per TMEM lane quarter = column interleave
  constexpr int kProd = kEpi, kMma = kEpi + 1;       // producer / MMA issuer: the highest warp ids
  constexpr int kBaseThreads = 64 + 32 * kEpi;
  // TMEM accumulator stages: the epilogue of tile i overlaps the MMAs of tile i+1
  constexpr int kAccStages = 2;
  const uint32_t cta_rank = CTA2 ? cluster_ctarank() : 0u;   // 0 = leader (issues the MMAs)
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();     // 128B-swizzle tiles need 1024B alignment
  // A_MODE 2: the two resident A buffers sit between the ring and the staging tiles (1024-byte aligned)
  constexpr int kABufBytes = ENCODE ? 2 * kEncMaxKb * Cfg::kABytes : 0;
  uint8_t* abuf = smem + Cfg::kStages * Cfg::kStageBytes;   // [2][kEncMaxKb][16 KB]
  uint8_t* staging = abuf + kABufBytes;
  uint64_t* bars = (uint64_t*)(staging + Cfg::kStagingBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::kStages;
  uint64_t* tfull = bars + 2 * Cfg::kStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = (uint32_t*)(tempty + 2);
  uint64_t* aready = bars + 32;                      // A_MODE 2: [2] generated A tiles of an m-tile complete
  uint64_t* afree = bars + 34;                       // A_MODE 2: [2] the MMAs reading that buffer have retired
  uint64_t* zbar = bars + 40;                        // TC_DGRAD_ACT: [kEpi][kZRing] z tile landed
  float* sbias = (float*)(staging + Cfg::kStagingBytes + kBarBytes);  // [2][256]: s_l * bias of the tile's columns
  float* colacc = (float*)(smem + Cfg::kSmem - Cfg::kAccFloats * 4);  // TC_DGRAD_ACT / TC_FWD_HEAD partial sums
  EncUnitT* etab = (EncUnitT*)(sbias + 2 * 256);     // A_MODE 2: unit table of the current network
  float* esx = (float*)(etab + kEncMaxUnits);        // A_MODE 2: [128][kMaxD+1] scaled inputs (+ raw time) of the tile rows
  float* gtile = sbias + 2 * 256;                    // TC_DGRAD_ENC: [128][BLOCK_N+1] f32 dfeat tile
  float* eacc = sbias;                               // TC_DGRAD_ENC: [2*kMaxD+3] partial sums
  float* kos_s = sbias + 2 * 256;                    // TC_FWD_HEAD: [2][256] Dense_L kernel of the tile's network
  uint8_t* hscr = (uint8_t*)(kos_s + 2 * 256 + 4 * 128);   // TC_FWD_HEAD: [kEpi][chunks][2 KB] bf16 h tiles
  float* rowdot = kos_s + 2 * 256;                   // TC_FWD_HEAD: [kParts][128] partial h.Ko of the quarter's warps
#ifdef BNF_HEAD_ROLLED
  float* hcol_s = reinterpret_cast<float*>(hscr + Cfg::kHeadScratch);   // [kEpi][chunks][2][32] column sums
#endif
  float* sxt = gtile + 128 * (BLOCK_N + 1);          // TC_DGRAD_ENC: [2][128][kMaxD+1] scaled inputs + raw time

  // the warp index through a shuffle: the compiler then knows it (and every address / coordinate
  // derived from it) is warp-uniform, so the TMA / mbarrier instructions take uniform-register
  // operands directly instead of a per-lane R2UR loop in front of every one of them
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  pdl_trigger();                       // the next kernel's CTAs may start their own prologue
  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full[s], CTA2 ? 2 : 1);
      mbar_init(&empty[s], 1);
    }
    if (ENCODE) for (int s = 0; s < 2; ++s) { mbar_init(&aready[s], 32 * kEncWarps); mbar_init(&afree[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], (CTA2 ? 2 : 1) * 32 * kEpi); }
    if (MODE == TC_DGRAD_ACT) for (int s = 0; s < kEpi * kZRing; ++s) mbar_init(&zbar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < Cfg::kAccFloats; i += Cfg::kThreads) colacc[i] = 0.f;
#ifdef BNF_HEAD_ROLLED
  for (int i = threadIdx.x; i < Cfg::kHeadColBytes / 4; i += Cfg::kThreads) hcol_s[i] = 0.f;
#endif
  if (warp == kMma) {
    if (CTA2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                   ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)Cfg::kTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                   ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)Cfg::kTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  if (ENCODE && warp > kMma) {
    // zero the A buffers once: pad columns [F, Fp) are never written again
    const int et = threadIdx.x - kBaseThreads;
    uint4* pz = reinterpret_cast<uint4*>(abuf);
    for (int i = et; i < 2 * kEncMaxKb * Cfg::kABytes / 16; i += 32 * kEncWarps) pz[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  // the TMA descriptors into the descriptor cache (kernel parameters: nothing the previous kernel
  // writes), so the first load / store of every role does not wait for its descriptor fetch
  if (warp == kProd && lane == 0) {
    if (!ENCODE) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
  }
  if (warp == 0 && lane == 0) {
    if ((MODE == TC_FWD && a.out0) || MODE == TC_FWD_HEAD || MODE == TC_DGRAD_ACT || MODE == TC_DGRAD_BF16)
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_o0)) : "memory");
    if (MODE == TC_FWD || MODE == TC_DGRAD_ACT)
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_o1)) : "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync_all();        // both CTAs' barriers are initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work items: (net, m unit, split, n tile); a unit is one 128-row tile, or a PAIR of them for
  // a CTA pair (this CTA takes rows of tile 2*unit + cta_rank)
  const int m_units = CTA2 ? (a.m_tiles + 1) / 2 : a.m_tiles;
  const int tiles_per_net = m_units * a.n_tiles * a.k_splits;
  const int total_tiles = a.n_net * tiles_per_net;
  const int kb_per_split = (a.k_blocks + a.k_splits - 1) / a.k_splits;
  // tile -> CTA mapping: round-robin, except the epilogues that keep per-network partial sums on
  // chip or per-network constants staged (TC_DGRAD_ENC, TC_FWD_HEAD, single-n-tile TC_DGRAD_ACT /
  // TC_FWD): there a CTA takes a CONTIGUOUS
  // range of tiles, i.e. mostly one network, and flushes its sums only when the network changes
  const int cta_id = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int n_ctas = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const bool contiguous = MODE == TC_DGRAD_ENC || MODE == TC_FWD_HEAD || ENCODE ||
                          ((MODE == TC_DGRAD_ACT || MODE == TC_FWD) && a.n_tiles == 1);
  const int tile0 = contiguous ? (int)((long long)total_tiles * cta_id / n_ctas) : cta_id;
  const int tile_end = contiguous ? (int)((long long)total_tiles * (cta_id + 1) / n_ctas) : total_tiles;
  const int tile_step = contiguous ? 1 : n_ctas;

  pdl_wait();                          // everything above overlapped the previous kernel's tail
  if (threadIdx.x == 0) TL(0, 15);


  if (warp == kProd) {
    // ===================== TMA producer (whole warp loops, one elected lane issues) =====================
    {
      int stage = 0; uint32_t phase = 0;
      for (int t = tile0; t < tile_end; t += tile_step) {
        const int net = t / tiles_per_net;
        int r = t % tiles_per_net;
        const int n_t = r % a.n_tiles; r /= a.n_tiles;
        // (the m unit varies faster than the split: the CTAs running side by side then share one
        // block of reduction rows across all their m- and n-tiles, so wgrad's operand re-reads hit L2)
        const int split = r / m_units;
        const int m_t = CTA2 ? 2 * (r % m_units) + (int)cta_rank : r % m_units;
        const int kb0 = split * kb_per_split;
        const int kb1 = min(a.k_blocks, kb0 + kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (lane == 0 && kb == kb0) TL((t - tile0) / tile_step, 0);
          if (lane == 0 && kb == kb1 - 1) TL((t - tile0) / tile_step, 1);
          // x3: k-block kb of the six-segment reduction = k-block kk of plane pair (pa, pb)
          int kk = kb, ca = 0, cb = 0;
          if (a.x3) {
            int seg;
            if (a.x3_nseg > 0) { kk = kb / a.x3_nseg; seg = kb - kk * a.x3_nseg; }   // wgrad: planes of one row block back to back
            else { seg = kb / a.kseg; kk = kb - seg * a.kseg; }
            ca = (int)((a.x3_pa >> (4 * seg)) & 3u) * a.a_pstride;
            cb = (int)((a.x3_pb >> (4 * seg)) & 3u) * a.b_pstride;
          }
          if (elect_one()) {
            uint8_t* sa = smem + stage * Cfg::kStageBytes;
            uint8_t* sb = sa + Cfg::kABytes;
            if (ENCODE) {
              // only the B tile travels (natural (in,out) kernel copy, MN-major); A is generated on chip
              mbar_arrive_expect_tx(&full[stage], Cfg::kBBytes);
              for (int j = 0; j < BLOCK_N / 64; ++j)
                tma_load_3d(sa + j * 8192, &map_b, &full[stage], n_t * BLOCK_N + j * 64, kb * 64, net);
            } else if (CTA2) {
              // the LEADER's full barrier collects both CTAs' bytes (count 2: its own
              // arrive.expect_tx + the peer's remote arrive)
              if (cta_rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * Cfg::kStageBytes);
              else mbar_arrive_remote(&full[stage], 0);
              const int nb = n_t * BLOCK_N + (int)cta_rank * (BLOCK_N / 2);   // this CTA's half of B
              if (!A_MN) {
                tma_load_3d_2sm(sa, &map_a, &full[stage], kk * 64 + ca, m_t * 128, net);
              } else {
                for (int j = 0; j < 2; ++j)
                  tma_load_3d_2sm(sa + j * 8192, &map_a, &full[stage], m_t * 128 + j * 64 + ca, kk * 64, net);
              }
              if (!B_MN) {
                tma_load_3d_2sm(sb, &map_b, &full[stage], kk * 64 + cb, nb, net);
              } else {
                for (int j = 0; j < BLOCK_N / 128; ++j)
                  tma_load_3d_2sm(sb + j * 8192, &map_b, &full[stage], nb + j * 64 + cb, kk * 64, net);
              }
            } else {
              mbar_arrive_expect_tx(&full[stage], Cfg::kStageBytes);
              if (!A_MN) {
                tma_load_3d(sa, &map_a, &full[stage], kk * 64 + ca, m_t * 128, net);
              } else {   // MN-major: boxes of [64 reduction rows][64 MN elements]
                for (int j = 0; j < 2; ++j)
                  tma_load_3d(sa + j * 8192, &map_a, &full[stage], m_t * 128 + j * 64 + ca, kk * 64, net);
              }
              if (!B_MN) {
                tma_load_3d(sb, &map_b, &full[stage], kk * 64 + cb, n_t * BLOCK_N, net);
              } else {
                for (int j = 0; j < BLOCK_N / 64; ++j)
                  tma_load_3d(sb + j * 8192, &map_b, &full[stage], n_t * BLOCK_N + j * 64 + cb, kk * 64, net);
              }
            }
          }
          __syncwarp();
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == kMma) {
    // ===================== MMA issuer (whole warp loops, one elected lane issues) =====================
    if (!CTA2 || cta_rank == 0) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): f32 accum, bf16 x bf16
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) |
                             ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BLOCK_N >> 3) << 17) |
                             ((uint32_t)((CTA2 ? 256 : 128) >> 4) << 24);
      // smem descriptors differ only in the 14-bit start address: build them once and add offsets
      const uint64_t adesc0 = A_MN ? make_smem_desc(smem_u32(smem), 8192, 1024)
                                   : make_smem_desc(smem_u32(smem), 16, 1024);
      const uint64_t bdesc0 = B_MN ? make_smem_desc(smem_u32(smem) + Cfg::kBOff, 8192, 1024)
                                   : make_smem_desc(smem_u32(smem) + Cfg::kBOff, 16, 1024);
      // A_MODE 2: A comes from the resident buffer of the tile's m-tile (two buffers, alternating)
      const uint64_t aenc0 = make_smem_desc(smem_u32(abuf), 16, 1024);
      int enc_buf = 1, enc_key = -1; uint32_t enc_phase[2] = {0u, 0u};
      constexpr uint32_t kStepA = (A_MN ? 2048 : 32) >> 4;      // one UMMA_K=16 slice
      constexpr uint32_t kStepB = (B_MN ? 2048 : 32) >> 4;
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int t = tile0; t < tile_end; t += tile_step) {
        int r = t % tiles_per_net;
        r /= a.n_tiles;
        const int split = r / m_units;
        const int kb0 = split * kb_per_split;
        const int kb1 = min(a.k_blocks, kb0 + kb_per_split);
        bool enc_last = false;
        if constexpr (ENCODE) {
          // tiles of one (network, m-tile) are consecutive (n fastest): the first waits for the
          // encoder warps, the last releases the buffer once its MMAs have retired
          const int key = t / a.n_tiles;
          if (key != enc_key) {
            enc_key = key;
            enc_buf ^= 1;
            mbar_wait(&aready[enc_buf], enc_phase[enc_buf]);
            enc_phase[enc_buf] ^= 1u;
            tc_fence_after();
          }
          enc_last = t + 1 >= tile_end || (t + 1) / a.n_tiles != key;
        }
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        if (lane == 0) TL((t - tile0) / tile_step, 2);
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          if (lane == 0 && kb == kb0) TL((t - tile0) / tile_step, 3);
          if (lane == 0 && kb == kb1 - 1) TL((t - tile0) / tile_step, 4);
          if (elect_one()) {
            const uint64_t so = (uint64_t)((stage * Cfg::kStageBytes) >> 4);
            const uint64_t ao = ENCODE ? aenc0 + (uint64_t)(((enc_buf * kEncMaxKb + kb) * Cfg::kABytes) >> 4) : adesc0 + so;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (CTA2) umma_bf16_2sm(d_tmem, ao + k * kStepA, bdesc0 + so + k * kStepB, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
              else umma_bf16(d_tmem, ao + k * kStepA, bdesc0 + so + k * kStepB, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            }
            if (CTA2) umma_commit_2sm(&empty[stage]);   // frees the slot in BOTH CTAs
            else umma_commit(&empty[stage]);            // frees the smem slot when these MMAs retire
          }
          __syncwarp();
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) {                            // accumulator complete -> epilogue(s)
          if (CTA2) umma_commit_2sm(&tfull[acc]); else umma_commit(&tfull[acc]);
          if (ENCODE && enc_last) umma_commit(&afree[enc_buf]);   // the encoder may refill this A buffer
        }
        __syncwarp();
        if (lane == 0) TL((t - tile0) / tile_step, 5);
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp < kEpi) {
    // ===================== epilogue (warps 0..kEpi-1) =====================
    // warp%4 selects the TMEM lane quarter it may read; the kParts warps of a quarter take
    // interleaved 32-column chunks, so every SM sub-partition has two or three epilogue warps to
    // hide tcgen05.ld / MUFU / store latency behind each other.
    const int q = warp & 3;
    const int half = warp >> 2;        // which of the quarter's kParts warps: column chunks half, half+kParts, ...
    const int epi_tid = threadIdx.x;
    int acc = 0; uint32_t acc_phase = 0;
    uint32_t zc = 0;                 // TC_DGRAD_ACT: this warp's running chunk count (z ring slot / phase)
    int acc_net = -1;                // TC_DGRAD_ACT / TC_FWD_HEAD: network whose partial sums sit in colacc
    int cbuf = 0;                    // TC_FWD_HEAD: which [2][256] constant buffer holds acc_net's bias / Dense_L kernel
    float h_sl = 0.f, h_w = 0.f, h_sout = 0.f, h_bo = 0.f, h_fls = 0.f, h_flik = 0.f, h_fpi = 0.f, h_fos = 0.f;
    float p_wact = 0.f, p_sprev = 0.f, p_fls = 0.f;   // TC_DGRAD_ACT: per-network constants
    f32x2 gw2 = 0ull, gs2 = 0ull;    // TC_DGRAD_ACT: per-lane packed partial sums of the current network
    float xg_w = 0.f, xg_s = 0.f;    // the same sums in the x3 epilogue (scalar math)
    // this warp's share of the two scalar gradients -> the CTA's shared-memory sums (before a flush)
    auto dact_scalars = [&](int fnet) {
      const float g_w = warp_sum(f2_lo(gw2) + f2_hi(gw2) + xg_w) * a.isf;     // sum dh*diff,  dh = acc*isf
      const float g_s = warp_sum(f2_lo(gs2) + f2_hi(gs2) + xg_s) * a.isf;     // sum dz*z,     dz = dh*act'(z)
      xg_w = 0.f; xg_s = 0.f;
      if (lane == 0) {
        if (a.skip_bias) {
          // no column sums to combine across warps: every warp adds its two scalars straight to the
          // gradient (two atomics per warp and network change) and the CTA-wide barriers go away
          float* g = a.gradp + (size_t)fnet * a.P;
          atomicAdd(g + a.off_actw, g_w * p_wact * (1.f - p_wact));
          atomicAdd(g + a.off_ls_prev, g_s * p_fls);
        } else {
          atomicAdd(&colacc[kAccCols], g_w * p_wact * (1.f - p_wact));
          atomicAdd(&colacc[kAccCols + 1], g_s * p_fls);
        }
      }
      gw2 = 0ull; gs2 = 0ull;
    };
    // TC_FWD_HEAD: per-lane partial sums of the current network, kept in registers across the CTA's
    // tiles and reduced only when the network changes: column sums (lane L <-> column c+L of this
    // warp's chunk i) of dU (bias gradient) and r*h (Dense_L kernel), and the scalar gradients
    constexpr int kHeadChunks = MODE == TC_FWD_HEAD ? (BLOCK_N / 32 + kParts - 1) / kParts : 1;
    // (kEpi16: one slot per 16-column half; lanes 2c and 2c+1 both hold column c of the half)
    constexpr int kHeadSlots = (kEpi16 && MODE == TC_FWD_HEAD) ? 2 * kHeadChunks : kHeadChunks;
    float hcol_b[kHeadSlots], hcol_k[kHeadSlots], hsc[7];
#pragma unroll
    for (int i = 0; i < kHeadSlots; ++i) { hcol_b[i] = 0.f; hcol_k[i] = 0.f; }
#pragma unroll
    for (int i = 0; i < 7; ++i) hsc[i] = 0.f;
    auto head_flush = [&](int fnet) {
      float* g = a.gradp + (size_t)fnet * a.P;
#pragma unroll
      for (int i = 0; i < kHeadSlots; ++i) {
        if constexpr (kEpi16 && MODE == TC_FWD_HEAD) {
          const int c = half * 32 + (i >> 1) * 32 * kParts + 16 * (i & 1) + (lane >> 1);
          if (half * 32 + (i >> 1) * 32 * kParts < BLOCK_N && (lane & 1) == 0) {
            atomicAdd(g + a.off_bias + c, hcol_b[i]);
            atomicAdd(g + dm.off_kernel[dm.L] + c, hcol_k[i] * (h_sout * dm.inv_sqrt_W));
          }
        } else {
          const int c = half * 32 + i * 32 * kParts;
#ifdef BNF_HEAD_ROLLED
          float* hc = hcol_s + ((warp * kHeadChunks + i) * 2) * 32 + lane;     // this lane's own words
          hcol_b[i] = hc[0]; hcol_k[i] = hc[32];
          hc[0] = 0.f; hc[32] = 0.f;
#endif
          if (c < BLOCK_N) {
            atomicAdd(g + a.off_bias + c + lane, hcol_b[i]);
            atomicAdd(g + dm.off_kernel[dm.L] + c + lane, hcol_k[i] * (h_sout * dm.inv_sqrt_W));
          }
        }
        hcol_b[i] = 0.f; hcol_k[i] = 0.f;
      }
      const float s0 = warp_sum(hsc[0]), s1 = warp_sum(hsc[1]);
      if (lane == 0) {
        atomicAdd(g + dm.off_actw, s0 * h_w * (1.f - h_w));
        atomicAdd(g + dm.off_layer_scale[a.layer], s1 * h_fls);
      }
      if (half == 0) {
        const float s2 = warp_sum(hsc[2]), s3 = warp_sum(hsc[3]), s4 = warp_sum(hsc[4]);
        const float s5 = warp_sum(hsc[5]), s6 = warp_sum(hsc[6]);
        if (lane == 0) {
          atomicAdd(a.ll + fnet, s2);
          atomicAdd(g + (dm.likelihood == BNF_NORMAL ? 0 : 1), s3 * h_flik);
          if (dm.likelihood == BNF_ZINB) atomicAdd(g + 2, s4 * h_fpi);
          atomicAdd(g + dm.off_out_scale, s5 * h_fos);
          atomicAdd(g + dm.off_bias[dm.L], s6 * h_sout);
        }
      }
#pragma unroll
      for (int i = 0; i < 7; ++i) hsc[i] = 0.f;
    };
    if (MODE == TC_DGRAD_ENC && epi_tid < 2 * kMaxD + 3) eacc[epi_tid] = 0.f;
    for (int t = tile0; t < tile_end; t += tile_step) {
      const int net = t / tiles_per_net;
      int r = t % tiles_per_net;
      const int n_t = r % a.n_tiles; r /= a.n_tiles;
      const int m_t = CTA2 ? 2 * (r % m_units) + (int)cta_rank : r % m_units;
      const float* dv = a.derived ? a.derived + (size_t)net * kDerivedStride : nullptr;
      float c1 = a.isf, w_act = 0.f;
      float* sb = sbias + acc * 256;
      if constexpr (MODE == TC_FWD_HEAD) {
        // ================= last hidden layer + head + log-likelihood + activation backward =================
        // One tile holds whole rows (n_tiles == 1), so everything downstream of the GEMM happens on
        // the accumulator while it sits in TMEM (models.py:263-273 forward, :157-191 likelihood,
        // and the backward of both).  Pass 1 does all the transcendental math once: h = act(z),
        // the row's h.Ko partial dot, kd = Ko*act'(z) and the row-local parts of the scalar
        // gradients; kd replaces the accumulator in TMEM (tcgen05.st) and h goes, as bf16, to this
        // warp's swizzled shared-memory tiles.  The warps of a lane quarter exchange their partial dots through shared
        // memory; every lane evaluates log p(y|o) and r = dlogp/do of its row.  Pass 2 re-reads
        // kd and h and emits dU = (s_l*r*s_out/sqrt(W))*kd (the only HBM output) plus the bias /
        // Dense_L-kernel column sums.  z and h never leave the SM.  Column and scalar sums
        // collect in shared memory and are flushed when the CTA's network changes (tiles are
        // assigned in contiguous ranges, so that is at most twice per CTA).
        if (net != acc_net) {
          if (acc_net >= 0) head_flush(acc_net);     // this warp's partial sums of the previous network
          acc_net = net;
          cbuf ^= 1;
          const float* pnet = a.params + (size_t)net * a.P;
          h_sl = dv[kDvSLayer + a.layer];
          h_w = dv[kDvActW];
          h_sout = dv[kDvSOut];
          h_bo = pnet[dm.off_bias[dm.L]];
          h_fls = sigmoid_f(pnet[dm.off_layer_scale[a.layer]]) / h_sl;
          h_flik = dm.likelihood == BNF_NORMAL ? expf(pnet[0]) : sigmoid_f(pnet[1]);
          h_fpi = dv[kDvPi] * (1.f - dv[kDvPi]);
          h_fos = sigmoid_f(pnet[dm.off_out_scale]);
          if (epi_tid < BLOCK_N) {
            sbias[cbuf * 256 + epi_tid] = h_sl * pnet[a.off_bias + epi_tid];
            kos_s[cbuf * 256 + epi_tid] = pnet[dm.off_kernel[dm.L] + epi_tid];
          }
          asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpi) : "memory");       // constants staged
        }
        const float* hsb = sbias + cbuf * 256;
        const float* kos = kos_s + cbuf * 256;
        const float s_l = h_sl, w = h_w, s_out = h_sout, bo = h_bo;
        const float cz = s_l * a.isf;
        const float hc = s_out * dm.inv_sqrt_W;
        const int row = m_t * 128 + q * 32 + lane;
        const bool row_ok = row < a.m_valid;
        float yv = 0.f;
        if (row_ok) yv = a.y[a.idx ? (size_t)a.idx[(size_t)net * a.idx_stride + row] : (size_t)row];
        if (epi_tid == 0) TL((t - tile0) / tile_step, 6);
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        if (epi_tid == 0) TL((t - tile0) / tile_step, 7);
        const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N);   // kd replaces the accumulator
        if constexpr (X3) {
          // ---- bf16x3 (f32-parity) variant: nothing is stashed -- pass 1 evaluates h = act(z) only for
          // the row dots, pass 2 re-reads the untouched accumulator and RECOMPUTES the activation
          // (one ex2 + one rcp per element and pass; the tile's six-segment MMAs take as long as
          // both passes), then emits dU as three bf16 planes.  z, h of the layer never exist in HBM.
          float dot = 0.f;
#pragma unroll 1
          for (int c = half * 32; c < BLOCK_N; c += 32 * kParts) {
            uint32_t v[32];
            tmem_ld32(tacc + (uint32_t)c, v);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(hsb + c + j);
              const float4 k4 = *reinterpret_cast<const float4*>(kos + c + j);
              const float bb[4] = {b4.x, b4.y, b4.z, b4.w}, kk[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float d, hh;
                act_grad_x3(fmaf(__uint_as_float(v[j + e]), cz, bb[e]), w, &d, &hh);
                dot = fmaf(hh, kk[e], dot);
              }
            }
          }
          rowdot[half * 128 + q * 32 + lane] = dot;
          asm volatile("bar.sync %0, %1;" ::"r"(2 + q), "n"(32 * kParts) : "memory");   // the warps of this lane quarter
          float dsum = rowdot[q * 32 + lane];
#pragma unroll
          for (int pp = 1; pp < kParts; ++pp) dsum += rowdot[pp * 128 + q * 32 + lane];
          const float opre = dsum * dm.inv_sqrt_W + bo;
          float gl3[3] = {0.f, 0.f, 0.f};
          float rr = 0.f, logp = 0.f;
          if (row_ok) logp = head_row_loglik(dm.likelihood, dv, s_out * opre, yv, &rr, gl3);
          if (!row_ok) rr = 0.f;
          const float rk = rr * hc;                     // dh[col] = rk * Ko[col]
          const float rks = rk * s_l;
          float gw = 0.f, gs = 0.f;
#ifdef BNF_HEAD_ROLLED
#pragma unroll 1
#else
#pragma unroll
#endif
          for (int ci = 0; ci < kHeadChunks; ++ci) {
            const int c = half * 32 + ci * 32 * kParts;
            if (c >= BLOCK_N) break;
            uint32_t v[32], pk[kDuPlanes][16];
            float du[32], gk[32];
            tmem_ld32(tacc + (uint32_t)c, v);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(hsb + c + j);
              const float4 k4 = *reinterpret_cast<const float4*>(kos + c + j);
              const float bb[4] = {b4.x, b4.y, b4.z, b4.w}, kk[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float zz = fmaf(__uint_as_float(v[j + e]), cz, bb[e]);
                float d, hh;
                const float da = act_grad_x3(zz, w, &d, &hh);
                const float kd = kk[e] * da;
                gw = fmaf(kk[e], d, gw);              // sum Ko*diff   (x rk  = sum dh*diff)
                gs = fmaf(kd, zz, gs);                // sum kd*z      (x rk  = sum dz*z)
                du[j + e] = kd * rks;
                gk[j + e] = hh * rr;
              }
              split2_pair(du[j], du[j + 1], &pk[0][j / 2], &pk[1][j / 2]);
              split2_pair(du[j + 2], du[j + 3], &pk[0][j / 2 + 1], &pk[1][j / 2 + 1]);
            }
            uint8_t* stg = staging + warp * Cfg::kStgWarp;       // dU planes 2 x 2 KB
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
#pragma unroll
            for (int pl = 0; pl < kDuPlanes; ++pl)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                *reinterpret_cast<uint4*>(stg + pl * 2048 + lane * 64 + ((k ^ ((lane >> 1) & 3)) << 4)) =
                    make_uint4(pk[pl][4 * k], pk[pl][4 * k + 1], pk[pl][4 * k + 2], pk[pl][4 * k + 3]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
#pragma unroll
              for (int pl = 0; pl < kDuPlanes; ++pl) tma_store_3d(&map_o0, stg + pl * 2048, c + pl * BLOCK_N, m_t * 128 + q * 32, net);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            warp_transpose_sum(du, lane);
            warp_transpose_sum(gk, lane);
#ifdef BNF_HEAD_ROLLED
            float* hc = hcol_s + ((warp * kHeadChunks + ci) * 2) * 32 + lane;
            hc[0] += du[0];
            hc[32] += gk[0];
#else
            hcol_b[ci] += du[0];
            hcol_k[ci] += gk[0];
#endif
          }
          tc_fence_before();
          if (CTA2) mbar_arrive_remote(&tempty[acc], 0);
          else mbar_arrive(&tempty[acc]);
          if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
          hsc[0] += rk * gw;
          hsc[1] += rk * gs;
          if (half == 0) {
            hsc[2] += logp;
            hsc[3] += dm.likelihood == BNF_NORMAL ? gl3[0] : gl3[1];
            hsc[4] += gl3[2];
            hsc[5] += rr * opre;
            hsc[6] += rr;
          }
          continue;
        }
        constexpr int kMaxChunks = (BLOCK_N / 32 + kParts - 1) / kParts;   // chunks per warp and tile
        uint8_t* hw_scr = hscr + (size_t)warp * kMaxChunks * 2048;        // this warp's h tiles
        // ---- pass 1 (packed f32x2 math: FFMA2)
        const ActConst2 ak(w);
        const f32x2 cz2 = f2_dup(cz);
        f32x2 dot2 = 0ull, gw2 = 0ull, gs2 = 0ull;
#pragma unroll 1
        for (int c = half * 32; c < BLOCK_N; c += 32 * kParts) {
          if constexpr (kEpi16) {
            uint8_t* ht = hw_scr + ((c - half * 32) / (32 * kParts)) * 2048 + lane * 64;
#pragma unroll 1
            for (int hh = 0; hh < 2; ++hh) {
              const int ch = c + 16 * hh;
              uint32_t v[16], hv[8];
              tmem_ld16(tacc + (uint32_t)ch, v);
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                const float4 b4 = *reinterpret_cast<const float4*>(hsb + ch + j);
                const float4 k4 = *reinterpret_cast<const float4*>(kos + ch + j);
#pragma unroll
                for (int jj = 0; jj < 4; jj += 2) {
                  const f32x2 b2 = jj ? f2_pack(b4.z, b4.w) : f2_pack(b4.x, b4.y);
                  const f32x2 k2 = jj ? f2_pack(k4.z, k4.w) : f2_pack(k4.x, k4.y);
                  const f32x2 z2 = f2_fma(f2_pack(v[j + jj], v[j + jj + 1]), cz2, b2);
                  f32x2 d2, h2;
                  const f32x2 da2 = act_grad_fast2(z2, ak, &d2, &h2);
                  const f32x2 kd2 = f2_mul(k2, da2);
                  dot2 = f2_fma(h2, k2, dot2);
                  gw2 = f2_fma(k2, d2, gw2);
                  gs2 = f2_fma(kd2, z2, gs2);
                  v[j + jj] = __float_as_uint(f2_lo(kd2)); v[j + jj + 1] = __float_as_uint(f2_hi(kd2));
                  hv[(j + jj) >> 1] = f2_to_bf16x2(h2);
                }
              }
              tmem_st16(tacc + (uint32_t)ch, v);
#pragma unroll
              for (int k2 = 0; k2 < 2; ++k2)
                *reinterpret_cast<uint4*>(ht + (((2 * hh + k2) ^ ((lane >> 1) & 3)) << 4)) =
                    make_uint4(hv[4 * k2], hv[4 * k2 + 1], hv[4 * k2 + 2], hv[4 * k2 + 3]);
            }
            continue;
          }
          uint32_t v[32], hv[16];
          tmem_ld32(tacc + (uint32_t)c, v);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(hsb + c + j);
            const float4 k4 = *reinterpret_cast<const float4*>(kos + c + j);
#pragma unroll
            for (int jj = 0; jj < 4; jj += 2) {
              const f32x2 b2 = jj ? f2_pack(b4.z, b4.w) : f2_pack(b4.x, b4.y);
              const f32x2 k2 = jj ? f2_pack(k4.z, k4.w) : f2_pack(k4.x, k4.y);
              const f32x2 z2 = f2_fma(f2_pack(v[j + jj], v[j + jj + 1]), cz2, b2);
              f32x2 d2, h2;
              const f32x2 da2 = act_grad_fast2(z2, ak, &d2, &h2);
              const f32x2 kd2 = f2_mul(k2, da2);
              dot2 = f2_fma(h2, k2, dot2);
              gw2 = f2_fma(k2, d2, gw2);              // sum Ko*diff   (x rk  = sum dh*diff)
              gs2 = f2_fma(kd2, z2, gs2);             // sum kd*z      (x rk  = sum dz*z)
              v[j + jj] = __float_as_uint(f2_lo(kd2)); v[j + jj + 1] = __float_as_uint(f2_hi(kd2));
              hv[(j + jj) >> 1] = f2_to_bf16x2(h2);
            }
          }
          tmem_st32(tacc + (uint32_t)c, v);
          uint8_t* ht = hw_scr + ((c - half * 32) / (32 * kParts)) * 2048 + lane * 64;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            *reinterpret_cast<uint4*>(ht + ((k ^ ((lane >> 1) & 3)) << 4)) =
                make_uint4(hv[4 * k], hv[4 * k + 1], hv[4 * k + 2], hv[4 * k + 3]);
        }
        tmem_st_wait();
        const float dot = f2_lo(dot2) + f2_hi(dot2);
        if (epi_tid == 0) TL((t - tile0) / tile_step, 8);
        rowdot[half * 128 + q * 32 + lane] = dot;
        asm volatile("bar.sync %0, %1;" ::"r"(2 + q), "n"(32 * kParts) : "memory");   // the warps of this lane quarter
        float dsum = rowdot[q * 32 + lane];
#pragma unroll
        for (int pp = 1; pp < kParts; ++pp) dsum += rowdot[pp * 128 + q * 32 + lane];
        const float opre = dsum * dm.inv_sqrt_W + bo;
        // ---- likelihood of the row (every warp of the quarter evaluates it; part 0 owns the sums)
        float gl3[3] = {0.f, 0.f, 0.f};
        float rr = 0.f, logp = 0.f;
        if (row_ok) logp = head_row_loglik(dm.likelihood, dv, s_out * opre, yv, &rr, gl3);
        if (!row_ok) rr = 0.f;
        const float rk = rr * hc;                       // dh[col] = rk * Ko[col]
        if (epi_tid == 0) TL((t - tile0) / tile_step, 9);
        // ---- pass 2: dU = kd*(rk*s_l), column sums of dU (bias gradient) and of r*h (Dense_L kernel)
        const f32x2 rr2 = f2_dup(rr);
        const f32x2 rks2 = f2_dup(rk * s_l);
#ifdef BNF_HEAD_ROLLED
#pragma unroll 1
#else
#pragma unroll
#endif
        for (int ci = 0; ci < kHeadChunks; ++ci) {
          const int c = half * 32 + ci * 32 * kParts;
          if (c >= BLOCK_N) break;
          if constexpr (kEpi16) {
            const uint8_t* ht = hw_scr + ci * 2048 + lane * 64;
            uint8_t* stg = staging + warp * Cfg::kStgWarp;
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              uint32_t v[16], hv[8], pk[8];
              float du[16], gk[16];
              tmem_ld16(tacc + (uint32_t)(c + 16 * hh), v);
#pragma unroll
              for (int k2 = 0; k2 < 2; ++k2) {
                const uint4 q4 = *reinterpret_cast<const uint4*>(ht + (((2 * hh + k2) ^ ((lane >> 1) & 3)) << 4));
                hv[4 * k2] = q4.x; hv[4 * k2 + 1] = q4.y; hv[4 * k2 + 2] = q4.z; hv[4 * k2 + 3] = q4.w;
              }
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                const f32x2 du2 = f2_mul(f2_pack(v[j], v[j + 1]), rks2);
                const uint32_t hb = hv[j >> 1];
                const f32x2 gk2 = f2_mul(f2_pack(hb << 16, hb & 0xffff0000u), rr2);      // r * h
                du[j] = f2_lo(du2); du[j + 1] = f2_hi(du2);
                gk[j] = f2_lo(gk2); gk[j + 1] = f2_hi(gk2);
                pk[j >> 1] = f2_to_bf16x2(du2);
              }
#pragma unroll
              for (int k2 = 0; k2 < 2; ++k2)
                *reinterpret_cast<uint4*>(stg + lane * 64 + (((2 * hh + k2) ^ ((lane >> 1) & 3)) << 4)) =
                    make_uint4(pk[4 * k2], pk[4 * k2 + 1], pk[4 * k2 + 2], pk[4 * k2 + 3]);
              warp_transpose_sum16(du, lane);
              warp_transpose_sum16(gk, lane);
              hcol_b[2 * ci + hh] += du[0];
              hcol_k[2 * ci + hh] += gk[0];
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(&map_o0, stg, c, m_t * 128 + q * 32, net);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            continue;
          }
          uint32_t v[32], hv[16];
          tmem_ld32(tacc + (uint32_t)c, v);
          const uint8_t* ht = hw_scr + ci * 2048 + lane * 64;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint4 q4 = *reinterpret_cast<const uint4*>(ht + ((k ^ ((lane >> 1) & 3)) << 4));
            hv[4 * k] = q4.x; hv[4 * k + 1] = q4.y; hv[4 * k + 2] = q4.z; hv[4 * k + 3] = q4.w;
          }
          uint32_t pk[16];
          float du[32], gk[32];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const f32x2 du2 = f2_mul(f2_pack(v[j], v[j + 1]), rks2);
            const uint32_t hb = hv[j >> 1];
            const f32x2 gk2 = f2_mul(f2_pack(hb << 16, hb & 0xffff0000u), rr2);      // r * h
            du[j] = f2_lo(du2); du[j + 1] = f2_hi(du2);
            gk[j] = f2_lo(gk2); gk[j + 1] = f2_hi(gk2);
            pk[j >> 1] = f2_to_bf16x2(du2);
          }
          uint8_t* stg = staging + warp * Cfg::kStgWarp;
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
#pragma unroll
          for (int k = 0; k < 4; ++k)
            *reinterpret_cast<uint4*>(stg + lane * 64 + ((k ^ ((lane >> 1) & 3)) << 4)) =
                make_uint4(pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&map_o0, stg, c, m_t * 128 + q * 32, net);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          // column sums over this warp's 32 rows by transpose-reduce (lane L ends with column L)
          warp_transpose_sum(du, lane);
          warp_transpose_sum(gk, lane);
#ifdef BNF_HEAD_ROLLED
          float* hc = hcol_s + ((warp * kHeadChunks + ci) * 2) * 32 + lane;
          hc[0] += du[0];
          hc[32] += gk[0];
#else
          hcol_b[ci] += du[0];
          hcol_k[ci] += gk[0];
#endif
        }
        if (epi_tid == 0) TL((t - tile0) / tile_step, 10);
        tc_fence_before();
        if (CTA2) mbar_arrive_remote(&tempty[acc], 0);
        else mbar_arrive(&tempty[acc]);
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
        // ---- this row's share of the scalar gradients (reduced over lanes at the flush)
        hsc[0] += rk * (f2_lo(gw2) + f2_hi(gw2));       // sum dh*diff,  dh = rk*Ko
        hsc[1] += rk * (f2_lo(gs2) + f2_hi(gs2));       // sum dz*z,     dz = dh*act'(z)
        if (half == 0) {
          hsc[2] += logp;
          hsc[3] += dm.likelihood == BNF_NORMAL ? gl3[0] : gl3[1];
          hsc[4] += gl3[2];
          hsc[5] += rr * opre;
          hsc[6] += rr;
        }
        continue;
      }
      if (MODE == TC_FWD) {
        // per-(network, n-tile) constants: reloaded and restaged only when the key changes (with
        // one n-tile the CTA's tiles are a contiguous range, i.e. mostly one network)
        const int key = net * a.n_tiles + n_t;
        if (key != acc_net) {
          acc_net = key;
          cbuf ^= 1;
          const float s_l = dv[kDvSLayer + a.layer];
          p_sprev = s_l * a.isf;
          p_wact = dv[kDvActW];
          if (epi_tid < BLOCK_N)
            sbias[cbuf * 256 + epi_tid] = s_l * a.params[(size_t)net * a.P + a.off_bias + n_t * BLOCK_N + epi_tid];
          asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpi) : "memory");
        }
        c1 = p_sprev;
        w_act = p_wact;
        sb = sbias + cbuf * 256;
      }
      const int row = m_t * 128 + q * 32 + lane;
      const bool row_ok = row < a.m_valid;
      // TC_DGRAD_ACT: the z tile (bf16, previous layer's forward epilogue) of every 32x32 chunk
      // arrives by TMA in this warp's ring of kZRing 64B-swizzled 2 KB slots, kZRing chunks ahead;
      // each lane then reads its own row with conflict-free 16-byte LDS.  The epilogue issues NO
      // global loads or atomics inside the chunk loop: the fence.proxy.async in front of every
      // TMA store is a MEMBAR.ALL.CTA, which waits for all of the thread's outstanding LSU
      // traffic -- with per-lane z loads and per-chunk atomics in flight every chunk paid a full
      // DRAM round trip (measured: 15.1 -> 10.1 ms on the wind shard with both removed).
      float s_prev = 0.f;
      f32x2 cdu2 = 0ull;
      constexpr int kZSlot = X3 ? 4096 : 2048;      // x3: z is f32 (32x32 f32 tile, 128B swizzle)
      uint8_t* zring = staging + warp * Cfg::kStgWarp + (X3 ? 0 : 2048);
      uint64_t* zb = zbar + warp * kZRing;
      if (MODE == TC_DGRAD_ACT) {
        if (net != acc_net) {
          // the network changed: flush the CTA's partial sums of the previous one
          if (acc_net >= 0) {
            dact_scalars(acc_net);
            if (!a.skip_bias) {
              asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpi) : "memory");     // all adds are in
              float* g = a.gradp + (size_t)acc_net * a.P;
              for (int i = epi_tid; i < a.n_valid; i += 32 * kEpi) {
                atomicAdd(g + a.off_bias_prev + i, colacc[i]);
                colacc[i] = 0.f;
              }
              if (epi_tid < 2) {
                atomicAdd(g + (epi_tid == 0 ? a.off_actw : a.off_ls_prev), colacc[kAccCols + epi_tid]);
                colacc[kAccCols + epi_tid] = 0.f;
              }
              asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpi) : "memory");     // zeroed before new adds
            }
          }
          acc_net = net;
          p_wact = dv[kDvActW];
          p_sprev = dv[kDvSLayer + a.layer_prev];
          p_fls = sigmoid_f(a.params[(size_t)net * a.P + a.off_ls_prev]) / p_sprev;
        }
        w_act = p_wact;
        s_prev = p_sprev;
        cdu2 = f2_dup(a.isf * s_prev);
        const uint32_t zc_u = __shfl_sync(0xffffffffu, zc, 0);   // (warp-uniform for the compiler: see `warp`)
        if (lane == 0 && !DBG(1)) {
#pragma unroll
          for (int i = 0; i < kZRing; ++i) {
            if (half * 32 + 32 * kParts * i >= BLOCK_N) break;
            const uint32_t sl = (zc_u + i) % kZRing;
            mbar_arrive_expect_tx(&zb[sl], kZSlot);
            tma_load_3d(zring + sl * kZSlot, &map_o1, &zb[sl], n_t * BLOCK_N + half * 32 + 32 * kParts * i, m_t * 128 + q * 32, net);
          }
        }
      }
      // TC_DGRAD_ENC: a tile's input rows -> shared memory (scaled by 1/(input_scale*exp(lsa));
      // slot D keeps the raw time for the seasonal units), double-buffered: tile i+1's rows are
      // fetched while tile i's encode backward runs
      auto stage_x = [&](int tt, float* dst) {
        const int net_x = tt / tiles_per_net;
        const int m_x = (tt % tiles_per_net) / (a.n_tiles * a.k_splits);
        const float* dvx = a.derived + (size_t)net_x * kDerivedStride;
        for (int e = epi_tid; e < 128 * dm.D; e += 32 * kEpi) {
          const int r = e / dm.D, i = e - r * dm.D;
          const int b = min(m_x * 128 + r, a.m_valid - 1);
          const float xv = a.x[(a.idx ? (size_t)a.idx[(size_t)net_x * a.idx_stride + b] : (size_t)b) * dm.D + i];
          dst[r * (kMaxD + 1) + i] = xv / dvx[kDvDenom + i];
          if (i == 0) dst[r * (kMaxD + 1) + dm.D] = xv;
        }
      };
      float* sx_cur = sxt + ((t - tile0) & 1) * 128 * (kMaxD + 1);
      float* sx_nxt = sxt + (((t - tile0) & 1) ^ 1) * 128 * (kMaxD + 1);
      if (MODE == TC_DGRAD_ENC && t == tile0) stage_x(t, sx_cur);
      if (epi_tid == 0) TL((t - tile0) / tile_step, 6);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      if (epi_tid == 0) TL((t - tile0) / tile_step, 7);
      if constexpr (X3 && (MODE == TC_FWD || MODE == TC_DGRAD_ACT)) {
        // ---- bf16x3 epilogues
        auto chunk_x3 = [&](int c, uint32_t* v) {
          const int col0 = n_t * BLOCK_N + c;
          if constexpr (MODE == TC_FWD) {
            // ---- bf16x3 forward: z = acc*c1 + s_l*b stays f32 (the backward pass needs it), h = act(z)
            // leaves as its three bf16 planes (the next GEMM's split A operand).  Staging per warp:
            // [z 32x32 f32, 128B swizzle | h planes 3 x (32x32 bf16, 64B swizzle)]
            uint8_t* stg = staging + warp * Cfg::kStgWarp;
            uint32_t hp[3][16];
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
  #pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float4 b4 = *reinterpret_cast<const float4*>(sb + c + 4 * k);
              const float z0 = fmaf(__uint_as_float(v[4 * k]), c1, b4.x), z1 = fmaf(__uint_as_float(v[4 * k + 1]), c1, b4.y);
              const float z2 = fmaf(__uint_as_float(v[4 * k + 2]), c1, b4.z), z3 = fmaf(__uint_as_float(v[4 * k + 3]), c1, b4.w);
              if (a.out0) *reinterpret_cast<float4*>(stg + lane * 128 + ((k ^ (lane & 7)) << 4)) = make_float4(z0, z1, z2, z3);
              float d, h0, h1, h2, h3;
              act_grad_x3(z0, w_act, &d, &h0);
              act_grad_x3(z1, w_act, &d, &h1);
              act_grad_x3(z2, w_act, &d, &h2);
              act_grad_x3(z3, w_act, &d, &h3);
              split3_pair(h0, h1, &hp[0][2 * k], &hp[1][2 * k], &hp[2][2 * k]);
              split3_pair(h2, h3, &hp[0][2 * k + 1], &hp[1][2 * k + 1], &hp[2][2 * k + 1]);
            }
  #pragma unroll
            for (int p = 0; p < 3; ++p)
  #pragma unroll
              for (int k = 0; k < 4; ++k)
                *reinterpret_cast<uint4*>(stg + 4096 + p * 2048 + lane * 64 + ((k ^ ((lane >> 1) & 3)) << 4)) =
                    make_uint4(hp[p][4 * k], hp[p][4 * k + 1], hp[p][4 * k + 2], hp[p][4 * k + 3]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
  #pragma unroll
              for (int p = 0; p < 3; ++p) tma_store_3d(&map_o1, stg + 4096 + p * 2048, col0 + p * a.n_valid, m_t * 128 + q * 32, net);
              if (a.out0) tma_store_3d(&map_o0, stg, col0, m_t * 128 + q * 32, net);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          } else {
            // ---- bf16x3 dgrad + activation backward: the f32 z tile of the chunk arrives by TMA
            // (ring of kZRing 4 KB slots, 128B swizzle); dU leaves as two bf16 planes.
            // Staging per warp: [z ring | dU planes 2 x 2 KB]
            uint8_t* so = staging + warp * Cfg::kStgWarp + kZRing * 4096;
            const uint32_t zsl = __shfl_sync(0xffffffffu, zc, 0) % kZRing;
            mbar_wait(&zb[zsl], (zc / kZRing) & 1u);
            const uint8_t* zt = zring + zsl * 4096 + lane * 128;
            uint32_t pk[kDuPlanes][16];
            float du[32];
            const float cdu = a.isf * s_prev;
  #pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float4 z4 = *reinterpret_cast<const float4*>(zt + ((k ^ (lane & 7)) << 4));
              const float zz[4] = {z4.x, z4.y, z4.z, z4.w};
  #pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float vv = __uint_as_float(v[4 * k + e]);
                float diff, hh;
                const float da = act_grad_x3(zz[e], w_act, &diff, &hh);
                const float x = vv * da;
                xg_w = fmaf(vv, diff, xg_w);
                xg_s = fmaf(x, zz[e], xg_s);
                du[4 * k + e] = x * cdu;
              }
              split2_pair(du[4 * k], du[4 * k + 1], &pk[0][2 * k], &pk[1][2 * k]);
              split2_pair(du[4 * k + 2], du[4 * k + 3], &pk[0][2 * k + 1], &pk[1][2 * k + 1]);
            }
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
            // every lane has read its z row: refill the slot with the tile kZRing chunks ahead
            if (lane == 0 && c + 32 * kParts * kZRing < BLOCK_N) {
              mbar_arrive_expect_tx(&zb[zsl], 4096);
              tma_load_3d(zring + zsl * 4096, &map_o1, &zb[zsl], col0 + 32 * kParts * kZRing, m_t * 128 + q * 32, net);
            }
            ++zc;
  #pragma unroll
            for (int p = 0; p < 3; ++p)
  #pragma unroll
              for (int k = 0; k < 4; ++k)
                *reinterpret_cast<uint4*>(so + p * 2048 + lane * 64 + ((k ^ ((lane >> 1) & 3)) << 4)) =
                    make_uint4(pk[p][4 * k], pk[p][4 * k + 1], pk[p][4 * k + 2], pk[p][4 * k + 3]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
  #pragma unroll
              for (int p = 0; p < kDuPlanes; ++p) tma_store_3d(&map_o0, so + p * 2048, col0 + p * a.n_valid, m_t * 128 + q * 32, net);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            if (!a.skip_bias) {                       // else the Dense_0 wgrad GEMM delivers these sums
              warp_transpose_sum(du, lane);
              atomicAdd(&colacc[col0 + lane], du[0]);
            }
          }
        };
        const uint32_t tb = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N);
#pragma unroll 1
        for (int c = half * 32; c < BLOCK_N; c += 32 * kParts) {
          // (prefetching the next chunk's accumulator into a second register set while this one is
          // processed was measured -- r2i -- and gains nothing: the exposed tcgen05.ld latency just
          // moves to the next dependent instruction, and the extra 32 registers spill)
          uint32_t v[32];
          tmem_ld32(tb + (uint32_t)c, v);
          chunk_x3(c, v);
        }
      } else
#pragma unroll 1
      for (int c = half * 32; c < BLOCK_N; c += 32 * kParts) {
        if constexpr (kEpi16 && A_MODE != 2 && (MODE == TC_FWD || MODE == TC_DGRAD_ACT)) {
          // ---- sixteen-warp variant: the 32x32 chunk as two 16-column halves (same staging tile,
          // same TMA store / z ring; half the live registers)
          const int col0 = n_t * BLOCK_N + c;
          const uint32_t tchunk = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + c);
          uint8_t* stg = staging + warp * Cfg::kStgWarp;
          const ActConst2 ak(w_act);
          if constexpr (MODE == TC_FWD) {
            const f32x2 c12 = f2_dup(c1);
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
#pragma unroll 1
            for (int hh = 0; hh < 2; ++hh) {
              uint32_t v[16], zp[8], hp[8];
              tmem_ld16(tchunk + (uint32_t)(16 * hh), v);
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                const float4 b4 = *reinterpret_cast<const float4*>(sb + c + 16 * hh + j);
#pragma unroll
                for (int jj = 0; jj < 4; jj += 2) {
                  const f32x2 z2 = f2_fma(f2_pack(v[j + jj], v[j + jj + 1]), c12, jj ? f2_pack(b4.z, b4.w) : f2_pack(b4.x, b4.y));
                  zp[(j + jj) >> 1] = f2_to_bf16x2(z2);
                  hp[(j + jj) >> 1] = f2_to_bf16x2(act_fast2(z2, ak));
                }
              }
#pragma unroll
              for (int k2 = 0; k2 < 2; ++k2) {
                const int off = lane * 64 + (((2 * hh + k2) ^ ((lane >> 1) & 3)) << 4);
                *reinterpret_cast<uint4*>(stg + off) = make_uint4(hp[4 * k2], hp[4 * k2 + 1], hp[4 * k2 + 2], hp[4 * k2 + 3]);
                if (a.out0)
                  *reinterpret_cast<uint4*>(stg + 2048 + off) = make_uint4(zp[4 * k2], zp[4 * k2 + 1], zp[4 * k2 + 2], zp[4 * k2 + 3]);
              }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(&map_o1, stg, col0, m_t * 128 + q * 32, net);
              if (a.out0) tma_store_3d(&map_o0, stg + 2048, col0, m_t * 128 + q * 32, net);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          } else {
            const uint32_t zsl = __shfl_sync(0xffffffffu, zc, 0) % kZRing;
            mbar_wait(&zb[zsl], (zc / kZRing) & 1u);
            const uint8_t* zt = zring + zsl * 2048 + lane * 64;
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
#pragma unroll 1
            for (int hh = 0; hh < 2; ++hh) {
              uint32_t v[16], zw[8], pk[8];
              float du[16];
              tmem_ld16(tchunk + (uint32_t)(16 * hh), v);
#pragma unroll
              for (int k2 = 0; k2 < 2; ++k2) {
                const uint4 q4 = *reinterpret_cast<const uint4*>(zt + (((2 * hh + k2) ^ ((lane >> 1) & 3)) << 4));
                zw[4 * k2] = q4.x; zw[4 * k2 + 1] = q4.y; zw[4 * k2 + 2] = q4.z; zw[4 * k2 + 3] = q4.w;
              }
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                const uint32_t zb = zw[j >> 1];
                const f32x2 z2 = f2_pack(zb << 16, zb & 0xffff0000u);
                const f32x2 v2 = f2_pack(v[j], v[j + 1]);
                f32x2 d2;
                const f32x2 da2 = act_grad_fast2(z2, ak, &d2);
                const f32x2 x2 = f2_mul(v2, da2);
                gw2 = f2_fma(v2, d2, gw2);
                gs2 = f2_fma(x2, z2, gs2);
                const f32x2 du2 = f2_mul(x2, cdu2);
                du[j] = f2_lo(du2);
                du[j + 1] = f2_hi(du2);
                pk[j >> 1] = f2_to_bf16x2(du2);
              }
#pragma unroll
              for (int k2 = 0; k2 < 2; ++k2)
                *reinterpret_cast<uint4*>(stg + lane * 64 + (((2 * hh + k2) ^ ((lane >> 1) & 3)) << 4)) =
                    make_uint4(pk[4 * k2], pk[4 * k2 + 1], pk[4 * k2 + 2], pk[4 * k2 + 3]);
              warp_transpose_sum16(du, lane);
              if ((lane & 1) == 0) atomicAdd(&colacc[col0 + 16 * hh + (lane >> 1)], du[0]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(&map_o0, stg, col0, m_t * 128 + q * 32, net);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
              // every lane has read its z values (syncwarp above): refill the slot kZRing ahead
              if (c + 32 * kParts * kZRing < BLOCK_N) {
                mbar_arrive_expect_tx(&zb[zsl], 2048);
                tma_load_3d(zring + zsl * 2048, &map_o1, &zb[zsl], col0 + 32 * kParts * kZRing, m_t * 128 + q * 32, net);
              }
            }
            ++zc;
          }
          continue;
        }
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + c), v);
        const int col0 = n_t * BLOCK_N + c;
        if (epi_tid == 0 && c == 0) TL((t - tile0) / tile_step, 11);       // first chunk: accumulator in registers
        if (MODE == TC_FWD) {
          uint32_t zp[16], hp[16];
          const ActConst2 ak(w_act);
          const f32x2 c12 = f2_dup(c1);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(sb + c + j);
#pragma unroll
            for (int jj = 0; jj < 4; jj += 2) {
              // packed f32x2 math (FFMA2): z = acc*c1 + s_l*b, h = act(z)
              const f32x2 z2 = f2_fma(f2_pack(v[j + jj], v[j + jj + 1]), c12, jj ? f2_pack(b4.z, b4.w) : f2_pack(b4.x, b4.y));
              zp[(j + jj) >> 1] = f2_to_bf16x2(z2);
              hp[(j + jj) >> 1] = f2_to_bf16x2(DBG(8) ? z2 : act_fast2(z2, ak));
            }
          }
          // registers -> 64B-swizzled smem tile (conflict-free 16B stores) -> TMA store:
          // full 64-byte rows leave the SM as bulk writes instead of 32 scattered
          // 16-byte stores per instruction; rows >= B are clipped by the tensor map.
          uint8_t* stg = staging + warp * Cfg::kStgWarp;
          if (epi_tid == 0 && c == 0) TL((t - tile0) / tile_step, 12);     // math done
          if (lane == 0 && !DBG(4)) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
          if (epi_tid == 0 && c == 0) TL((t - tile0) / tile_step, 13);     // staging tile free
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int off = lane * 64 + ((k ^ ((lane >> 1) & 3)) << 4);
            *reinterpret_cast<uint4*>(stg + off) = make_uint4(hp[4 * k], hp[4 * k + 1], hp[4 * k + 2], hp[4 * k + 3]);
            if (a.out0)
              *reinterpret_cast<uint4*>(stg + 2048 + off) = make_uint4(zp[4 * k], zp[4 * k + 1], zp[4 * k + 2], zp[4 * k + 3]);
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0 && !DBG(4)) {
            tma_store_3d(&map_o1, stg, col0, m_t * 128 + q * 32, net);
            if (a.out0) tma_store_3d(&map_o0, stg + 2048, col0, m_t * 128 + q * 32, net);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          if (epi_tid == 0 && c == 0) TL((t - tile0) / tile_step, 14);     // stores issued
        } else if (MODE == TC_DGRAD_ACT) {
          // dh = acc/sqrt(fan_in); dz = dh*act'(z); dU = s*dz; plus the reductions that the
          // separate act_bwd kernel used to do (bias column sums, activation-mix and
          // layer-scale scalars).  models.py:255-268 backward.
          uint32_t zw[16];
          const uint32_t zsl = __shfl_sync(0xffffffffu, zc, 0) % kZRing;
          if (!DBG(1)) {
            mbar_wait(&zb[zsl], (zc / kZRing) & 1u);
            const uint8_t* zt = zring + zsl * 2048 + lane * 64;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint4 q4 = *reinterpret_cast<const uint4*>(zt + ((k ^ ((lane >> 1) & 3)) << 4));
              zw[4 * k] = q4.x; zw[4 * k + 1] = q4.y; zw[4 * k + 2] = q4.z; zw[4 * k + 3] = q4.w;
            }
          } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) zw[k] = 0u;
          }
          // packed f32x2 math on adjacent column pairs (FFMA2): x = acc*act'(z), dU = x*(isf*s_prev);
          // the scalar sums are kept unscaled (sum acc*diff, sum x*z) and scaled once per tile
          uint32_t pk[16];
          float du[32];
          const ActConst2 ak(w_act);
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const uint32_t zb = zw[j >> 1];
            const f32x2 z2 = f2_pack(zb << 16, zb & 0xffff0000u);
            const f32x2 v2 = f2_pack(v[j], v[j + 1]);
            f32x2 d2;
            f32x2 da2 = act_grad_fast2(z2, ak, &d2);
            if (DBG(8)) { da2 = z2; d2 = z2; }
            const f32x2 x2 = f2_mul(v2, da2);
            gw2 = f2_fma(v2, d2, gw2);
            gs2 = f2_fma(x2, z2, gs2);
            const f32x2 du2 = f2_mul(x2, cdu2);
            du[j] = f2_lo(du2);
            du[j + 1] = f2_hi(du2);
            pk[j >> 1] = f2_to_bf16x2(du2);
          }
          uint8_t* stg = staging + warp * Cfg::kStgWarp;
          if (epi_tid == 0 && c == 0) TL((t - tile0) / tile_step, 12);     // math done
          if (lane == 0 && !DBG(4)) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
          if (epi_tid == 0 && c == 0) TL((t - tile0) / tile_step, 13);     // staging tile free
          // every lane holds its z values in registers: refill the slot with the tile kZRing ahead
          if (lane == 0 && c + 32 * kParts * kZRing < BLOCK_N && !DBG(1)) {
            mbar_arrive_expect_tx(&zb[zsl], 2048);
            tma_load_3d(zring + zsl * 2048, &map_o1, &zb[zsl], col0 + 32 * kParts * kZRing, m_t * 128 + q * 32, net);
          }
          ++zc;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            *reinterpret_cast<uint4*>(stg + lane * 64 + ((k ^ ((lane >> 1) & 3)) << 4)) =
                make_uint4(pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0 && !DBG(4)) {
            tma_store_3d(&map_o0, stg, col0, m_t * 128 + q * 32, net);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          if (epi_tid == 0 && c == 0) TL((t - tile0) / tile_step, 14);     // store issued
          if (DBG(2) || a.skip_bias) continue;      // skip_bias: the Dense_0 wgrad GEMM delivers these sums
          // bias gradient: column sums over this warp's 32 rows by a transpose-reduce (31
          // shuffles; lane L ends up with column L), added to the CTA's shared-memory partial
          // sums.  (mma.sync on the staged tile was tried for this and is far slower: the legacy
          // HMMA queues behind the tcgen05 MMAs in flight -- profiles/experiments/README.md.)
          warp_transpose_sum(du, lane);
          atomicAdd(&colacc[col0 + lane], du[0]);
        } else if (MODE == TC_DGRAD_BF16) {
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            __nv_bfloat162 t2 = __floats2bfloat162_rn(__uint_as_float(v[j]) * a.isf, __uint_as_float(v[j + 1]) * a.isf);
            pk[j >> 1] = *reinterpret_cast<uint32_t*>(&t2);
          }
          uint8_t* stg = staging + warp * Cfg::kStgWarp;
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
#pragma unroll
          for (int k = 0; k < 4; ++k)
            *reinterpret_cast<uint4*>(stg + lane * 64 + ((k ^ ((lane >> 1) & 3)) << 4)) =
                make_uint4(pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&map_o0, stg, col0, m_t * 128 + q * 32, net);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        } else if (MODE == TC_DGRAD_ENC) {
          // dfeat chunk -> shared-memory tile (row stride BLOCK_N+1: conflict-free for the
          // row-per-lane reads of the encode backward below); nothing goes to HBM
          float* gr = gtile + (q * 32 + lane) * (BLOCK_N + 1) + c;
#pragma unroll
          for (int j = 0; j < 32; ++j) gr[j] = __uint_as_float(v[j]) * a.isf;
        } else if (MODE == TC_DGRAD_F32 && a.out_cm) {
          // dfeat goes out COLUMN-major [net][col][row]: the 32 lanes (= 32 consecutive rows)
          // write one full 128-byte line per column, and encode_bwd reads it back coalesced
          if (row_ok) {
            float* oc = a.outf + (size_t)net * a.out_batch + (size_t)col0 * a.m_valid + row;
#pragma unroll
            for (int j = 0; j < 32; ++j) oc[(size_t)j * a.m_valid] = __uint_as_float(v[j]) * a.isf;
          }
        } else if (MODE == TC_DGRAD_F32 || MODE == TC_PLAIN_F32) {
          if (row_ok) {
            float4* o4 = reinterpret_cast<float4*>(a.outf + (size_t)net * a.out_batch + (size_t)row * a.ld_out + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              o4[j] = make_float4(__uint_as_float(v[4 * j]) * a.isf, __uint_as_float(v[4 * j + 1]) * a.isf,
                                  __uint_as_float(v[4 * j + 2]) * a.isf, __uint_as_float(v[4 * j + 3]) * a.isf);
          }
        } else {  // TC_WGRAD: grad[net*P + off + row*ld + col] += acc*isf  (split-K partial sums)
          // The 32x32 f32 chunk is transposed through this warp's staging tile so that every
          // store / atomic instruction covers 32 CONSECUTIVE floats of one output row (one or
          // two 128-byte lines) instead of one float in each of 32 rows: 32x fewer L2 requests,
          // which is what bounds the split-K reduction at small shapes.  (Parameter leaves sit
          // at arbitrary 4-byte offsets of the flat vector, so no 16-byte vector atomics.)
          float* st32 = reinterpret_cast<float*>(staging + warp * Cfg::kStgWarp);
          __syncwarp();
#pragma unroll
          for (int k = 0; k < 8; ++k)
            *reinterpret_cast<float4*>(st32 + lane * 32 + ((k ^ (lane & 7)) << 2)) =
                make_float4(__uint_as_float(v[4 * k]) * a.isf, __uint_as_float(v[4 * k + 1]) * a.isf,
                            __uint_as_float(v[4 * k + 2]) * a.isf, __uint_as_float(v[4 * k + 3]) * a.isf);
          __syncwarp();
          const int row_base = m_t * 128 + q * 32;
          float* o = a.outf + (size_t)net * a.out_batch + a.grad_off + (size_t)row_base * a.ld_out + col0 + lane;
          const bool col_ok = col0 + lane < a.n_valid;
          const int rows_here = min(32, a.m_valid - row_base);
          if (col_ok) {
            // the constant-one feature row (Dense_0): its sums are the bias gradient
            const int rb = a.bias_row - row_base;        // in [0, rows_here) only in the row group that holds it
            const int rows_k = (rb >= 0 && rb < rows_here) ? rb : rows_here;
            if (a.k_splits == 1) {
#pragma unroll 8
              for (int r = 0; r < rows_k; ++r)
                o[(size_t)r * a.ld_out] = st32[r * 32 + ((((lane >> 2) ^ (r & 7)) << 2) | (lane & 3))];
            } else {
#pragma unroll 8
              for (int r = 0; r < rows_k; ++r)
                atomicAdd(o + (size_t)r * a.ld_out, st32[r * 32 + ((((lane >> 2) ^ (r & 7)) << 2) | (lane & 3))]);
            }
            if (rows_k < rows_here)
              atomicAdd(a.outf + (size_t)net * a.out_batch + a.bias_off + col0 + lane,
                        st32[rb * 32 + ((((lane >> 2) ^ (rb & 7)) << 2) | (lane & 3))] * a.bias_rescale);
          }
        }
      }
      if (epi_tid == 0) TL((t - tile0) / tile_step, 10);
      tc_fence_before();
      if (CTA2) mbar_arrive_remote(&tempty[acc], 0);   // the leader's MMA warp owns both TMEMs' reuse
      else mbar_arrive(&tempty[acc]);
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
      if (MODE == TC_DGRAD_ENC) {
        // ---- feature-encode backward of this tile's 128 rows (same math
        // as encode_bwd_kernel<true>): a warp owns whole units, its lanes stride over the rows.
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpi) : "memory");     // tile complete
        if (t + 1 < tile_end) stage_x(t + 1, sx_nxt);
        const float* dvp = a.derived + (size_t)net * kDerivedStride;
        const int U = num_units(dm);
        const float two_pi = 6.283185307179586f;
        for (int u = warp; u < U; u += kEpi) {
          const UnitInfo ui = decode_unit(dm, u);
          int slot = 0, dim_a = 0, dim_b = -1;
          if (ui.kind == 0) { slot = 0; dim_a = ui.a; }
          else if (ui.kind == 1) { slot = 3 + ui.a; dim_a = ui.a; }
          else if (ui.kind == 2) { slot = 1; dim_a = -1; }
          else { slot = 2; dim_a = dm.inter_a[ui.a]; dim_b = dm.inter_b[ui.a]; }
          float gs = 0.f, gl_a = 0.f, gl_b = 0.f;
          // four independent rows per lane (rows >= B carry a zero dfeat row: their dU rows were
          // zero-filled by TMA, so they add nothing)
          if (ui.kind == 0) {
            const float s_x = dvp[kDvSX];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int r = lane + 32 * j;
              const float sx = sx_cur[r * (kMaxD + 1) + ui.a];
              const float G = gtile[r * (BLOCK_N + 1) + dm.col_x + ui.a];
              gs = fmaf(G, sx, gs);
              gl_a = fmaf(s_x * G, -sx, gl_a);
            }
          } else if (ui.kind == 1) {
            const int i = ui.a, d = ui.b, deg = dm.fourier_deg[i];
            const float cc = two_pi * (float)(1 << d);
            const float rden = 1.f / (float)(d + 1);
            const float kf = dvp[kDvSFourier + i] * (cc * rden);
            const int cc0 = dm.fourier_col[i] + d, cs0 = cc0 + deg;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int r = lane + 32 * j;
              const float sx = sx_cur[r * (kMaxD + 1) + i];
              float sn, cs;
              if constexpr (X3) sincosf(cc * sx, &sn, &cs); else sincos_reduced(cc * sx, &sn, &cs);
              const float Gc = gtile[r * (BLOCK_N + 1) + cc0], Gs = gtile[r * (BLOCK_N + 1) + cs0];
              gs += (Gc * cs + Gs * sn) * rden;
              gl_a = fmaf(kf * (cs * Gs - sn * Gc), -sx, gl_a);
            }
          } else if (ui.kind == 2) {
            const int k = ui.a;
            const float wk = dm.seasonal_w[k], rh = 1.f / dm.seasonal_h[k];
            const int c0 = dm.col_seasonal + k, c1i = c0 + dm.n_seasonal;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int r = lane + 32 * j;
              float sn, cs;
              if constexpr (X3) sincosf(wk * sx_cur[r * (kMaxD + 1) + dm.D], &sn, &cs);
              else sincos_reduced(wk * sx_cur[r * (kMaxD + 1) + dm.D], &sn, &cs);
              gs += (gtile[r * (BLOCK_N + 1) + c0] * cs + gtile[r * (BLOCK_N + 1) + c1i] * sn) * rh;
            }
          } else {
            const float s_i = dvp[kDvSInter];
            const int cj = dm.col_inter + ui.a;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int r = lane + 32 * j;
              const float pr = sx_cur[r * (kMaxD + 1) + dim_a] * sx_cur[r * (kMaxD + 1) + dim_b];
              const float G = gtile[r * (BLOCK_N + 1) + cj];
              gs = fmaf(G, pr, gs);
              gl_a = fmaf(s_i * G, -pr, gl_a);
            }
          }
          if (ui.kind == 3) gl_b = gl_a;
          gs = warp_sum(gs);
          gl_a = warp_sum(gl_a);
          gl_b = warp_sum(gl_b);
          if (lane == 0) {
            atomicAdd(&eacc[dm.D + slot], gs);
            if (dim_a >= 0) atomicAdd(&eacc[dim_a], gl_a);
            if (dim_b >= 0) atomicAdd(&eacc[dim_b], gl_b);
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpi) : "memory");     // sums complete, tile free
        const int nacc = dm.D + 3 + dm.D;
        const bool flush = t + 1 >= tile_end || (t + 1) / tiles_per_net != net;
        if (flush && epi_tid < nacc) {
          const float val = eacc[epi_tid];
          eacc[epi_tid] = 0.f;               // ordered before the next tile's atomics by its bar.sync
          const float* pp = a.params + (size_t)net * a.P;
          float* gp = a.gradp + (size_t)net * a.P;
          if (epi_tid < dm.D) {
            atomicAdd(&gp[dm.off_lsa + epi_tid], val);
          } else {
            const int slot = epi_tid - dm.D;
            const int off = slot == 0 ? dm.off_scale_x : (slot == 1 ? dm.off_scale_seasonal
                          : (slot == 2 ? dm.off_scale_inter : dm.fourier_scale_off[slot - 3]));
            if (off >= 0) atomicAdd(&gp[off], val * sigmoid_f(pp[off]));   // d softplus = sigmoid
          }
        }
      }
    }
    if (MODE == TC_FWD_HEAD && acc_net >= 0) head_flush(acc_net);
    if (MODE == TC_DGRAD_ACT && acc_net >= 0) {
      // the CTA's last network: flush its partial sums
      dact_scalars(acc_net);
      if (!a.skip_bias) {
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpi) : "memory");
        float* g = a.gradp + (size_t)acc_net * a.P;
        for (int i = epi_tid; i < a.n_valid; i += 32 * kEpi) atomicAdd(g + a.off_bias_prev + i, colacc[i]);
        if (epi_tid < 2) atomicAdd(g + (epi_tid == 0 ? a.off_actw : a.off_ls_prev), colacc[kAccCols + epi_tid]);
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else if (ENCODE) {
    // ===================== feature encoder (4 warps = one thread per tile row, A_MODE 2) =====================
    // models.py:216-252 fused into the Dense_0 GEMM: for every (network, m-tile) the encoder warps
    // write the 128 x Fp bf16 feature tile straight into the 128B-swizzled K-major A buffers the MMA
    // reads (two buffers: tile i+1 is generated while the MMAs / epilogues of tile i run, and the
    // tile is reused by every n-tile of the layer).  A thread owns one row: it scales its D inputs
    // once, then walks the network's unit table (column indices, argument multiplier, output scale
    // with the 1/(d+1), 1/h divisions folded in -- built when the CTA's network changes) and stores
    // one or two bf16 values per unit.  `feat` (the Dense_0 wgrad operand) is TMA-stored from the
    // same buffers when the backward pass needs it; forward-only calls never write it.
    const int et = threadIdx.x - kBaseThreads;       // = tile row
    const int U = num_units(dm);
    int enc_buf = 1, enc_key = -1, tab_net = -1; uint32_t free_phase[2] = {0u, 0u};
    for (int t = tile0; t < tile_end; t += tile_step) {
      const int key = t / a.n_tiles;
      if (key == enc_key) continue;                  // another n-tile of the m-tile already generated
      enc_key = key;
      enc_buf ^= 1;
      const int net = t / tiles_per_net;
      const int m_t = (t % tiles_per_net) / (a.n_tiles * a.k_splits);
      const float* dv = a.derived + (size_t)net * kDerivedStride;
      // the MMAs that read this buffer two m-tiles ago have retired (first use passes at once) ...
      mbar_wait(&afree[enc_buf], free_phase[enc_buf] ^ 1u);
      free_phase[enc_buf] ^= 1u;
      // ... and so has the TMA store that read it
      if (et == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      if (net != tab_net) {
        asm volatile("bar.sync 2, %0;" ::"n"(32 * kEncWarps) : "memory");     // nobody reads the old table any more
        for (int u = et; u < U; u += 32 * kEncWarps) {
          const UnitInfo ui = decode_unit(dm, u);
          EncUnitT e;
          e.kind = ui.kind; e.dim = 0; e.dim2 = 0; e.c0 = 0; e.c1 = 0; e.mult = 0.f; e.k0 = 0.f; e.pad = 0;
          if (ui.kind == 0) {
            e.dim = ui.a; e.c0 = dm.col_x + ui.a; e.k0 = dv[kDvSX];
          } else if (ui.kind == 1) {
            e.dim = ui.a; e.mult = 6.283185307179586f * (float)(1 << ui.b);
            e.c0 = dm.fourier_col[ui.a] + ui.b; e.c1 = e.c0 + dm.fourier_deg[ui.a];
            e.k0 = dv[kDvSFourier + ui.a] / (float)(ui.b + 1);
          } else if (ui.kind == 2) {
            e.dim = dm.D; e.mult = dm.seasonal_w[ui.a];
            e.c0 = dm.col_seasonal + ui.a; e.c1 = e.c0 + dm.n_seasonal;
            e.k0 = dv[kDvSSeas] / dm.seasonal_h[ui.a];
          } else {
            e.dim = dm.inter_a[ui.a]; e.dim2 = dm.inter_b[ui.a]; e.c0 = dm.col_inter + ui.a; e.k0 = dv[kDvSInter];
          }
          etab[u] = e;
        }
        tab_net = net;
      }
      asm volatile("bar.sync 2, %0;" ::"n"(32 * kEncWarps) : "memory");       // table ready, buffer free for all rows
      // this thread's row: scaled inputs x / (input_scale * exp(lsa)) and the raw time (slot D)
      const int row = min(m_t * 128 + et, a.m_valid - 1);      // rows >= B: a copy of the last row (never stored)
      const float* xr = a.x + (a.idx ? (size_t)a.idx[(size_t)net * a.idx_stride + row] : (size_t)row) * dm.D;
      float* sx = esx + et * (kMaxD + 1);
      for (int i = 0; i < dm.D; ++i) {
        const float xv = xr[i];
        sx[i] = xv / dv[kDvDenom + i];
        if (i == 0) sx[dm.D] = xv;
      }
      uint8_t* ab = abuf + enc_buf * kEncMaxKb * Cfg::kABytes;
      const int r = et;
      const uint32_t rbase = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);
      auto put = [&](int c, float v) {              // column c of this row, swizzled K-major tile of k-block c / 64
        const int cc = c & 63;
        const uint32_t off = (uint32_t)(c >> 6) * Cfg::kABytes + rbase + (uint32_t)((((cc >> 3) ^ (r & 7)) & 7) << 4) + (uint32_t)(cc & 7) * 2u;
        *reinterpret_cast<__nv_bfloat16*>(ab + off) = __float2bfloat16_rn(v);
      };
#pragma unroll 2
      for (int u = 0; u < U; ++u) {
        const EncUnitT e = etab[u];                  // warp-uniform: broadcast reads
        if (e.kind == 0) {
          put(e.c0, sx[e.dim] * e.k0);
        } else if (e.kind == 3) {
          put(e.c0, (sx[e.dim] * sx[e.dim2]) * e.k0);
        } else {
          float sn, cs;
          sincos_reduced(e.mult * sx[e.dim], &sn, &cs);   // MUFU: the bf16 rounding of the feature dominates
          put(e.c0, cs * e.k0);
          put(e.c1, sn * e.k0);
        }
      }
      if (dm.F < dm.Fp) put(dm.F, 1.f);            // constant-one feature: Dense_0 bias gradient via the wgrad GEMM
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (a.write_feat) {
        asm volatile("bar.sync 2, %0;" ::"n"(32 * kEncWarps) : "memory");     // the whole tile is written
        if (et == 0) {
          for (int kb = 0; kb < a.k_blocks; ++kb) tma_store_3d(&map_a, ab + kb * Cfg::kABytes, kb * 64, m_t * 128, net);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      mbar_arrive(&aready[enc_buf]);
    }
    if (et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync_all();        // the peer's smem / TMEM stay alive until both CTAs are done
  if (warp == kMma) {
    __syncwarp();
    tc_fence_after();
    if (CTA2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::kTmemCols) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::kTmemCols) : "memory");
  }
}

// -----------------------------------------------------------------------------
// host side: tensor maps + launch
// -----------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// bf16 tensor [net][rows][cols] (cols contiguous), box = [1][box_rows][64]
static int make_map(CUtensorMap* map, const bf16* base, uint64_t cols, uint64_t rows, uint64_t nets,
                    uint64_t row_stride_elems, uint64_t net_stride_elems, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return tc_fail(BNF_ERR_CUDA, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[3] = {cols, rows, nets};
  cuuint64_t strides[2] = {row_stride_elems * 2, net_stride_elems * 2};
  cuuint32_t box[3] = {64, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_tc_err, sizeof(g_tc_err), "cuTensorMapEncodeTiled failed (%d): cols=%llu rows=%llu nets=%llu box_rows=%u",
             (int)r, (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)nets, box_rows);
    return BNF_ERR_CUDA;
  }
  return 0;
}

// bf16 output tensor [net][rows][cols]: 32x32 boxes, 64-byte swizzle (epilogue staging tiles)
static int make_out_map(CUtensorMap* map, const bf16* base, uint64_t cols, uint64_t rows, uint64_t nets) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return tc_fail(BNF_ERR_CUDA, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[3] = {cols, rows, nets};
  cuuint64_t strides[2] = {cols * 2, rows * cols * 2};
  cuuint32_t box[3] = {32, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return tc_fail(BNF_ERR_CUDA, "cuTensorMapEncodeTiled (output map) failed");
  return 0;
}

// f32 tensor [net][rows][cols]: 32x32 boxes (128-byte rows), 128-byte swizzle (x3: z tiles)
static int make_f32_map(CUtensorMap* map, const float* base, uint64_t cols, uint64_t rows, uint64_t nets) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return tc_fail(BNF_ERR_CUDA, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[3] = {cols, rows, nets};
  cuuint64_t strides[2] = {cols * 4, rows * cols * 4};
  cuuint32_t box[3] = {32, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return tc_fail(BNF_ERR_CUDA, "cuTensorMapEncodeTiled (f32 map) failed");
  return 0;
}

struct OutMaps { CUtensorMap o0, o1; };

template <int BLOCK_N, int MN, int MODE, bool CTA2, bool X3 = false>
static int launch_tc_k(const CUtensorMap& ma, const CUtensorMap& mb, const OutMaps& om, const TcArgs& a, int sm_count, cudaStream_t st,
                       const DevModel* dm) {
  using Cfg = TcCfg<BLOCK_N, MN, CTA2, MODE, X3>;
  static DevModel dm_zero;   // zero-initialised placeholder for the non-encode instantiations
  // the attribute is per DEVICE: a process that moves to another GPU must set it there too
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    if (cudaFuncSetAttribute(tc_gemm_kernel<BLOCK_N, MN, MODE, CTA2, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem) != cudaSuccess)
      return tc_fail(BNF_ERR_CUDA, "cudaFuncSetAttribute(smem) failed");
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const int m_units = CTA2 ? (a.m_tiles + 1) / 2 : a.m_tiles;
  long long total = (long long)a.n_net * m_units * a.n_tiles * a.k_splits;
  const int slots = (CTA2 ? sm_count / 2 : sm_count) * Cfg::kMinBlocks;
  int grid = (int)(total < slots ? total : slots);
  if (grid < 1) grid = 1;
  if (CTA2) grid *= 2;
  BNF_PROF(X3 && MODE == TC_FWD_HEAD ? "tc_fwd_head_x3" : X3 && MODE == TC_FWD ? "tc_gemm_fwd_x3" : X3 && MODE == TC_DGRAD_ACT ? "tc_gemm_dgrad_x3" : MN == 2 ? "tc_encode_fwd0" : MODE == TC_DGRAD_ENC ? "tc_dgrad0_enc" : MODE == TC_FWD_HEAD ? "tc_fwd_head" : MODE == TC_FWD ? "tc_gemm_fwd" : (MODE == TC_WGRAD ? "tc_gemm_wgrad" : (MODE == TC_PLAIN_F32 ? "tc_gemm_plain" : "tc_gemm_dgrad")), st);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(Cfg::kThreads);
  cfg.dynamicSmemBytes = Cfg::kSmem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTA2 ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (pdl_active()) {
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  const DevModel& dmr = dm ? *dm : dm_zero;
#ifdef BNF_TC_EXPERIMENT
  TcArgs ax = a;
  { const char* d = getenv("BNF_TC_DBG"); ax.dbg = d ? atoi(d) : 0; }
  static long long* d_tl = nullptr;
  const char* tle = getenv("BNF_TC_TL");            // = MODE number to trace (needs BNF_NO_GRAPH=1)
  const bool trace = tle && atoi(tle) == MODE;
  if (trace) {
    if (!d_tl) cudaMalloc(&d_tl, 512 * 256 * sizeof(long long));
    cudaMemsetAsync(d_tl, 0, 512 * 256 * sizeof(long long), st);
    ax.tl = d_tl;
  }
  cudaError_t e = cudaLaunchKernelEx(&cfg, tc_gemm_kernel<BLOCK_N, MN, MODE, CTA2, X3>, ma, mb, om.o0, om.o1, ax, dmr);
  if (trace) {
    static int n_traced = 0;
    cudaStreamSynchronize(st);
    if (++n_traced == 5) {                          // one warm launch
      static long long h[512 * 256];
      cudaMemcpy(h, d_tl, sizeof(h), cudaMemcpyDeviceToHost);
      for (int c : {0, 1, 2, 100, 147}) {
        const long long t0 = h[(c * 16 + 0) * 16 + 15];
        for (int ti = 0; ti < 6; ++ti) {
          fprintf(stderr, "TL mode %d cta %3d tile %d:", MODE, c, ti);
          for (int ev = 0; ev < 15; ++ev) {
            const long long v = h[(c * 16 + ti) * 16 + ev];
            fprintf(stderr, " %7lld", v ? v - t0 : -1);
          }
          fprintf(stderr, "\n");
        }
      }
    }
  }
#else
  cudaError_t e = cudaLaunchKernelEx(&cfg, tc_gemm_kernel<BLOCK_N, MN, MODE, CTA2, X3>, ma, mb, om.o0, om.o1, a, dmr);
#endif
  if (e != cudaSuccess) {
    snprintf(g_tc_err, sizeof(g_tc_err), "tc_gemm_kernel launch failed: %s", cudaGetErrorString(e));
    return BNF_ERR_CUDA;
  }
  return 0;
}

// CTA pairs (cta_group::2) serve the 256-wide GEMMs (measured on the wind shard: fwd -18 %,
// wgrad -15 %, fused dgrad -5 % time); BNF_CTA2=0 falls back to single-CTA tiles.
static bool want_cta2(const TcArgs& a, int block_n) {
  const char* e = getenv("BNF_CTA2");
  if (e && e[0] == '0') return false;
  return block_n == 256 && a.m_tiles >= 2;
}

template <int BLOCK_N, int MN, int MODE, bool X3 = false>
static int launch_tc_m(const CUtensorMap& ma, const CUtensorMap& mb, const OutMaps& om, const TcArgs& a, int sm_count, cudaStream_t st,
                       const DevModel* dm = nullptr) {
  if constexpr (BLOCK_N == 256 && MN != 2 && (MODE == TC_FWD || MODE == TC_FWD_HEAD || MODE == TC_DGRAD_ACT || MODE == TC_WGRAD || MODE == TC_PLAIN_F32)) {
    if (want_cta2(a, BLOCK_N)) return launch_tc_k<BLOCK_N, MN, MODE, true, X3>(ma, mb, om, a, sm_count, st, dm);
  }
  return launch_tc_k<BLOCK_N, MN, MODE, false, X3>(ma, mb, om, a, sm_count, st, dm);
}

// one kernel instantiation per (tile width, operand mode, epilogue): each carries only its own
// epilogue code (the whole hot loop stays resident in the instruction cache)
template <int BLOCK_N, int MN>
static int launch_tc(const CUtensorMap& ma, const CUtensorMap& mb, const OutMaps& om, const TcArgs& a, int sm, cudaStream_t st,
                     const DevModel* dm = nullptr) {
  if constexpr (MN == 2) {
    return launch_tc_m<BLOCK_N, 2, TC_FWD>(ma, mb, om, a, sm, st, dm);
  } else if constexpr (MN == 1) {
    if (a.mode == TC_WGRAD) return launch_tc_m<BLOCK_N, 1, TC_WGRAD>(ma, mb, om, a, sm, st, dm);
    return launch_tc_m<BLOCK_N, 1, TC_PLAIN_F32>(ma, mb, om, a, sm, st, dm);
  } else if constexpr (MN == 3) {
    if (a.mode == TC_FWD_HEAD && a.x3) return launch_tc_m<BLOCK_N, 3, TC_FWD_HEAD, true>(ma, mb, om, a, sm, st, dm);
    if (a.mode == TC_FWD_HEAD) return launch_tc_m<BLOCK_N, 3, TC_FWD_HEAD>(ma, mb, om, a, sm, st, dm);
    if (a.mode == TC_FWD && a.x3) return launch_tc_m<BLOCK_N, 3, TC_FWD, true>(ma, mb, om, a, sm, st, dm);
    if (a.mode == TC_FWD) return launch_tc_m<BLOCK_N, 3, TC_FWD>(ma, mb, om, a, sm, st, dm);
    return launch_tc_m<BLOCK_N, 3, TC_PLAIN_F32>(ma, mb, om, a, sm, st, dm);
  } else {
    switch (a.mode) {
      case TC_FWD: return launch_tc_m<BLOCK_N, 0, TC_FWD>(ma, mb, om, a, sm, st, dm);
      case TC_DGRAD_ACT:
        if (a.x3) return launch_tc_m<BLOCK_N, 0, TC_DGRAD_ACT, true>(ma, mb, om, a, sm, st, dm);
        return launch_tc_m<BLOCK_N, 0, TC_DGRAD_ACT>(ma, mb, om, a, sm, st, dm);
      case TC_DGRAD_BF16: return launch_tc_m<BLOCK_N, 0, TC_DGRAD_BF16>(ma, mb, om, a, sm, st, dm);
      case TC_DGRAD_F32: return launch_tc_m<BLOCK_N, 0, TC_DGRAD_F32>(ma, mb, om, a, sm, st, dm);
      case TC_DGRAD_ENC:
        if constexpr (BLOCK_N <= 128) {
          if (a.x3) return launch_tc_m<BLOCK_N, 0, TC_DGRAD_ENC, true>(ma, mb, om, a, sm, st, dm);
          return launch_tc_m<BLOCK_N, 0, TC_DGRAD_ENC>(ma, mb, om, a, sm, st, dm);
        } else return tc_fail(BNF_ERR_INVALID, "TC_DGRAD_ENC needs one n-tile of <= 128 columns");
      default: return launch_tc_m<BLOCK_N, 0, TC_PLAIN_F32>(ma, mb, om, a, sm, st, dm);
    }
  }
}

template <int MN>
static int launch_tc_n(int block_n, const CUtensorMap& ma, const CUtensorMap& mb, const OutMaps& om, const TcArgs& a, int sm, cudaStream_t st,
                       const DevModel* dm = nullptr) {
  switch (block_n) {
    case 256: return launch_tc<256, MN>(ma, mb, om, a, sm, st, dm);
    case 128: return launch_tc<128, MN>(ma, mb, om, a, sm, st, dm);
    case 64: return launch_tc<64, MN>(ma, mb, om, a, sm, st, dm);
  }
  return tc_fail(BNF_ERR_INVALID, "bad BLOCK_N");
}

static int pick_block_n(int n) { return n % 256 == 0 ? 256 : (n % 128 == 0 ? 128 : 64); }
// experiment hook: BNF_BN_FWD / BNF_BN_DGRAD = 128 or 64 force narrower output tiles for the plain
// forward / fused dgrad kernels (more, smaller tiles: less wave quantisation, more per-tile overhead)
static int pick_block_n_env(int n, const char* env) {
  const char* e = getenv(env);
  const int want = e ? atoi(e) : 0;
  if ((want == 128 || want == 64) && n % want == 0) return want;
  return pick_block_n(n);
}
static int sm_count_of(const bnf_plan* p) {
  if (p->sm_count > 0) return p->sm_count;
  int dev = 0, n = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n;
}

// The forward B operand: the natural (in,out) bf16 copy `wn` read MN-major (default), or the
// transposed copy `wt` read K-major (BNF_FWD_WT=1, the round-1 layout; kept for A/B timing).
bool tc_fwd_uses_wt() {
  const char* e = getenv("BNF_FWD_WT");   // read per call: tests flip it
  return e && e[0] == '1';
}

int tc_fwd_layer(const bnf_plan* p, int layer, const float* params, const float* derived, const bf16* a_in,
                 const bf16* wt, const bf16* wn, bf16* z, bf16* h, int n_net, int B, cudaStream_t st,
                 bool x3, float* zf) {
  const DevModel& m = p->m;
  const int Kp = kp_of(m, layer);
  int bn = pick_block_n_env(m.W, "BNF_BN_FWD");
  // bf16x3 Dense_0 at W = 256: the GEMM is short (K = Fp) and the epilogue writes 10 bytes per element,
  // so the finer 128-wide tiles win (less wave quantisation: measured 62.3 -> 56.3 us, r2j-2); every
  // other forward shape is faster with one 256-wide CTA-pair tile per row block
  if (x3 && layer == 0 && m.W == 256 && !getenv("BNF_BN_FWD")) bn = 128;
  if (x3) {
    // split operands: a_in [n_net,B,3*Kp], wn [layer][Kp][3*W]; outputs zf [n_net,B,W] f32 (may be
    // NULL: forward only) and h [n_net,B,3*W]
    CUtensorMap ma, mb;
    int rc = make_map(&ma, a_in, 3 * (uint64_t)Kp, B, n_net, 3 * (uint64_t)Kp, (uint64_t)B * 3 * Kp, 128);
    if (rc) return rc;
    if ((rc = make_map(&mb, wn + 3 * layer_off(m, layer), 3 * (uint64_t)m.W, Kp, n_net, 3 * (uint64_t)m.W, 3 * tc_weight_elems(m), 64))) return rc;
    TcArgs a;
    memset(&a, 0, sizeof(a));
    a.mode = TC_FWD; a.n_net = n_net;
    a.m_tiles = (B + 127) / 128; a.n_tiles = m.W / bn; a.k_splits = 1;
    a.x3 = 1; a.kseg = Kp / 64; a.k_blocks = 6 * a.kseg; a.a_pstride = Kp; a.b_pstride = m.W;
    a.x3_pa = kX3PlaneA; a.x3_pb = kX3PlaneB;
    a.m_valid = B; a.n_valid = m.W;
    a.params = params; a.derived = derived; a.P = m.P; a.off_bias = m.off_bias[layer]; a.layer = layer;
    a.isf = layer == 0 ? m.inv_sqrt_F : m.inv_sqrt_W;
    a.out0 = (bf16*)zf; a.out1 = h; a.out_batch = (long long)B * m.W; a.ld_out = m.W;
    OutMaps om;
    memset(&om, 0, sizeof(om));
    if ((rc = make_out_map(&om.o1, h, 3 * (uint64_t)m.W, B, n_net))) return rc;
    if (zf && (rc = make_f32_map(&om.o0, zf, m.W, B, n_net))) return rc;
    return launch_tc_n<3>(bn, ma, mb, om, a, sm_count_of(p), st);
  }
  const bool use_wt = tc_fwd_uses_wt() || wn == nullptr;
  CUtensorMap ma, mb;
  int rc = make_map(&ma, a_in, Kp, B, n_net, Kp, (uint64_t)B * Kp, 128);
  if (rc) return rc;
  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.mode = TC_FWD; a.n_net = n_net;
  a.m_tiles = (B + 127) / 128; a.n_tiles = m.W / bn; a.k_splits = 1; a.k_blocks = Kp / 64;
  if (use_wt) {
    // a CTA pair loads the B tile in two halves (one per CTA)
    rc = make_map(&mb, wt + layer_off(m, layer), Kp, m.W, n_net, Kp, tc_weight_elems(m), want_cta2(a, bn) ? bn / 2 : bn);
  } else {
    // wn [Kp][W]: boxes of [64 reduction rows][64 output columns]
    rc = make_map(&mb, wn + layer_off(m, layer), m.W, Kp, n_net, m.W, tc_weight_elems(m), 64);
  }
  if (rc) return rc;
  a.m_valid = B; a.n_valid = m.W;
  a.params = params; a.derived = derived; a.P = m.P; a.off_bias = m.off_bias[layer]; a.layer = layer;
  a.isf = layer == 0 ? m.inv_sqrt_F : m.inv_sqrt_W;
  a.out0 = z; a.out1 = h; a.out_batch = (long long)B * m.W; a.ld_out = m.W;
  OutMaps om;
  memset(&om, 0, sizeof(om));
  if ((rc = make_out_map(&om.o1, h, m.W, B, n_net))) return rc;
  if (z && (rc = make_out_map(&om.o0, z, m.W, B, n_net))) return rc;
  if (use_wt) return launch_tc_n<0>(bn, ma, mb, om, a, sm_count_of(p), st);
  return launch_tc_n<3>(bn, ma, mb, om, a, sm_count_of(p), st);
}

// Last hidden layer of a TRAINING step: forward GEMM + head + log-likelihood + the layer's own
// activation backward in one kernel (TC_FWD_HEAD).  Needs whole rows in one tile (W <= 256) and
// the natural-layout weight copy.  Outputs: dU [n_net,B,W] bf16, ll[n_net] += loglik, grad +=
// Dense_L / bias / scale / likelihood-parameter gradients.
bool tc_fwd_head_supported(const DevModel& m) {
  const char* e = getenv("BNF_NO_FUSED_HEAD_EPI");
  if (e && e[0] == '1') return false;
  return pick_block_n(m.W) == m.W && !tc_fwd_uses_wt();
}

int tc_fwd_head(const bnf_plan* p, const float* params, const float* derived, const bf16* a_in, const bf16* wn,
                const float* y, const int32_t* idx, int64_t idx_stride, bf16* dU, float* ll, float* grad,
                int n_net, int B, cudaStream_t st, bool x3) {
  const DevModel& m = p->m;
  const int layer = m.L - 1;
  const int Kp = kp_of(m, layer), bn = pick_block_n(m.W);
  if (!tc_fwd_head_supported(m)) return tc_fail(BNF_ERR_UNSUPPORTED, "fused head epilogue needs W in {64,128,256}");
  CUtensorMap ma, mb;
  const uint64_t pl = x3 ? 3 : 1;     // planes side by side in every operand row
  int rc = make_map(&ma, a_in, pl * Kp, B, n_net, pl * Kp, (uint64_t)B * pl * Kp, 128);
  if (rc) return rc;
  if ((rc = make_map(&mb, wn + pl * layer_off(m, layer), pl * m.W, Kp, n_net, pl * m.W, pl * tc_weight_elems(m), 64))) return rc;
  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.mode = TC_FWD_HEAD; a.n_net = n_net;
  a.m_tiles = (B + 127) / 128; a.n_tiles = 1; a.k_splits = 1; a.k_blocks = Kp / 64;
  if (x3) { a.x3 = 1; a.kseg = Kp / 64; a.k_blocks = 6 * a.kseg; a.a_pstride = Kp; a.b_pstride = m.W; a.x3_pa = kX3PlaneA; a.x3_pb = kX3PlaneB; }
  a.m_valid = B; a.n_valid = m.W;
  a.params = params; a.derived = derived; a.P = m.P; a.off_bias = m.off_bias[layer]; a.layer = layer;
  a.isf = layer == 0 ? m.inv_sqrt_F : m.inv_sqrt_W;
  a.y = y; a.ll = ll; a.idx = idx; a.idx_stride = idx_stride; a.gradp = grad;
  OutMaps om;
  memset(&om, 0, sizeof(om));
  if ((rc = make_out_map(&om.o0, dU, (x3 ? (uint64_t)kDuPlanes : 1) * m.W, B, n_net))) return rc;
  return launch_tc_n<3>(bn, ma, mb, om, a, sm_count_of(p), st, &m);
}

// Fused feature encode + Dense_0 (models.py:216-268 for the first layer): the A operand is
// generated on-chip from (x, idx) by encoder warps and stays resident for all n-tiles of its
// m-tile; B is the natural (in,out) bf16 kernel copy read MN-major (the copy the fused MAP update
// maintains).  `feat` (may be NULL) receives the tiles as a by-product for the Dense_0 wgrad.
bool tc_fused_encode_supported(const DevModel& m) {
  int U = m.D + m.n_seasonal + m.n_inter;
  for (int i = 0; i < m.D; ++i) U += m.fourier_deg[i] > 0 ? m.fourier_deg[i] : 0;
  return m.Fp <= 64 * kEncMaxKb && U <= kEncMaxUnits && !tc_fwd_uses_wt();
}
// Policy (measured, profiles/experiments/README.md r2f): forward-only calls (predict) use the fused
// kernel -- `feat` then never exists in HBM.  In a TRAINING step `feat` is needed again by the
// Dense_0 wgrad, so fusing saves only its (L2-resident) re-read, while the encoder warps share the
// issue slots of an epilogue-bound kernel and one programmatic-launch overlap is lost: 0.184 vs
// 0.174 ms/step at W=256, a tie at W=1024.  Training therefore keeps encode kernel + GEMM unless
// BNF_FUSED_ENCODE=1; BNF_NO_FUSED_ENCODE=1 switches the fused kernel off everywhere.
bool tc_fused_encode_wanted(const DevModel& m, bool training) {
  if (!tc_fused_encode_supported(m)) return false;
  const char* off = getenv("BNF_NO_FUSED_ENCODE");
  if (off && off[0] == '1') return false;
  const char* on = getenv("BNF_FUSED_ENCODE");
  if (on && on[0] == '1') return true;
  return !training;
}

int tc_fwd_layer0_fused(const bnf_plan* p, const float* params, const float* derived, const float* x,
                        const int32_t* idx, int64_t idx_stride, const bf16* wn, bf16* feat, bf16* z, bf16* h,
                        int n_net, int B, cudaStream_t st) {
  const DevModel& m = p->m;
  const int Kp = m.Fp, bn = pick_block_n(m.W);
  if (!tc_fused_encode_supported(m)) return tc_fail(BNF_ERR_UNSUPPORTED, "fused encode + Dense_0 needs <= 128 padded features");
  CUtensorMap ma, mb;
  int rc;
  memset(&ma, 0, sizeof(ma));
  if (feat && (rc = make_map(&ma, feat, Kp, B, n_net, Kp, (uint64_t)B * Kp, 128))) return rc;
  // wn [Kp][W]: boxes of [64 reduction rows][64 output columns]
  if ((rc = make_map(&mb, wn, m.W, Kp, n_net, m.W, tc_weight_elems(m), 64))) return rc;
  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.mode = TC_FWD; a.n_net = n_net;
  a.m_tiles = (B + 127) / 128; a.n_tiles = m.W / bn; a.k_splits = 1; a.k_blocks = Kp / 64;
  a.m_valid = B; a.n_valid = m.W;
  a.params = params; a.derived = derived; a.P = m.P; a.off_bias = m.off_bias[0]; a.layer = 0;
  a.isf = m.inv_sqrt_F;
  a.out0 = z; a.out1 = h; a.out_batch = (long long)B * m.W; a.ld_out = m.W;
  a.x = x; a.idx = idx; a.idx_stride = idx_stride;
  a.write_feat = feat ? 1 : 0;
  OutMaps om;
  memset(&om, 0, sizeof(om));
  if ((rc = make_out_map(&om.o1, h, m.W, B, n_net))) return rc;
  if (z && (rc = make_out_map(&om.o0, z, m.W, B, n_net))) return rc;
  return launch_tc_n<2>(bn, ma, mb, om, a, sm_count_of(p), st, &m);
}

// bf16x3 variant of the fused dgrad + activation backward: dU [n_net,B,3*W] and wn [layer][Kp][3*W]
// are split operands, z_prev [n_net,B,Kp] is f32, out [n_net,B,3*Kp] receives dU of layer-1 split.
int tc_dgrad_act_x3(const bnf_plan* p, int layer, const bf16* wn3, const bf16* dU, bf16* out, const float* z_prev,
                    const float* params, const float* derived, float* grad, int n_net, int B, cudaStream_t st) {
  const DevModel& m = p->m;
  const int Kp = kp_of(m, layer), bn = pick_block_n_env(Kp, "BNF_BN_DGRAD");
  if (layer < 1 || Kp > kAccCols) return tc_fail(BNF_ERR_UNSUPPORTED, "fused dgrad + activation backward needs W <= 1024");
  CUtensorMap ma, mb;
  int rc = make_map(&ma, dU, kDuPlanes * (uint64_t)m.W, B, n_net, kDuPlanes * (uint64_t)m.W, (uint64_t)B * kDuPlanes * m.W, 128);
  if (rc) return rc;
  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.mode = TC_DGRAD_ACT; a.n_net = n_net;
  a.m_tiles = (B + 127) / 128; a.n_tiles = Kp / bn; a.k_splits = 1;
  if ((rc = make_map(&mb, wn3 + 3 * layer_off(m, layer), 3 * (uint64_t)m.W, Kp, n_net, 3 * (uint64_t)m.W, 3 * tc_weight_elems(m),
                     want_cta2(a, bn) ? bn / 2 : bn))) return rc;
  a.x3 = 1; a.kseg = m.W / 64; a.k_blocks = kX3BwdSegs * a.kseg; a.a_pstride = m.W; a.b_pstride = m.W;
  a.x3_pa = kX3DgradA; a.x3_pb = kX3DgradB;
  a.m_valid = B; a.n_valid = Kp;
  a.isf = m.inv_sqrt_W;
  a.zin = (const bf16*)z_prev; a.gradp = grad; a.params = params; a.derived = derived; a.P = m.P;
  a.skip_bias = (layer == 1 && tc_bias0_via_wgrad(m)) ? 1 : 0;
  a.layer_prev = layer - 1; a.off_bias_prev = m.off_bias[layer - 1];
  a.off_ls_prev = m.off_layer_scale[layer - 1]; a.off_actw = m.off_actw;
  a.out0 = out; a.out_batch = (long long)B * Kp; a.ld_out = Kp;
  OutMaps om;
  memset(&om, 0, sizeof(om));
  if ((rc = make_out_map(&om.o0, out, kDuPlanes * (uint64_t)Kp, B, n_net))) return rc;
  if ((rc = make_f32_map(&om.o1, z_prev, Kp, B, n_net))) return rc;
  return launch_tc_n<0>(bn, ma, mb, om, a, sm_count_of(p), st);
}

int tc_dgrad(const bnf_plan* p, int layer, const bf16* wn, const bf16* dU, bf16* out_bf, float* out_f32,
             int n_net, int B, cudaStream_t st, const bf16* z_prev, const float* params,
             const float* derived, float* grad) {
  const DevModel& m = p->m;
  const int Kp = kp_of(m, layer);          // output columns (in-features, padded)
  const int bn = z_prev ? pick_block_n_env(Kp, "BNF_BN_DGRAD") : pick_block_n(Kp);
  CUtensorMap ma, mb;
  int rc = make_map(&ma, dU, m.W, B, n_net, m.W, (uint64_t)B * m.W, 128);
  if (rc) return rc;
  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.m_tiles = (B + 127) / 128;
  a.mode = out_bf ? (z_prev ? TC_DGRAD_ACT : TC_DGRAD_BF16) : TC_DGRAD_F32; a.n_net = n_net;
  rc = make_map(&mb, wn + layer_off(m, layer), m.W, Kp, n_net, m.W, tc_weight_elems(m),
                (a.mode == TC_DGRAD_ACT && want_cta2(a, bn)) ? bn / 2 : bn);
  if (rc) return rc;
  if (z_prev) {
    if (Kp > kAccCols) return tc_fail(BNF_ERR_UNSUPPORTED, "fused dgrad + activation backward needs W <= 1024");
    a.zin = z_prev; a.gradp = grad; a.params = params; a.derived = derived; a.P = m.P;
    a.skip_bias = (layer == 1 && tc_bias0_via_wgrad(m)) ? 1 : 0;
    a.layer_prev = layer - 1; a.off_bias_prev = m.off_bias[layer - 1];
    a.off_ls_prev = m.off_layer_scale[layer - 1]; a.off_actw = m.off_actw;
  }
  a.m_tiles = (B + 127) / 128; a.n_tiles = Kp / bn; a.k_splits = 1; a.k_blocks = m.W / 64;
  a.m_valid = B; a.n_valid = Kp;
  a.isf = layer == 0 ? m.inv_sqrt_F : m.inv_sqrt_W;
  a.out0 = out_bf; a.outf = out_f32; a.out_batch = (long long)B * Kp; a.ld_out = Kp;
  a.out_cm = out_f32 != nullptr;   // layer-0 dgrad: dfeat is column-major (see launch_encode_bwd)
  OutMaps om;
  memset(&om, 0, sizeof(om));
  if (out_bf && (rc = make_out_map(&om.o0, out_bf, Kp, B, n_net))) return rc;
  // the z tiles of the fused activation backward are TMA-loaded through the second map
  if (z_prev && (rc = make_out_map(&om.o1, z_prev, Kp, B, n_net))) return rc;
  return launch_tc_n<0>(bn, ma, mb, om, a, sm_count_of(p), st);
}

bool tc_dgrad_act_supported(const DevModel& m) { return m.W <= kAccCols; }
// Dense_0 bias gradient as row F of the Dense_0 wgrad GEMM (constant-one feature in the first pad
// column of feat): needs a pad column and the fused dgrad + activation backward feeding layer 0.
bool tc_bias0_via_wgrad(const DevModel& m) {
  const char* e = getenv("BNF_NO_BIAS0_WGRAD");
  if (e && e[0] == '1') return false;
  return m.F < m.Fp && m.L >= 2 && tc_dgrad_act_supported(m);
}

// Layer-0 dgrad fused with the feature-encode backward: dfeat = isf * dU_0 @ K_0^T never leaves
// the SM (TMEM -> shared-memory tile); the epilogue warps reduce it against the regenerated
// features into the feature-scale / log_scale_adjustment gradients.  One n-tile must cover all
// padded features (Fp <= 128).
bool tc_dgrad0_enc_supported(const DevModel& m) {
  const char* e = getenv("BNF_NO_FUSED_ENC_BWD");
  if (e && e[0] == '1') return false;
  return m.Fp <= 128 && pick_block_n(m.Fp) == m.Fp;
}

int tc_dgrad0_enc(const bnf_plan* p, const bf16* wn, const bf16* dU, const float* x, const int32_t* idx,
                  int64_t idx_stride, const float* params, const float* derived, float* grad, int n_net,
                  int B, cudaStream_t st, bool x3) {
  const DevModel& m = p->m;
  const int Kp = m.Fp, bn = pick_block_n(Kp);
  if (!tc_dgrad0_enc_supported(m)) return tc_fail(BNF_ERR_UNSUPPORTED, "fused dgrad0 + encode backward needs Fp <= 128");
  CUtensorMap ma, mb;
  const uint64_t pl = x3 ? 3 : 1, pd = x3 ? kDuPlanes : 1;     // planes side by side in a kernel / dU row
  int rc = make_map(&ma, dU, pd * m.W, B, n_net, pd * m.W, (uint64_t)B * pd * m.W, 128);
  if (rc) return rc;
  if ((rc = make_map(&mb, wn, pl * m.W, Kp, n_net, pl * m.W, pl * tc_weight_elems(m), bn))) return rc;
  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.mode = TC_DGRAD_ENC; a.n_net = n_net;
  a.m_tiles = (B + 127) / 128; a.n_tiles = 1; a.k_splits = 1; a.k_blocks = m.W / 64;
  if (x3) { a.x3 = 1; a.kseg = m.W / 64; a.k_blocks = kX3BwdSegs * a.kseg; a.a_pstride = m.W; a.b_pstride = m.W; a.x3_pa = kX3DgradA; a.x3_pb = kX3DgradB; }
  a.m_valid = B; a.n_valid = Kp;
  a.isf = m.inv_sqrt_F;
  a.x = x; a.idx = idx; a.idx_stride = idx_stride;
  a.params = params; a.derived = derived; a.gradp = grad; a.P = m.P;
  OutMaps om;
  memset(&om, 0, sizeof(om));
  return launch_tc_n<0>(bn, ma, mb, om, a, sm_count_of(p), st, &m);
}

int tc_wgrad(const bnf_plan* p, int layer, const bf16* a_in, const bf16* dU, float* grad, int n_net, int B,
             cudaStream_t st, bool x3, bool bias0) {
  const DevModel& m = p->m;
  const int Kp = kp_of(m, layer), Kin = layer == 0 ? m.F : m.W, bn = pick_block_n(m.W);
  CUtensorMap ma, mb;
  const uint64_t pl = x3 ? 3 : 1;     // planes side by side in every operand row
  // MN-major operands: boxes of [64 batch rows][64 contiguous features]
  int rc = make_map(&ma, a_in, pl * Kp, B, n_net, pl * Kp, (uint64_t)B * pl * Kp, 64);
  if (rc) return rc;
  const uint64_t pd = x3 ? kDuPlanes : 1;
  rc = make_map(&mb, dU, pd * m.W, B, n_net, pd * m.W, (uint64_t)B * pd * m.W, 64);
  if (rc) return rc;
  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.mode = TC_WGRAD; a.n_net = n_net;
  a.m_tiles = (Kp + 127) / 128; a.n_tiles = m.W / bn; a.k_blocks = (B + 63) / 64;
  if (x3) {
    a.x3 = 1; a.kseg = a.k_blocks; a.k_blocks = kX3BwdSegs * a.kseg; a.a_pstride = Kp; a.b_pstride = m.W;
    a.x3_pa = kX3WgradA; a.x3_pb = kX3WgradB;
    // The reduction runs over the BATCH here: segment after segment would sweep all rows of the
    // operand planes five times, and the planes (3 + 2 per activation pair) exceed the L2 at the
    // benchmark shapes -- ncu: 339 MB from DRAM for 256 MB of operands.  Interleaved, the five
    // products of one 64-row block are consecutive k-blocks, so every re-read of a plane tile hits
    // L2.  Price: the accumulator is at full magnitude for all five products of a block (the
    // round-toward-zero bias grows ~5x); the accumulations are short (<= 64 k-blocks per split) and
    // gradients are judged at 1e-4.
    const char* il = getenv("BNF_X3_WGRAD_SEGMENTED");
    a.x3_nseg = (il && il[0] == '1') ? 0 : kX3BwdSegs;
  }
  const int sm = sm_count_of(p);
  const bool pair = want_cta2(a, bn);
  const int slots = pair ? sm / 2 : sm;
  long long base_tiles = (long long)n_net * (pair ? (a.m_tiles + 1) / 2 : a.m_tiles) * a.n_tiles;
  int splits = 1;
  if (base_tiles < slots) splits = (int)(slots / base_tiles);   // fill one wave, never a ragged second one
  if (splits > a.k_blocks / 4) splits = a.k_blocks / 4;   // >= 4 k-blocks per split
  if (splits < 1) splits = 1;
  if (x3) {
    // tcgen05 accumulates with round-toward-zero, one truncation per UMMA_K step: keep every TMEM
    // accumulation short (<= 64 k-blocks = 4096 batch rows, ~4e-6 of the partial sum; measured in
    // profiles/experiments/README.md r2a-2) and let the f32 atomics (round-to-nearest) add the splits
    const int min_splits = (a.k_blocks + 63) / 64;
    if (splits < min_splits) splits = min_splits;
  }
  // every split must own at least one k-block
  while (splits > 1 && (a.k_blocks + splits - 1) / splits * (splits - 1) >= a.k_blocks) --splits;
  a.k_splits = splits;
  a.m_valid = Kin; a.n_valid = m.W;
  a.bias_row = -1;
  if (layer == 0 && bias0) {        // row F of the accumulator = isf * sum_b dU_0 = the Dense_0 bias gradient
    a.m_valid = Kin + 1; a.bias_row = Kin; a.bias_off = m.off_bias[0]; a.bias_rescale = 1.f / m.inv_sqrt_F;
  }
  a.isf = layer == 0 ? m.inv_sqrt_F : m.inv_sqrt_W;
  a.outf = grad; a.out_batch = m.P; a.ld_out = m.W; a.grad_off = m.off_kernel[layer];
  OutMaps om;
  memset(&om, 0, sizeof(om));
  return launch_tc_n<1>(bn, ma, mb, om, a, sm, st);
}

// debug / test entry: plain C[net][M][N] (f32) = A x B in either operand layout
int tc_debug_gemm(int mn_major, const bf16* A, const bf16* Bm, float* C, int n_net, int M, int N, int K,
                  int sm_count, cudaStream_t st) {
  const int bn = pick_block_n(N);
  if (N % 64 != 0 || (K % 64 != 0 && mn_major != 1 && mn_major != 4)) return tc_fail(BNF_ERR_INVALID, "N, K must be multiples of 64");
  CUtensorMap ma, mb;
  int rc;
  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.mode = TC_PLAIN_F32; a.n_net = n_net; a.k_splits = 1; a.isf = 1.f;
  a.m_tiles = (M + 127) / 128; a.n_tiles = N / bn; a.k_blocks = (K + 63) / 64;
  a.m_valid = M; a.n_valid = N; a.outf = C; a.out_batch = (long long)M * N; a.ld_out = N;
  OutMaps om;
  memset(&om, 0, sizeof(om));
  if (mn_major >= 3 && mn_major <= 5) {
    // split-operand (bf16x3) GEMMs: every operand row holds three bf16 planes side by side
    //   3: A [net][M][3K] K-major,  B [net][K][3N] MN-major   (forward)
    //   4: A [net][K][3M] MN-major, B [net][K][3N] MN-major   (wgrad)
    //   5: A [net][M][3K] K-major,  B [net][N][3K] K-major    (dgrad)
    if (K % 64 != 0 && mn_major != 4) return tc_fail(BNF_ERR_INVALID, "K must be a multiple of 64");
    a.x3 = 1; a.kseg = a.k_blocks; a.k_blocks = 6 * a.kseg; a.x3_pa = kX3PlaneA; a.x3_pb = kX3PlaneB;
    const uint64_t M3 = 3 * (uint64_t)M, N3 = 3 * (uint64_t)N, K3 = 3 * (uint64_t)K;
    if (mn_major == 3) {
      a.a_pstride = K; a.b_pstride = N;
      if ((rc = make_map(&ma, A, K3, M, n_net, K3, (uint64_t)M * K3, 128))) return rc;
      if ((rc = make_map(&mb, Bm, N3, K, n_net, N3, (uint64_t)K * N3, 64))) return rc;
      return launch_tc_n<3>(bn, ma, mb, om, a, sm_count, st);
    }
    if (mn_major == 4) {
      a.a_pstride = M; a.b_pstride = N;
      if ((rc = make_map(&ma, A, M3, K, n_net, M3, (uint64_t)K * M3, 64))) return rc;
      if ((rc = make_map(&mb, Bm, N3, K, n_net, N3, (uint64_t)K * N3, 64))) return rc;
      return launch_tc_n<1>(bn, ma, mb, om, a, sm_count, st);
    }
    a.a_pstride = K; a.b_pstride = K;
    if ((rc = make_map(&ma, A, K3, M, n_net, K3, (uint64_t)M * K3, 128))) return rc;
    if ((rc = make_map(&mb, Bm, K3, N, n_net, K3, (uint64_t)N * K3, want_cta2(a, bn) ? bn / 2 : bn))) return rc;
    return launch_tc_n<0>(bn, ma, mb, om, a, sm_count, st);
  }
  if (mn_major == 2) {   // A [net][M][K] (K-major), B [net][K][N] (MN-major)
    if ((rc = make_map(&ma, A, K, M, n_net, K, (uint64_t)M * K, 128))) return rc;
    if ((rc = make_map(&mb, Bm, N, K, n_net, N, (uint64_t)N * K, 64))) return rc;
    return launch_tc_n<3>(bn, ma, mb, om, a, sm_count, st);
  }
  if (!mn_major) {   // A [net][M][K], B [net][N][K]
    if ((rc = make_map(&ma, A, K, M, n_net, K, (uint64_t)M * K, 128))) return rc;
    if ((rc = make_map(&mb, Bm, K, N, n_net, K, (uint64_t)N * K, want_cta2(a, bn) ? bn / 2 : bn))) return rc;
    return launch_tc_n<0>(bn, ma, mb, om, a, sm_count, st);
  }
  // A [net][K][M], B [net][K][N]
  if ((rc = make_map(&ma, A, M, K, n_net, M, (uint64_t)M * K, 64))) return rc;
  if ((rc = make_map(&mb, Bm, N, K, n_net, N, (uint64_t)N * K, 64))) return rc;
  return launch_tc_n<1>(bn, ma, mb, om, a, sm_count, st);
}

}  // namespace bnf
