// C ABI of libbnf_sm100.so (see include/bnf.h): host bookkeeping (parameter
// layout in the reference's tree_leaves order), workspace carving and the
// orchestration of the kernels for forward / loglik+grad / MAP steps / VI step.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "bnf_device.cuh"
#include "bnf_kernels.h"
#include "bnf_prof.h"
#include "bnf_tc.h"

using namespace bnf;
typedef __nv_bfloat16 bf16;

static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CU(x)                                                                            \
  do {                                                                                   \
    cudaError_t e_ = (x);                                                                \
    if (e_ != cudaSuccess)                                                               \
      return fail(BNF_ERR_CUDA, "%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
#define CUK() CU(cudaGetLastError())

extern "C" int bnf_abi_version(void) { return BNF_ABI_VERSION; }
extern "C" const char* bnf_last_error(void) { return g_err; }

// -----------------------------------------------------------------------------
// plan: models.py:216-252 bookkeeping + Flax/tree_leaves parameter order
// -----------------------------------------------------------------------------
extern "C" int bnf_plan_create(const bnf_config_t* c, bnf_plan_t** out) {
  if (!c || !out) return fail(BNF_ERR_INVALID, "null argument");
  if (c->abi_version != BNF_ABI_VERSION) return fail(BNF_ERR_INVALID, "ABI version mismatch");
  const int D = c->input_dim, W = c->width, L = c->depth;
  if (D < 1 || D > kMaxD) return fail(BNF_ERR_INVALID, "input_dim must be in [1,%d]", kMaxD);
  if (W < 1) return fail(BNF_ERR_INVALID, "width must be positive");
  if (L < 1 || L > kMaxLayers) return fail(BNF_ERR_INVALID, "depth must be in [1,%d]", kMaxLayers);
  if (c->n_seasonal < 0 || c->n_seasonal > kMaxSeasonal)
    return fail(BNF_ERR_INVALID, "n_seasonal must be in [0,%d]", kMaxSeasonal);
  if (c->n_interactions < 0 || c->n_interactions > kMaxInter)
    return fail(BNF_ERR_INVALID, "n_interactions must be in [0,%d]", kMaxInter);
  if (c->likelihood < BNF_NORMAL || c->likelihood > BNF_ZINB)
    return fail(BNF_ERR_INVALID, "unknown likelihood %d", c->likelihood);
  if (!c->fourier_degrees || !c->input_scales) return fail(BNF_ERR_INVALID, "null config array");

  bnf_plan* p = new bnf_plan();
  DevModel& m = p->m;
  memset(&m, 0, sizeof(m));
  m.D = D; m.W = W; m.L = L; m.likelihood = c->likelihood;
  m.n_seasonal = c->n_seasonal; m.n_inter = c->n_interactions;
  for (int i = 0; i < D; ++i) {
    m.input_scales[i] = (float)c->input_scales[i];
    m.fourier_deg[i] = c->fourier_degrees[i];
    if (m.fourier_deg[i] > 24) { delete p; return fail(BNF_ERR_INVALID, "fourier degree > 24"); }
  }
  const float two_pi = (float)(2.0 * M_PI);  // f32(2*pi): models.py:73 evaluates (2*pi*f) in f32 first
  for (int k = 0; k < m.n_seasonal; ++k) {
    m.seasonal_w[k] = two_pi * c->seasonal_freq[k];
    m.seasonal_h[k] = c->seasonal_harm[k];
  }
  for (int j = 0; j < m.n_inter; ++j) {
    m.inter_a[j] = c->interactions[2 * j];
    m.inter_b[j] = c->interactions[2 * j + 1];
    if (m.inter_a[j] < 0 || m.inter_a[j] >= D || m.inter_b[j] < 0 || m.inter_b[j] >= D) {
      delete p;
      return fail(BNF_ERR_INVALID, "interaction index out of range");
    }
  }
  // feature groups: [scaled_x, fourier_i (deg>0, filtered BEFORE enumeration), seasonal,
  // interactions]; the name index counts empty groups too (models.py:242-251).
  std::map<std::string, std::pair<int, int>> shapes;  // name -> (rows, cols)
  int col = 0, gidx = 0, ngroups = 0;
  std::vector<std::pair<int*, std::string>> scale_refs;
  auto add_group = [&](int size, int* col_out, int* off_out) {
    if (size > 0) {
      *col_out = col;
      col += size;
      std::string nm = "feature_inv_sp_scale" + std::to_string(gidx);
      shapes[nm] = {0, 0};
      scale_refs.push_back({off_out, nm});
      ++ngroups;
    } else {
      *col_out = -1;
      *off_out = -1;
    }
    ++gidx;
  };
  add_group(D, &m.col_x, &m.off_scale_x);
  for (int i = 0; i < D; ++i) {
    if (m.fourier_deg[i] > 0) add_group(2 * m.fourier_deg[i], &m.fourier_col[i], &m.fourier_scale_off[i]);
    else { m.fourier_col[i] = -1; m.fourier_scale_off[i] = -1; }
  }
  add_group(2 * m.n_seasonal, &m.col_seasonal, &m.off_scale_seasonal);
  add_group(m.n_inter, &m.col_inter, &m.off_scale_inter);
  m.F = col;
  m.Fp = (m.F + kFeatPad - 1) / kFeatPad * kFeatPad;
  m.inv_sqrt_F = 1.0f / sqrtf((float)m.F);
  m.inv_sqrt_W = 1.0f / sqrtf((float)W);
  p->n_groups = ngroups;

  for (int l = 0; l <= L; ++l) {
    int fan = l == 0 ? m.F : W, outw = l == L ? 1 : W;
    shapes["Dense_" + std::to_string(l) + "/bias"] = {outw, 0};
    shapes["Dense_" + std::to_string(l) + "/kernel"] = {fan, outw};
    if (l < L) shapes["inv_sp_layer_scale" + std::to_string(l)] = {0, 0};
  }
  shapes["inv_sp_output_scale"] = {0, 0};
  shapes["log_scale_adjustment"] = {D, 0};
  shapes["logit_activation_weight"] = {0, 0};
  // jax.tree_util.tree_leaves on nested dicts: keys sorted at every level.
  std::vector<std::string> names;
  for (auto& kv : shapes) names.push_back(kv.first);
  auto key = [](const std::string& s) {
    size_t slash = s.find('/');
    return slash == std::string::npos ? std::make_pair(s, std::string())
                                      : std::make_pair(s.substr(0, slash), s.substr(slash + 1));
  };
  std::sort(names.begin(), names.end(),
            [&](const std::string& a, const std::string& b) { return key(a) < key(b); });
  int64_t off = 3;
  std::map<std::string, int64_t> offs;
  for (auto& nm : names) {
    auto sh = shapes[nm];
    int64_t n = sh.first == 0 ? 1 : (sh.second == 0 ? sh.first : (int64_t)sh.first * sh.second);
    p->leaves.push_back({nm, off, sh.first, sh.second});
    offs[nm] = off;
    off += n;
  }
  if (off > 0x7fffffffLL) { delete p; return fail(BNF_ERR_INVALID, "too many parameters"); }
  m.P = (int)off;
  for (auto& r : scale_refs) *r.first = (int)offs[r.second];
  for (int l = 0; l <= L; ++l) {
    m.off_bias[l] = (int)offs["Dense_" + std::to_string(l) + "/bias"];
    m.off_kernel[l] = (int)offs["Dense_" + std::to_string(l) + "/kernel"];
    if (l < L) m.off_layer_scale[l] = (int)offs["inv_sp_layer_scale" + std::to_string(l)];
  }
  m.off_out_scale = (int)offs["inv_sp_output_scale"];
  m.off_lsa = (int)offs["log_scale_adjustment"];
  m.off_actw = (int)offs["logit_activation_weight"];

  p->sm_count = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) == cudaSuccess) p->sm_count = prop.multiProcessorCount;
  }
  cudaGetLastError();  // no GPU is fine for bookkeeping
  *out = p;
  return BNF_OK;
}

static void free_graph_cache(bnf_plan* p);

extern "C" void bnf_plan_destroy(bnf_plan_t* p) {
  if (!p) return;
  if (p->graph_stream) cudaDeviceSynchronize();
  free_graph_cache(p);
  if (p->graph_stream) cudaStreamDestroy((cudaStream_t)p->graph_stream);
  delete p;
}

extern "C" int bnf_plan_info(const bnf_plan_t* p, bnf_plan_info_t* o) {
  if (!p || !o) return fail(BNF_ERR_INVALID, "null argument");
  o->num_params = p->m.P;
  o->num_features = p->m.F;
  o->padded_features = p->m.Fp;
  o->num_leaves = (int)p->leaves.size();
  o->num_feature_groups = p->n_groups;
  o->sm_count = p->sm_count;
  return BNF_OK;
}

extern "C" int bnf_plan_leaf(const bnf_plan_t* p, int32_t leaf, char* name, int32_t name_len,
                             int64_t* offset, int32_t* rows, int32_t* cols) {
  if (!p || leaf < 0 || leaf >= (int)p->leaves.size()) return fail(BNF_ERR_INVALID, "bad leaf index");
  const Leaf& l = p->leaves[leaf];
  if (name && name_len > 0) snprintf(name, name_len, "%s", l.name.c_str());
  if (offset) *offset = l.offset;
  if (rows) *rows = l.rows;
  if (cols) *cols = l.cols;
  return BNF_OK;
}

// -----------------------------------------------------------------------------
// workspace
// -----------------------------------------------------------------------------
namespace {
struct Carver {
  char* base; size_t off;
  explicit Carver(void* b) : base((char*)b), off(0) {}
  void* take(size_t bytes) {
    void* r = base ? base + off : nullptr;
    off += (bytes + 255) / 256 * 256;
    return r;
  }
};

struct Ws {
  float* derived;
  void* feat;
  void* z[kMaxLayers];
  void* h[kMaxLayers];
  float* opre; float* r;
  void* dU[2];
  float* dfeat;
  float* ll; float* prior;
  float* grad;             // MAP / VI internal gradient [n_net, P]
  float* vz; float* veps; float* vloss;  // VI
  int32_t* idxwin;         // device-drawn batch windows [n_net (MAP) or 1 (VI), B]
  bf16* wt; bf16* wn;      // bf16 weight staging for the tcgen05 path
  float* mm;               // small scratch
  size_t bytes;
};

Ws carve(const bnf_plan* p, int prec, int n_net, int B, int mode, void* base) {
  const DevModel& m = p->m;
  const bool x3 = prec == BNF_PREC_BF16X3;
  // bytes per element: f32 / bf16 / three bf16 planes (bf16x3 GEMM operands; its z stays f32)
  const size_t ts = prec == BNF_PREC_FP32 ? 4 : (x3 ? 6 : 2);
  const size_t tz = x3 ? 4 : ts;
  Carver c(base);
  Ws w;
  memset(&w, 0, sizeof(w));
  const size_t rows = (size_t)n_net * B;
  w.derived = (float*)c.take((size_t)n_net * kDerivedStride * 4);
  w.feat = c.take(rows * m.Fp * ts);
  const bool g = mode != BNF_WS_FORWARD;
  for (int l = 0; l < m.L; ++l) {
    if (g) {
      w.z[l] = c.take(rows * m.W * tz);
      w.h[l] = c.take(rows * m.W * ts);
    } else {
      w.z[l] = nullptr;
      w.h[l] = l < 2 ? c.take(rows * m.W * ts) : w.h[l - 2];  // ping-pong
    }
  }
  w.ll = (float*)c.take((size_t)n_net * 4);
  w.prior = (float*)c.take((size_t)n_net * 4);
  w.mm = (float*)c.take(64);
  if (g) {
    w.opre = (float*)c.take(rows * 4);
    w.r = (float*)c.take(rows * 4);
    const size_t tdu = x3 ? 4 : ts;               // bf16x3: the backpropagated dU carries two planes
    w.dU[0] = c.take(rows * m.W * tdu);
    w.dU[1] = m.L > 1 ? c.take(rows * m.W * tdu) : w.dU[0];
    w.dfeat = x3 ? nullptr : (float*)c.take(rows * m.Fp * 4);   // bf16x3: dfeat never leaves the SM
  }
  if (mode == BNF_WS_MAP || mode == BNF_WS_VI) w.grad = (float*)c.take((size_t)n_net * m.P * 4);
  if (mode == BNF_WS_MAP) w.idxwin = (int32_t*)c.take((size_t)n_net * B * 4);
  if (mode == BNF_WS_VI) w.idxwin = (int32_t*)c.take((size_t)B * 4);
  if (mode == BNF_WS_VI) {
    w.vz = (float*)c.take((size_t)n_net * m.P * 4);
    w.veps = (float*)c.take((size_t)n_net * m.P * 4);
    w.vloss = (float*)c.take((size_t)n_net * 4);
  }
  if (prec == BNF_PREC_BF16) {
    size_t per = tc_weight_elems(m);
    w.wt = (bf16*)c.take((size_t)n_net * per * 2);
    w.wn = (bf16*)c.take((size_t)n_net * per * 2);
  } else if (x3) {
    w.wn = (bf16*)c.take((size_t)n_net * tc_weight_elems(m) * 6);   // [layer][Kp][3*W]
  }
  w.bytes = c.off;
  return w;
}

int check_common(const bnf_plan* p, int prec, int n_net, int B) {
  if (!p) return fail(BNF_ERR_INVALID, "null plan");
  if (prec < 0 || prec > 3) return fail(BNF_ERR_INVALID, "unknown precision %d", prec);
  if (n_net < 1 || B < 1) return fail(BNF_ERR_INVALID, "n_networks and batch_rows must be positive");
  if (n_net > 65535) return fail(BNF_ERR_INVALID, "n_networks must be <= 65535");
  if ((double)n_net * B * std::max(p->m.W, p->m.Fp) > 9.0e18) return fail(BNF_ERR_INVALID, "problem too large");
  if (prec == BNF_PREC_BF16 || prec == BNF_PREC_BF16X3) {
    const char* why = tc_unsupported_reason(p->m);
    if (why) return fail(BNF_ERR_UNSUPPORTED, "tcgen05 path: %s", why);
  }
  if (prec == BNF_PREC_BF16X3) {
    // the split-operand mode exists only with its fused epilogues
    if (!head_fused_x3_supported(p->m)) return fail(BNF_ERR_UNSUPPORTED, "bf16x3 needs a width in {64, 128, 256, 512, 1024}");
    if (p->m.W > 1024) return fail(BNF_ERR_UNSUPPORTED, "bf16x3 needs width <= 1024");
    if (p->m.Fp > 128) return fail(BNF_ERR_UNSUPPORTED, "bf16x3 needs <= 128 encoded features");
  }
  return BNF_OK;
}

// the transposed bf16 kernel copy is only read by the BNF_FWD_WT=1 path
bool need_wt() { return tc_fwd_uses_wt(); }

// forward (+ optional backward) for n_net networks on B rows.
template <typename T>
int run_net(const bnf_plan* p, int prec, const float* params, int n_net, const float* x,
            const float* y, const int32_t* idx, int64_t idx_stride, int B, const Ws& w,
            float* out_loc, float* ll, float* grad, cudaStream_t st, bool prepped = false) {
  // prepped: the derived scalars and the bf16 weight copies are already current (the fused MAP
  // update of the previous step wrote them)
  const DevModel& m = p->m;
  if (prec == BNF_PREC_BF16X3) {
    // ---- split-operand tensor-core mode (f32-parity): feat / h / dU are rows of three bf16 planes,
    // z is f32; same kernel sequence as the bf16 mode without TC_FWD_HEAD
    if constexpr (sizeof(T) == 2) {
      const bool g = grad != nullptr;
      if (!prepped) {
        launch_prep(m, params, w.derived, n_net, nullptr, nullptr, nullptr, st);
        tc_cast_weights_x3(m, params, w.wn, n_net, st);
      }
      launch_encode_x3(m, w.derived, x, idx, idx_stride, B, (bf16*)w.feat, n_net, st);
      // training with W <= 256: the last hidden layer's GEMM also runs the head, the log-likelihood and
      // its own activation backward (TC_FWD_HEAD): z, h of that layer never reach HBM
      const bool head_epi = g && ll && tc_fwd_head_supported(m);
      for (int l = 0; l < m.L; ++l) {
        const bf16* a_in = l == 0 ? (const bf16*)w.feat : (const bf16*)w.h[l - 1];
        if (head_epi && l == m.L - 1) {
          int rc = tc_fwd_head(p, params, w.derived, a_in, w.wn, y, idx, idx_stride, (bf16*)w.dU[0], ll, grad, n_net, B, st, true);
          if (rc) return fail(rc, "tc_fwd_head (bf16x3) failed: %s", tc_last_error());
          continue;
        }
        int rc = tc_fwd_layer(p, l, params, w.derived, a_in, nullptr, w.wn, nullptr, (bf16*)w.h[l], n_net, B, st,
                              true, g ? (float*)w.z[l] : nullptr);
        if (rc) return fail(rc, "tc_fwd_layer (bf16x3) failed: %s", tc_last_error());
      }
      if (!(g && ll)) {
        launch_head_x3(m, params, w.derived, (const bf16*)w.h[m.L - 1], y, idx, idx_stride, B, out_loc,
                       ll ? w.opre : nullptr, ll ? w.r : nullptr, ll, nullptr, n_net, st);
        CUK();
        if (!g) return BNF_OK;
        return fail(BNF_ERR_INVALID, "gradient without log-likelihood output");
      }
      int cur = 0;
      if (!head_epi &&
          !launch_head_fused_x3(m, params, w.derived, (const bf16*)w.h[m.L - 1], (const float*)w.z[m.L - 1], y, idx,
                                idx_stride, B, (bf16*)w.dU[cur], ll, grad, n_net, st))
        return fail(BNF_ERR_UNSUPPORTED, "bf16x3 head kernel does not support this width");
      for (int l = m.L - 1; l >= 0; --l) {
        const bf16* a_in = l == 0 ? (const bf16*)w.feat : (const bf16*)w.h[l - 1];
        int rc = tc_wgrad(p, l, a_in, (const bf16*)w.dU[cur], grad, n_net, B, st, true, l == 0 && tc_bias0_via_wgrad(m));
        if (rc) return fail(rc, "tc_wgrad (bf16x3) failed: %s", tc_last_error());
        if (l > 0) {
          rc = tc_dgrad_act_x3(p, l, w.wn, (const bf16*)w.dU[cur], (bf16*)w.dU[cur ^ 1], (const float*)w.z[l - 1],
                               params, w.derived, grad, n_net, B, st);
          if (rc) return fail(rc, "tc_dgrad_act (bf16x3) failed: %s", tc_last_error());
          cur ^= 1;
        } else {
          rc = tc_dgrad0_enc(p, w.wn, (const bf16*)w.dU[cur], x, idx, idx_stride, params, w.derived, grad, n_net, B, st, true);
          if (rc) return fail(rc, "tc_dgrad0_enc (bf16x3) failed: %s", tc_last_error());
        }
      }
      CUK();
      return BNF_OK;
    } else {
      return fail(BNF_ERR_INVALID, "internal: bf16x3 runs on bf16 planes");
    }
  }
  const bool tc = prec == BNF_PREC_BF16;
  if (!prepped) launch_prep(m, params, w.derived, n_net, nullptr, nullptr, nullptr, st);
  // bf16 tensor-core mode: the feature encode fused into the Dense_0 GEMM (encoder warps generate the
  // A tiles in shared memory; `feat` never reaches HBM) is the forward-only default; training steps
  // need `feat` again for the Dense_0 wgrad and keep encode kernel + GEMM (tc_fused_encode_wanted).
  const bool fuse0 = tc && tc_fused_encode_wanted(m, grad != nullptr);
  if (!fuse0) launch_encode<T>(m, w.derived, x, idx, idx_stride, B, (T*)w.feat, n_net, st);
  if (tc && !prepped) tc_cast_weights(m, params, need_wt() ? w.wt : nullptr, w.wn, n_net, st);
  const bool g = grad != nullptr;
  // training in tensor-core mode with W <= 256: the last hidden layer's GEMM also runs the head,
  // the log-likelihood and its own activation backward (z, h of that layer never reach HBM)
  const bool head_epi = tc && g && ll && tc_fwd_head_supported(m) && !(fuse0 && m.L == 1);
  for (int l = 0; l < m.L; ++l) {
    const T* a_in = l == 0 ? (const T*)w.feat : (const T*)w.h[l - 1];
    if (head_epi && l == m.L - 1) {
      int rc = tc_fwd_head(p, params, w.derived, (const bf16*)a_in, w.wn, y, idx, idx_stride, (bf16*)w.dU[0], ll, grad, n_net, B, st);
      if (rc) return fail(rc, "tc_fwd_head failed: %s", tc_last_error());
      continue;
    }
    if (tc && l == 0 && fuse0) {
      int rc = tc_fwd_layer0_fused(p, params, w.derived, x, idx, idx_stride, w.wn, grad ? (bf16*)w.feat : nullptr,
                                   (bf16*)w.z[0], (bf16*)w.h[0], n_net, B, st);
      if (rc) return fail(rc, "tc_fwd_layer0_fused failed: %s", tc_last_error());
    } else if (tc) {
      int rc = tc_fwd_layer(p, l, params, w.derived, (const bf16*)a_in, w.wt, w.wn, (bf16*)w.z[l], (bf16*)w.h[l], n_net, B, st);
      if (rc) return fail(rc, "tc_fwd_layer failed: %s", tc_last_error());
    } else {
      launch_fwd_layer_simt_t<T>(m, l, params, w.derived, a_in, l == 0 ? m.F : m.W, l == 0 ? m.Fp : m.W,
                                 (T*)w.z[l], (T*)w.h[l], n_net, B, st);
    }
  }
  int cur = 0;
  // training: head + activation backward of the last layer in one kernel when the shape allows
  const char* nh = getenv("BNF_NO_FUSED_HEAD");
  bool head_done = head_epi;
  if (!head_done && g && ll && !(nh && nh[0] == '1'))
    head_done = launch_head_fused<T>(m, params, w.derived, (const T*)w.h[m.L - 1], (const T*)w.z[m.L - 1], y, idx,
                                     idx_stride, B, (T*)w.dU[cur], ll, grad, n_net, st);
  if (!head_done) {
    launch_head<T>(m, params, w.derived, (const T*)w.h[m.L - 1], y, idx, idx_stride, B, out_loc,
                   ll ? w.opre : nullptr, ll ? w.r : nullptr, ll, g ? grad : nullptr, n_net, st);
    CUK();
    if (!g) return BNF_OK;
    launch_act_bwd<T>(m, m.L - 1, true, params, w.derived, (const T*)w.z[m.L - 1], (const T*)w.h[m.L - 1],
                      w.r, (T*)w.dU[cur], B, grad, n_net, st);
  }
  for (int l = m.L - 1; l >= 0; --l) {
    const T* a_in = l == 0 ? (const T*)w.feat : (const T*)w.h[l - 1];
    const int Kin = l == 0 ? m.F : m.W, lda = l == 0 ? m.Fp : m.W;
    if (tc) {
      // layer 0: the constant-one feature column also yields the Dense_0 bias gradient when the
      // fused dgrad + activation backward (which then skips its column sums) produced dU_0
      const char* nfa = getenv("BNF_NO_FUSED_ACT_BWD");
      const bool bias0 = l == 0 && !(nfa && nfa[0] == '1') && tc_bias0_via_wgrad(m);
      int rc = tc_wgrad(p, l, (const bf16*)a_in, (const bf16*)w.dU[cur], grad, n_net, B, st, false, bias0);
      if (rc) return fail(rc, "tc_wgrad failed: %s", tc_last_error());
    } else {
      launch_wgrad_simt_t<T>(m, l, a_in, Kin, lda, (const T*)w.dU[cur], grad, n_net, B, st);
    }
    if (l > 0) {
      const char* nf = getenv("BNF_NO_FUSED_ACT_BWD");
      if (tc && !(nf && nf[0] == '1') && tc_dgrad_act_supported(m)) {
        // dgrad + activation backward of layer l-1 in one kernel (epilogue fusion)
        int rc = tc_dgrad(p, l, w.wn, (const bf16*)w.dU[cur], (bf16*)w.dU[cur ^ 1], nullptr, n_net, B, st,
                          (const bf16*)w.z[l - 1], params, w.derived, grad);
        if (rc) return fail(rc, "tc_dgrad (fused) failed: %s", tc_last_error());
      } else {
        if (tc) {
          int rc = tc_dgrad(p, l, w.wn, (const bf16*)w.dU[cur], (bf16*)w.dU[cur ^ 1], nullptr, n_net, B, st);
          if (rc) return fail(rc, "tc_dgrad failed: %s", tc_last_error());
        } else {
          launch_dgrad_simt_t<T, T>(m, l, params, (const T*)w.dU[cur], (T*)w.dU[cur ^ 1], m.W, m.W, n_net, B, st);
        }
        launch_act_bwd<T>(m, l - 1, false, params, w.derived, (const T*)w.z[l - 1], (const T*)nullptr,
                          nullptr, (T*)w.dU[cur ^ 1], B, grad, n_net, st);
      }
      cur ^= 1;
    } else {
      if (tc && tc_dgrad0_enc_supported(m)) {
        // dgrad_0 + feature-encode backward in one kernel: dfeat never goes to HBM
        int rc = tc_dgrad0_enc(p, w.wn, (const bf16*)w.dU[cur], x, idx, idx_stride, params, w.derived, grad, n_net, B, st);
        if (rc) return fail(rc, "tc_dgrad0_enc failed: %s", tc_last_error());
        continue;
      }
      if (tc) {
        int rc = tc_dgrad(p, 0, w.wn, (const bf16*)w.dU[cur], nullptr, w.dfeat, n_net, B, st);
        if (rc) return fail(rc, "tc_dgrad failed: %s", tc_last_error());
      } else {
        launch_dgrad_simt_t<T, float>(m, 0, params, (const T*)w.dU[cur], w.dfeat, m.F, m.Fp, n_net, B, st);
      }
      launch_encode_bwd(m, params, w.derived, x, idx, idx_stride, B, w.dfeat, grad, n_net, tc, /*col_major=*/tc, st);
    }
  }
  CUK();
  return BNF_OK;
}

int run_net_any(const bnf_plan* p, int prec, const float* params, int n_net, const float* x,
                const float* y, const int32_t* idx, int64_t idx_stride, int B, const Ws& w,
                float* out_loc, float* ll, float* grad, cudaStream_t st, bool prepped = false) {
  if (prec == BNF_PREC_FP32)
    return run_net<float>(p, prec, params, n_net, x, y, idx, idx_stride, B, w, out_loc, ll, grad, st, prepped);
  return run_net<bf16>(p, prec, params, n_net, x, y, idx, idx_stride, B, w, out_loc, ll, grad, st, prepped);
}
}  // namespace

extern "C" int bnf_precision_supported(const bnf_plan_t* p, int32_t prec) {
  return check_common(p, prec, 1, 1);
}

extern "C" size_t bnf_workspace_bytes(const bnf_plan_t* p, int32_t prec, int32_t n_net, int32_t B,
                                      int32_t mode) {
  if (!p || n_net < 1 || B < 1) return 0;
  return carve(p, prec, n_net, B, mode, nullptr).bytes;
}

extern "C" int bnf_forward(const bnf_plan_t* p, int32_t prec, const float* params, int32_t n_net,
                           const float* x, const int32_t* idx, int64_t idx_stride, int32_t B,
                           float* out_loc, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_common(p, prec, n_net, B);
  if (rc) return rc;
  if (!params || !x || !out_loc || !ws) return fail(BNF_ERR_INVALID, "null pointer");
  Ws w = carve(p, prec, n_net, B, BNF_WS_FORWARD, ws);
  if (w.bytes > ws_bytes) return fail(BNF_ERR_WORKSPACE, "workspace too small: need %zu, have %zu", w.bytes, ws_bytes);
  PdlScope pdl(true);
  return run_net_any(p, prec, params, n_net, x, nullptr, idx, idx_stride, B, w, out_loc, nullptr,
                     nullptr, (cudaStream_t)stream);
}

extern "C" int bnf_loglik_grad(const bnf_plan_t* p, int32_t prec, const float* params, int32_t n_net,
                               const float* x, const float* y, const int32_t* idx, int64_t idx_stride,
                               int32_t B, float* out_ll, float* out_grad, void* ws, size_t ws_bytes,
                               void* stream) {
  int rc = check_common(p, prec, n_net, B);
  if (rc) return rc;
  if (!params || !x || !y || !out_ll || !ws) return fail(BNF_ERR_INVALID, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  Ws w = carve(p, prec, n_net, B, BNF_WS_GRAD, ws);
  if (w.bytes > ws_bytes) return fail(BNF_ERR_WORKSPACE, "workspace too small: need %zu, have %zu", w.bytes, ws_bytes);
  CU(cudaMemsetAsync(out_ll, 0, (size_t)n_net * 4, st));
  float* grad = out_grad;
  if (grad) CU(cudaMemsetAsync(grad, 0, (size_t)n_net * p->m.P * 4, st));
  // value-only still needs r/opre scratch for the head kernel
  PdlScope pdl(true);
  return run_net_any(p, prec, params, n_net, x, y, idx, idx_stride, B, w, nullptr, out_ll, grad, st);
}

// ---- CUDA-graph replay of one training step -------------------------------------------------
// Signature of a captured step: every pointer / scalar baked into the graph's kernel nodes.
// (The loss buffer is NOT baked: the kernels that write a step's loss read its address from the
// workspace, where the prologue of every call stores it -- callers may pass a fresh buffer per
// call and still replay.)
struct StepGraphKey {
  const void* p0; const void* p1; const void* p2; const void* p3; const void* step_count; const void* x;
  const void* y; const void* idx; const void* ws;
  long long idx_stride, first_member; unsigned long long seed;
  int kind, prec, n_net, aux, B, n_total; float lr, w; int flags, shuffle;
  bool operator==(const StepGraphKey& o) const { return memcmp(this, &o, sizeof(*this)) == 0; }
};

static void free_graph_cache(bnf_plan* p) {
  for (int k = 0; k < 2 * bnf_plan::kGraphWays; ++k) {
    if (p->graph_exec[k]) cudaGraphExecDestroy((cudaGraphExec_t)p->graph_exec[k]);
    if (p->graph_exec_multi[k]) cudaGraphExecDestroy((cudaGraphExec_t)p->graph_exec_multi[k]);
    delete (StepGraphKey*)p->graph_key[k];
    p->graph_exec[k] = p->graph_exec_multi[k] = p->graph_key[k] = nullptr;
  }
  for (int k = 0; k < 2; ++k) {
    delete (StepGraphKey*)p->last_key[k];
    p->last_key[k] = nullptr;
  }
}
static bool pdl_scope_would_enable() {
  const char* e = getenv("BNF_PDL");
  return !(e && e[0] == '0');
}

// Steps per launch of the unrolled step graph (BNF_GRAPH_UNROLL; 1 = one graph launch per step).
static int graph_unroll() {
  const char* e = getenv("BNF_GRAPH_UNROLL");
  int u = e ? atoi(e) : 8;
  return u < 1 ? 1 : (u > 64 ? 64 : u);
}

// Captures `reps` consecutive steps on the plan's capture stream and instantiates the graph.
template <typename F>
static int capture_steps(const bnf_plan* p, int reps, F& one_step, cudaGraphExec_t* out, long long* launches_per_step) {
  if (!p->graph_stream) {
    cudaStream_t gs;
    CU(cudaStreamCreateWithFlags(&gs, cudaStreamNonBlocking));
    p->graph_stream = gs;
  }
  cudaStream_t gs = (cudaStream_t)p->graph_stream;
  cudaGraph_t graph = nullptr;
  CU(cudaStreamBeginCapture(gs, cudaStreamCaptureModeThreadLocal));
  const unsigned long long before = bnf_debug_launch_count();
  int rc = 0;
  for (int r = 0; r < reps && !rc; ++r) rc = one_step(gs);
  const long long captured = (long long)(bnf_debug_launch_count() - before);
  cudaError_t ce = cudaStreamEndCapture(gs, &graph);
  prof_add_launches(-captured);                                   // captured, not launched
  if (rc || ce != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    if (rc) return rc;
    return fail(BNF_ERR_CUDA, "CUDA graph capture of the training step failed (%s)", cudaGetErrorString(ce));
  }
  cudaGraphExec_t exec = nullptr;
  ce = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ce != cudaSuccess) return fail(BNF_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ce));
  *out = exec;
  if (launches_per_step) *launches_per_step = captured / reps;
  return BNF_OK;
}

// Replays `one_step` n_steps times on `st`: as a cached CUDA graph (slot `kind`: 0 MAP, 1 VI) when
// its arguments are those of the cached capture or the call is long enough to pay for a capture,
// else as direct launches.  one_step(stream) must enqueue a step whose launch arguments do not
// depend on the step number.
template <typename F>
static int replay_steps(const bnf_plan* p, int kind, const StepGraphKey& key, int n_steps, cudaStream_t st, F one_step) {
  bool use_graph = !prof_enabled() && !getenv("BNF_NO_GRAPH");
  int way = -1;                     // cache entry of this call's graph
  if (use_graph) {
    const int base = kind * bnf_plan::kGraphWays;
    for (int w = 0; w < bnf_plan::kGraphWays; ++w) {
      const StepGraphKey* c = (const StepGraphKey*)p->graph_key[base + w];
      if (p->graph_exec[base + w] && c && *c == key) way = base + w;
    }
    StepGraphKey* last = (StepGraphKey*)p->last_key[kind];
    const bool hit = way >= 0;
    const bool seen = last && *last == key;
    if (!last) { last = new StepGraphKey(); p->last_key[kind] = last; }
    *last = key;
    if (!hit && n_steps < 4 && !seen) use_graph = false;     // not worth a capture yet
    if (use_graph && !hit) {
      way = base;                                            // an empty way, else the least recently used
      for (int w = 0; w < bnf_plan::kGraphWays; ++w) {
        if (!p->graph_exec[base + w]) { way = base + w; break; }
        if (p->graph_age[base + w] < p->graph_age[way]) way = base + w;
      }
      if (p->graph_exec[way]) {
        // the old graphs may still be running on the caller's stream
        CU(cudaStreamSynchronize(st));
        cudaGraphExecDestroy((cudaGraphExec_t)p->graph_exec[way]);
        if (p->graph_exec_multi[way]) cudaGraphExecDestroy((cudaGraphExec_t)p->graph_exec_multi[way]);
        p->graph_exec[way] = p->graph_exec_multi[way] = nullptr;
      }
      cudaGraphExec_t exec = nullptr;
      const int rc = capture_steps(p, 1, one_step, &exec, &p->graph_launches[way]);
      if (rc) return rc;
      p->graph_exec[way] = exec;
      if (!p->graph_key[way]) p->graph_key[way] = new StepGraphKey();
      *(StepGraphKey*)p->graph_key[way] = key;
    }
    // long calls: the step captured `unroll` times back to back, one launch per `unroll` steps
    // (no graph-launch boundary and an unbroken programmatic-dependent-launch chain between the
    // fused update of one step and the first kernel of the next)
    const int unroll = graph_unroll();
    if (use_graph && unroll > 1 && n_steps >= 2 * unroll &&
        (!p->graph_exec_multi[way] || p->graph_multi_steps[way] != unroll)) {
      if (p->graph_exec_multi[way]) {
        CU(cudaStreamSynchronize(st));
        cudaGraphExecDestroy((cudaGraphExec_t)p->graph_exec_multi[way]);
        p->graph_exec_multi[way] = nullptr;
      }
      cudaGraphExec_t exec = nullptr;
      const int rc = capture_steps(p, unroll, one_step, &exec, nullptr);
      if (rc) return rc;
      p->graph_exec_multi[way] = exec;
      p->graph_multi_steps[way] = unroll;
    }
  }
  if (use_graph) {
    p->graph_age[way] = ++p->graph_clock;
    int s = 0;
    const int ms = p->graph_exec_multi[way] ? p->graph_multi_steps[way] : 0;
    if (ms > 1 && ms == graph_unroll())
      for (; s + ms <= n_steps; s += ms) CU(cudaGraphLaunch((cudaGraphExec_t)p->graph_exec_multi[way], st));
    for (; s < n_steps; ++s) CU(cudaGraphLaunch((cudaGraphExec_t)p->graph_exec[way], st));
    prof_add_launches(p->graph_launches[way] * n_steps);
    return BNF_OK;
  }
  for (int s = 0; s < n_steps; ++s) {
    int rc = one_step(st);
    if (rc) return rc;
  }
  CUK();
  return BNF_OK;
}

// shuffle: the batch windows are drawn on the device (idx must be NULL); else idx / idx == NULL as
// documented for bnf_map_steps
static int map_steps_impl(const bnf_plan_t* p, int32_t prec, float* params, float* am, float* av,
                          int32_t* step_count, int32_t n_net, const float* x, const float* y,
                          const int32_t* idx, int64_t idx_stride, int32_t B, int32_t n_total,
                          int32_t n_steps, float lr, float prior_weight, bool shuffle, uint64_t seed,
                          int64_t first_member, float* out_loss, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_common(p, prec, n_net, B);
  if (rc) return rc;
  if (!params || !am || !av || !step_count || !x || !y || !out_loss || !ws)
    return fail(BNF_ERR_INVALID, "null pointer");
  if (n_steps < 1 || n_total < B) return fail(BNF_ERR_INVALID, "bad n_steps / n_rows_total");
  cudaStream_t st = (cudaStream_t)stream;
  Ws w = carve(p, prec, n_net, B, BNF_WS_MAP, ws);
  if (w.bytes > ws_bytes) return fail(BNF_ERR_WORKSPACE, "workspace too small: need %zu, have %zu", w.bytes, ws_bytes);
  const DevModel& m = p->m;
  const float c_ll = (float)((double)n_total / (double)B);  // target.shape[0] / batch_size
  int32_t* slot = (int32_t*)(w.mm + 8);                      // loss-row cursor
  unsigned int* counter = (unsigned int*)(w.mm + 9);         // map_update's block ticket
  float** loss_slot = (float**)(w.mm + 10);                  // device copy of `out_loss` (8-byte aligned)
  const bool x3 = prec == BNF_PREC_BF16X3;
  const bool tc = prec == BNF_PREC_BF16;
  const int spe = n_total / B;                               // steps per epoch (ragged tail dropped)
  if (shuffle) { idx = w.idxwin; idx_stride = B; }
  // Paths that read the transposed weight copy (fused-encode experiment, BNF_FWD_WT=1) keep the
  // round-1 step (prep + cast every step): the fused update only maintains the natural copy.
  const bool legacy = (tc && need_wt()) || getenv("BNF_LEGACY_STEP");

  if (legacy) {
    CU(cudaMemsetAsync(slot, 0, 4, st));
    for (int s = 0; s < n_steps; ++s) {
      const int32_t* idx_s = (idx && !shuffle) ? idx + (size_t)s * B : idx;
      CU(cudaMemsetAsync(w.grad, 0, (size_t)n_net * m.P * 4, st));
      CU(cudaMemsetAsync(w.ll, 0, (size_t)n_net * 4, st));
      CU(cudaMemsetAsync(w.prior, 0, (size_t)n_net * 4, st));
      if (shuffle) launch_batch_window(n_total, B, spe, n_net, seed, first_member, step_count, w.idxwin, st);
      launch_tick(step_count, slot, st);
      rc = run_net_any(p, prec, params, n_net, x, y, idx_s, idx_stride, B, w, nullptr, w.ll, w.grad, st);
      if (rc) return rc;
      launch_map_adam(m.P, params, am, av, w.grad, step_count, c_ll, prior_weight, lr, w.prior, n_net, st);
      launch_map_loss(n_net, w.ll, w.prior, c_ll, prior_weight, out_loss, slot, st);
    }
    CUK();
    return BNF_OK;
  }

  // ---- prologue (once per call): zeroed accumulators, derived scalars, bf16 weight copies ----
  launch_prep(m, params, w.derived, n_net, w.ll, w.prior, slot /* + counter */, st, loss_slot, out_loss, w.grad);
  if (tc) tc_cast_weights(m, params, need_wt() ? w.wt : nullptr, w.wn, n_net, st);
  if (x3) tc_cast_weights_x3(m, params, w.wn, n_net, st);
  CUK();

  // One training step: [batch window] -> encode -> fwd GEMMs -> head -> bwd GEMMs -> encode_bwd ->
  // fused update.  With idx == NULL (full batch) or device-drawn windows the launch arguments do
  // not depend on the step (device-side cursors), so the sequence replays as one CUDA graph;
  // consecutive kernels are chained by programmatic dependent launch.
  auto one_step = [&](cudaStream_t s, const int32_t* idx_s) -> int {
    PdlScope pdl(true);
    if (shuffle) launch_batch_window(n_total, B, spe, n_net, seed, first_member, step_count, w.idxwin, s);
    int r = run_net_any(p, prec, params, n_net, x, y, idx_s, idx_stride, B, w, nullptr, w.ll, w.grad, s, /*prepped=*/true);
    if (r) return r;
    launch_map_update(m, params, am, av, w.grad, step_count, c_ll, prior_weight, lr, w.prior, w.ll, loss_slot,
                      slot, counter, w.derived, (tc || x3) ? w.wn : nullptr, (tc || x3) ? tc_weight_elems(m) : 0,
                      x3 ? 3 : 1, n_net, s);
    return BNF_OK;
  };

  if (idx == nullptr || shuffle) {
    StepGraphKey key;
    memset(&key, 0, sizeof(key));
    key.p0 = params; key.p1 = am; key.p2 = av; key.step_count = step_count; key.x = x; key.y = y;
    key.idx = idx; key.idx_stride = idx_stride; key.ws = ws; key.kind = 0; key.prec = prec; key.n_net = n_net;
    key.B = B; key.n_total = n_total; key.lr = lr; key.w = prior_weight;
    key.flags = (pdl_scope_would_enable() ? 1 : 0) | (need_wt() ? 2 : 0);
    key.shuffle = shuffle ? 1 : 0; key.seed = shuffle ? seed : 0; key.first_member = shuffle ? first_member : 0;
    return replay_steps(p, 0, key, n_steps, st, [&](cudaStream_t s) { return one_step(s, idx); });
  }
  for (int s = 0; s < n_steps; ++s) {
    // injected index rows: step s uses rows [s*B, (s+1)*B) of every network's index row
    rc = one_step(st, idx + (size_t)s * B);
    if (rc) return rc;
  }
  CUK();
  return BNF_OK;
}

extern "C" int bnf_map_steps(const bnf_plan_t* p, int32_t prec, float* params, float* am, float* av,
                             int32_t* step_count, int32_t n_net, const float* x, const float* y,
                             const int32_t* idx, int64_t idx_stride, int32_t B, int32_t n_total,
                             int32_t n_steps, float lr, float prior_weight, float* out_loss, void* ws,
                             size_t ws_bytes, void* stream) {
  return map_steps_impl(p, prec, params, am, av, step_count, n_net, x, y, idx, idx_stride, B, n_total, n_steps, lr,
                        prior_weight, false, 0, 0, out_loss, ws, ws_bytes, stream);
}

extern "C" int bnf_map_epochs(const bnf_plan_t* p, int32_t prec, float* params, float* am, float* av,
                              int32_t* step_count, int32_t n_net, const float* x, const float* y, int32_t B,
                              int32_t n_total, int32_t n_epochs, float lr, float prior_weight, uint64_t seed,
                              int64_t first_member, float* out_loss, void* ws, size_t ws_bytes, void* stream) {
  if (B < 1 || n_total < B || n_epochs < 1) return fail(BNF_ERR_INVALID, "bad batch_rows / n_rows_total / n_epochs");
  const long long steps = (long long)n_epochs * (n_total / B);
  if (steps > 0x7fffffffLL) return fail(BNF_ERR_INVALID, "too many steps");
  return map_steps_impl(p, prec, params, am, av, step_count, n_net, x, y, nullptr, 0, B, n_total, (int)steps, lr,
                        prior_weight, true, seed, first_member, out_loss, ws, ws_bytes, stream);
}

// One or more steps of the VI optimiser.  eps / idx injected (tests) only for single steps.
static int vi_steps_impl(const bnf_plan_t* p, int32_t prec, float* mu, float* rho, float* am, float* av,
                         int32_t* step_count, int32_t E, int32_t S, const float* eps, uint64_t seed,
                         const float* x, const float* y, const int32_t* idx, bool shuffle, int64_t device_id,
                         int32_t B, int32_t n_total, int32_t n_steps, float lr, float kl_weight,
                         float* out_loss, void* ws, size_t ws_bytes, void* stream) {
  if (E < 1 || S < 1) return fail(BNF_ERR_INVALID, "members and samples must be positive");
  if ((long long)E * S > 65535) return fail(BNF_ERR_INVALID, "members x samples must be <= 65535");
  const int n_net = E * S;
  int rc = check_common(p, prec, n_net, B);
  if (rc) return rc;
  if (!mu || !rho || !am || !av || !step_count || !x || !y || !out_loss || !ws)
    return fail(BNF_ERR_INVALID, "null pointer");
  if (!(kl_weight > 0.f)) return fail(BNF_ERR_INVALID, "kl_weight must be positive");
  if (n_steps < 1 || n_total < B) return fail(BNF_ERR_INVALID, "bad n_steps / n_rows_total");
  cudaStream_t st = (cudaStream_t)stream;
  Ws w = carve(p, prec, n_net, B, BNF_WS_VI, ws);
  if (w.bytes > ws_bytes) return fail(BNF_ERR_WORKSPACE, "workspace too small: need %zu, have %zu", w.bytes, ws_bytes);
  const DevModel& m = p->m;
  const float c = (float)((double)n_total / (double)B) / kl_weight;
  int32_t* slot = (int32_t*)(w.mm + 8);                      // loss-row cursor
  float** loss_slot = (float**)(w.mm + 10);                  // device copy of `out_loss`
  if (shuffle) idx = w.idxwin;
  // prologue: loss cursor = 0, loss buffer address -> workspace
  launch_arm_loss(loss_slot, out_loss, slot, st);
  auto one_step = [&](cudaStream_t s) -> int {
    // one shared random sub-batch per device per step (inference.py:704-709): the first B entries of
    // a fresh permutation keyed by (seed; device, number of completed steps)
    if (shuffle) launch_batch_window(n_total, B, 1, 1, seed ^ 0x5649424154434855ULL, device_id, step_count, w.idxwin, s);
    launch_tick(step_count, slot, s);
    // device draws are keyed by (seed, step count): every step has its own Philox stream
    launch_vi_sample(m.P, E, S, mu, rho, eps, w.veps, seed, 0x5649ULL, step_count, w.vz, s);
    CU(cudaMemsetAsync(w.grad, 0, (size_t)n_net * m.P * 4, s));
    CU(cudaMemsetAsync(w.ll, 0, (size_t)n_net * 4, s));
    CU(cudaMemsetAsync(w.vloss, 0, (size_t)n_net * 4, s));
    int r;
    {
      PdlScope pdl(true);
      r = run_net_any(p, prec, w.vz, n_net, x, y, idx, 0, B, w, nullptr, w.ll, w.grad, s);
    }
    if (r) return r;
    launch_vi_adam(m.P, E, S, mu, rho, am, av, w.vz, w.veps, w.grad, step_count, c, lr, w.vloss, s);
    launch_vi_loss(E, S, w.vloss, w.ll, c, nullptr, loss_slot, slot, s);
    return BNF_OK;
  };
  if (eps == nullptr && (idx == nullptr || shuffle)) {
    StepGraphKey key;
    memset(&key, 0, sizeof(key));
    key.p0 = mu; key.p1 = rho; key.p2 = am; key.p3 = av; key.step_count = step_count; key.x = x; key.y = y;
    key.idx = idx; key.ws = ws; key.kind = 1; key.prec = prec; key.n_net = E; key.aux = S; key.B = B;
    key.n_total = n_total; key.lr = lr; key.w = kl_weight; key.flags = pdl_scope_would_enable() ? 1 : 0;
    key.shuffle = shuffle ? 1 : 0; key.seed = seed; key.first_member = device_id;
    return replay_steps(p, 1, key, n_steps, st, one_step);
  }
  if (n_steps != 1) return fail(BNF_ERR_INVALID, "injected eps / index rows are single-step hooks");
  rc = one_step(st);
  if (rc) return rc;
  CUK();
  return BNF_OK;
}

extern "C" int bnf_vi_step(const bnf_plan_t* p, int32_t prec, float* mu, float* rho, float* am,
                           float* av, int32_t* step_count, int32_t E, int32_t S, const float* eps,
                           uint64_t seed, const float* x, const float* y, const int32_t* idx, int32_t B,
                           int32_t n_total, float lr, float kl_weight, float* out_loss, void* ws,
                           size_t ws_bytes, void* stream) {
  return vi_steps_impl(p, prec, mu, rho, am, av, step_count, E, S, eps, seed, x, y, idx, false, 0, B, n_total, 1, lr,
                       kl_weight, out_loss, ws, ws_bytes, stream);
}

extern "C" int bnf_vi_steps(const bnf_plan_t* p, int32_t prec, float* mu, float* rho, float* am,
                            float* av, int32_t* step_count, int32_t E, int32_t S, uint64_t seed,
                            int64_t device_id, const float* x, const float* y, int32_t B, int32_t n_total,
                            int32_t n_steps, float lr, float kl_weight, float* out_loss, void* ws,
                            size_t ws_bytes, void* stream) {
  return vi_steps_impl(p, prec, mu, rho, am, av, step_count, E, S, nullptr, seed, x, y, nullptr, B < n_total, device_id,
                       B, n_total, n_steps, lr, kl_weight, out_loss, ws, ws_bytes, stream);
}

// permute_dataset (inference.py:35-39) as the device draws it: out[i] = row at position i of the
// order of (seed; member, epoch).  Evaluated on the HOST (tests, replaying a device run).
extern "C" int bnf_debug_permutation(uint64_t seed, int64_t member, int32_t epoch, int32_t n, int32_t* out) {
  if (!out || n < 1 || epoch < 0) return fail(BNF_ERR_INVALID, "bad argument");
  const bnf::PermKeys pk = bnf::perm_keys(seed, (uint64_t)member, (uint32_t)epoch);
  const uint32_t hb = bnf::perm_half_bits((uint32_t)n);
  for (int32_t i = 0; i < n; ++i) out[i] = (int32_t)bnf::perm_index((uint32_t)i, (uint32_t)n, hb, pk);
  return BNF_OK;
}

extern "C" int bnf_vi_sample(const bnf_plan_t* p, const float* mu, const float* rho, int32_t E,
                             int32_t n_samples, const float* eps, uint64_t seed, float* out,
                             void* stream) {
  if (!p || !mu || !rho || !out || E < 1 || n_samples < 1) return fail(BNF_ERR_INVALID, "bad argument");
  launch_vi_sample(p->m.P, E, n_samples, mu, rho, eps, nullptr, seed, 0x504fULL, nullptr, out, (cudaStream_t)stream);
  CUK();
  return BNF_OK;
}

extern "C" int bnf_init_params(const bnf_plan_t* p, float lns_init, uint64_t seed, int64_t first_member,
                               int32_t n_net, float* out, void* stream) {
  if (!p || !out || n_net < 1) return fail(BNF_ERR_INVALID, "bad argument");
  launch_init_params(p->m, lns_init, seed, first_member, n_net, out, (cudaStream_t)stream);
  CUK();
  return BNF_OK;
}

extern "C" size_t bnf_quantile_workspace_bytes(int32_t, int32_t) { return 256; }

// inverse normal CDF (host, double): Acklam's rational approximation + one Halley step
static double ndtri(double pq) {
  static const double a[] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                             1.383577518672690e+02, -3.066479806614716e+01, 2.506628277459239e+00};
  static const double b[] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                             6.680131188771972e+01, -1.328068155288572e+01};
  static const double c[] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                             -2.549732539343734e+00, 4.374664141464968e+00, 2.938163982698783e+00};
  static const double d[] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00,
                             3.754408661907416e+00};
  double x;
  if (pq < 0.02425) {
    double q = sqrt(-2 * log(pq));
    x = (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
        ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
  } else if (pq <= 1 - 0.02425) {
    double q = pq - 0.5, r = q * q;
    x = (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q /
        (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1);
  } else {
    double q = sqrt(-2 * log(1 - pq));
    x = -(((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
        ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
  }
  double e = 0.5 * erfc(-x / sqrt(2.0)) - pq;
  double u = e * sqrt(2 * M_PI) * exp(x * x / 2);
  return x - u / (1 + x * u / 2);
}

extern "C" int bnf_mixture_quantiles(const float* means, const float* scales, int32_t M, int32_t N,
                                     const double* q, int32_t nq, int32_t approximate, float* out,
                                     void* ws, size_t ws_bytes, void* stream) {
  if (!means || !scales || !q || !out || M < 1 || N < 1 || nq < 1) return fail(BNF_ERR_INVALID, "bad argument");
  if (!ws || ws_bytes < 64) return fail(BNF_ERR_WORKSPACE, "workspace too small");
  std::vector<float> nd(nq);
  for (int i = 0; i < nq; ++i) {
    if (!(q[i] > 0.0 && q[i] < 1.0)) return fail(BNF_ERR_INVALID, "quantile must be in (0,1)");
    nd[i] = (float)ndtri(q[i]);
  }
  launch_quantiles(means, scales, M, N, q, nq, approximate != 0, nd.data(), out, (float*)ws, (cudaStream_t)stream);
  CUK();
  return BNF_OK;
}

// host-side evaluation of the device RNG's block function (known-answer test without a GPU)
extern "C" void bnf_debug_philox(const uint32_t* counter4, const uint32_t* key2, uint32_t* out4) {
  const uint4 r = bnf::philox4x32_10(make_uint4(counter4[0], counter4[1], counter4[2], counter4[3]),
                                     make_uint2(key2[0], key2[1]));
  out4[0] = r.x; out4[1] = r.y; out4[2] = r.z; out4[3] = r.w;
}

extern "C" int bnf_debug_gemm(int32_t mn_major, const void* a, const void* b, float* c, int32_t n_net,
                              int32_t m, int32_t n, int32_t k, void* stream) {
  if (!a || !b || !c || n_net < 1 || m < 1 || n < 1 || k < 1) return fail(BNF_ERR_INVALID, "bad argument");
  int dev = 0, sm = 148;
  CU(cudaGetDevice(&dev));
  CU(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev));
  int rc = tc_debug_gemm(mn_major, (const bf16*)a, (const bf16*)b, c, n_net, m, n, k, sm, (cudaStream_t)stream);
  if (rc) return fail(rc, "tc_debug_gemm: %s", tc_last_error());
  return BNF_OK;
}

extern "C" int bnf_nb_mixture_quantiles(const float* loc, const float* shape_raw, const float* pi_logit,
                                        int32_t M, int32_t N, const double* q, int32_t nq, float* out_means,
                                        float* out_q, void* ws, size_t ws_bytes, void* stream) {
  if (!loc || !shape_raw || !q || !out_means || !out_q || M < 1 || N < 1 || nq < 1)
    return fail(BNF_ERR_INVALID, "bad argument");
  if (!ws || ws_bytes < 64) return fail(BNF_ERR_WORKSPACE, "workspace too small");
  for (int i = 0; i < nq; ++i)
    if (!(q[i] > 0.0 && q[i] < 1.0)) return fail(BNF_ERR_INVALID, "quantile must be in (0,1)");
  launch_nb_quantiles(loc, shape_raw, pi_logit, M, N, q, nq, out_means, out_q, (float*)ws, (cudaStream_t)stream);
  CUK();
  return BNF_OK;
}
