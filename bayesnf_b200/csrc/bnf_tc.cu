// placeholder until the tcgen05 kernels land
#include "bnf_tc.h"
namespace bnf {
const char* tc_unsupported_reason(const DevModel&) { return "not built yet"; }
const char* tc_last_error() { return ""; }
size_t tc_weight_elems(const DevModel&) { return 0; }
void tc_cast_weights(const DevModel&, const float*, __nv_bfloat16*, __nv_bfloat16*, int, cudaStream_t) {}
int tc_fwd_layer(const bnf_plan*, int, const float*, const float*, const __nv_bfloat16*, const __nv_bfloat16*, __nv_bfloat16*, __nv_bfloat16*, int, int, cudaStream_t) { return BNF_ERR_UNSUPPORTED; }
int tc_dgrad(const bnf_plan*, int, const __nv_bfloat16*, const __nv_bfloat16*, __nv_bfloat16*, float*, int, int, cudaStream_t) { return BNF_ERR_UNSUPPORTED; }
int tc_wgrad(const bnf_plan*, int, const __nv_bfloat16*, const __nv_bfloat16*, float*, int, int, cudaStream_t) { return BNF_ERR_UNSUPPORTED; }
}
