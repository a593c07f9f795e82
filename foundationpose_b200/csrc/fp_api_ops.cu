// fp_api_ops.cu — C-ABI entry points that expose single operators (used by the parity tests to
// bisect the pipeline layer by layer; the product entry points live in fp_api.cu).
#include "../../include/fpose.h"
#include "fp_common.cuh"
#include "fp_attn.cuh"
#include "fp_gemm.cuh"

namespace fp {
const char* get_last_error();
int prof_collect(int kind, double* total_ms, double* total_work, int* launches);
}

extern "C" {

const char* fp_last_error(void) { return fp::get_last_error(); }

unsigned long long fp_launch_count(void) { return fp::g_launch_count; }

int fp_prof_enable(int on) {
  fp::g_prof_on = on != 0;
  return 0;
}

int fp_prof_collect(int kind, double* total_ms, double* total_work, int* launches) {
  if (!total_ms || !total_work || !launches) {
    fp::set_last_error("fp_prof_collect: null output");
    return -1;
  }
  return fp::prof_collect(kind, total_ms, total_work, launches);
}

int fp_op_attention(const void* qkv, void* out, int B, int impl, void* stream) {
  if (!qkv || !out) {
    fp::set_last_error("fp_op_attention: null argument");
    return -1;
  }
  fp::AttnParams ap;
  ap.qkv = reinterpret_cast<const __half*>(qkv);
  ap.ld = 1536;
  ap.q_off = 0;
  ap.k_off = 512;
  ap.v_off = 1024;
  ap.group_col_stride = 0;
  ap.n_groups = 1;
  ap.out = reinterpret_cast<__half*>(out);
  ap.ld_out = 512;
  ap.out_group_stride = 0;
  ap.B = B;
  ap.T = 400;
  ap.n_heads = 4;
  ap.scale = 0.08838834764831845f;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return impl == 0 ? fp::attn_legacy_launch(ap, st) : fp::attn_tc_launch(ap, st);
}

int fp_op_gemm_layer(const fp_gemm_layer_t* l, void* stream) {
  if (!l) {
    fp::set_last_error("fp_op_gemm_layer: null layer");
    return -1;
  }
  fp::GemmLayer L;
  L.kind = l->kind;
  L.n_img = l->n_img;
  L.Hin = l->Hin;
  L.Win = l->Win;
  L.Cin = l->Cin;
  L.Cout = l->Cout;
  L.in = l->in;
  L.w = l->w;
  L.bias = l->bias;
  L.res = l->res;
  L.res_ld = l->res_ld;
  L.out = l->out;
  L.out_ld = l->out_ld;
  L.out_split = l->out_split;
  L.post_add = l->post_add;
  L.relu = l->relu;
  return fp::gemm_layer_launch(L, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
