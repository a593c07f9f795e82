// fp_attn.cuh — launchers of the non-GEMM transformer / scorer-tail / pose-update kernels.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stddef.h>

namespace fp {

struct AttnParams {
  const __half* qkv;  // [B*T][ld]; q at q_off + h*128, k at k_off + h*128, v at v_off + h*128
  int ld;
  int q_off, k_off, v_off;
  int group_col_stride;  // groups (e.g. trans / rot head) are further column blocks of the same rows
  int n_groups;
  __half* out;  // [group][B*T][ld_out], head h at column h*128
  int ld_out;
  size_t out_group_stride;
  int B, T, n_heads;
  float scale;  // 1/sqrt(head_dim)
};
// the tcgen05 kernel (fp_attn_tc.cu)
int attn_core_launch(const AttnParams& p, cudaStream_t stream);
int attn_tc_launch(const AttnParams& p, cudaStream_t stream);

int layernorm_launch(const __half* x, __half* y, const float* gamma, const float* beta, int rows, cudaStream_t stream);
int head_final_launch(const __half* x, const float* gamma, const float* beta, const float* w, const float* bias,
                      float* out, int B, int T, int out_dim, cudaStream_t stream);
// w_f32: att.out_proj.weight as fp32 [512][512]; mean_ws: [B][512] fp32 workspace
int token_mean_proj_launch(const __half* x, const float* w_f32, const float* bias, float* mean_ws, float* out, int B, int T,
                           cudaStream_t stream);

struct ScoreTailParams {
  const float* feats;  // [L][512]
  int L;
  const float* w_in;    // att_cross.in_proj_weight  [1536][512] (fp32: the tail decides the argmax)
  const float* b_in;    // [1536]
  const float* fold_v;  // [512] = out_proj.weight^T linear.weight: out_proj followed by Linear(512, 1) is one dot product
  float fold_c;         // linear.weight . out_proj.bias + linear.bias
  float offset;         // +100 of predict_score.py:207
  float* qkv;           // workspace [L][1536]
  float* scores;        // out [L]
  int* best;            // out, optional
  unsigned int* counter;  // device word, zero between launches (the last CTA takes the arg-max and resets it)
};
int score_tail_launch(const ScoreTailParams& p, cudaStream_t stream);

int pose_update_launch(const float* pose_in, const float* trans, const float* rot, float* pose_out,
                       float* trans_delta_out, float* rot_delta_out, int N, float trans_scale, float rot_normalizer,
                       cudaStream_t stream);
int f32_to_f16_launch(const float* x, __half* y, size_t n, cudaStream_t stream);

}  // namespace fp
