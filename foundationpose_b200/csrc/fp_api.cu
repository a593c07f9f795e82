// fp_api.cu — the product C ABI (include/fpose.h): context, weights, mesh, frame, and the per-frame
// hot loop (crops -> encoder -> heads -> pose update, K times; then scoring) enqueued on one stream
// with no host synchronisation.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <exception>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/fpose.h"
#include "fp_attn.cuh"
#include "fp_common.cuh"
#include "fp_crop.cuh"
#include "fp_depth.cuh"
#include "fp_gemm.cuh"

namespace fp {
const char* get_last_error();

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
};

// (Re)allocates `b` to at least `bytes`.  `epoch` is the owning context's graph epoch: it is bumped whenever a device
// pointer or by-value kernel parameter that a captured CUDA graph may hold changes (re-allocation, new mesh /
// weights / intrinsics), and cached graphs older than it are rebuilt.  Never called while a stream is capturing:
// every workspace is sized by ensure_capacity / fp_set_mesh / fp_set_frame BEFORE run_graphed.
static int dev_alloc(unsigned long long& epoch, DevBuf& b, size_t bytes, bool zero = false) {
  if (b.bytes >= bytes && b.p) return 0;
  ++epoch;
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.bytes = 0;
  FP_CUDA_OK(cudaMalloc(&b.p, bytes));
  b.bytes = bytes;
  if (zero) {
    // legacy-stream memset + full synchronisation: the consumers run on the caller's (possibly non-blocking) stream
    FP_CUDA_OK(cudaMemset(b.p, 0, bytes));
    FP_CUDA_OK(cudaDeviceSynchronize());
  }
  return 0;
}
template <class T>
static int upload(unsigned long long& epoch, DevBuf& b, const std::vector<T>& v) {
  const size_t bytes = std::max<size_t>(v.size() * sizeof(T), 16);
  if (dev_alloc(epoch, b, bytes)) return -2;
  if (!v.empty()) FP_CUDA_OK(cudaMemcpy(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

struct Tensor {
  void* p = nullptr;
  int dtype = 0;  // 0 = f32, 1 = f16
  long long numel = 0;
};

struct Net {
  std::map<std::string, Tensor> t;
  bool loaded = false;
  // fp_load_network verified that every name the execution plan uses is present; a miss is a programming error
  // and surfaces as an exception that the extern "C" wrappers turn into an error code
  const __half* h(const char* name) const { return reinterpret_cast<const __half*>(t.at(name).p); }
  const float* f(const char* name) const { return reinterpret_cast<const float*>(t.at(name).p); }
};

constexpr int S = 160;
constexpr int T = 400;
constexpr size_t kCropImg = (size_t)(S + 6) * (S + 8) * 8;  // fp16 elements per padded crop image
// The encoder's first stage runs on the A (rendered) and B (observed) crops as one batch.  The 40x40
// 128-channel layers tile four images per MMA (gemm_swap_patch_kernel), and the layer that fuses
// torch.cat((a, b), 1) stores A and B tiles to different channel halves, so the A/B boundary must fall on a tile
// boundary: B starts at N rounded up to 4 (up to three never-read pad images).
static inline int b_img0_of(int N) { return (N + 3) & ~3; }

}  // namespace fp

struct fp_ctx {
  int device = 0;
  fp::Net net[2];  // 0 = refiner, 1 = scorer
  unsigned long long epoch = 1;  // graph epoch (see dev_alloc)
  // mesh (fp_meshlet.cu layout)
  fp::DevBuf vpos, vnrm, vatt, faces, meshlets, ml_verts, ml_tris, tex;
  int V = 0, F = 0, Ht = 0, Wt = 0, n_meshlets = 0, front_sign = 0, mesh_closed = 0;
  float mesh_bs[4] = {0.f, 0.f, 0.f, 0.f};
  bool has_tex = false, has_mesh = false;
  float diameter = 0.f, rot_normalizer = 0.3490658503988659f;
  float crop_ratio[2] = {1.2f, 1.2f};  // per predictor: each reads its own config.yml (predict_pose_refine.py:117, predict_score.py:137)
  // frame
  fp::DevBuf rgb_raw, rgba, depth_raw, depth_a, depth_b, xyz;
  const float* depth_cur = nullptr;
  float K[9] = {0};
  int H = 0, W = 0;
  bool has_frame = false;
  // workspaces (sized for cap_n hypotheses)
  int cap_n = 0;
  fp::DevBuf crops, act0, a1, a2, a3, ab0, ab1, ab2, c0, c1, c2, tok, qkv, att, x1pre, x1, ff, x2pre;
  fp::DevBuf head_out, poses_a, poses_b, feats, tail_qkv, tail_attn, tail_proj, scores, best;
  int tail_cap = 0;
  float fold_c = 0.f;      // linear.weight . out_proj.bias + linear.bias (by-value kernel parameter)
  fp::DevBuf fold_v, tail_counter;  // out_proj^T linear.weight [512]; arg-max ticket
  // CUDA graphs of the launch-bound inner loops, keyed by (kind, N, iterations)
  struct GraphEntry {
    cudaGraphExec_t exec = nullptr;
    unsigned long long epoch = 0;
  };
  std::map<std::tuple<int, int, int>, GraphEntry> graphs;
  std::map<std::tuple<int, int, int>, int> graph_nodes;
  cudaStream_t cap_stream = nullptr;
  // the refiner's two decoder heads are independent after the shared attention core: at small batches
  // (one linear layer = 1-2 waves of tiles) the second head runs on `side_stream` so that its kernels fill
  // the SMs the first head's tail wave leaves idle.  Fork / join are events, captured into the graph.
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int fork_max_n = 128;  // FPOSE_FORK_MAX_N; 0 disables.  Measured (profiles/r02_fork_probe.log): -5 % at 32, -1 % at 126, +1 % at 252 hypotheses
  bool use_graphs = true;
  bool cull_backfaces = true;  // FPOSE_NO_CULL=1: render both sides even for closed meshes (A/B checks)
  bool track_valid = false;
  int crop_tile = 0;
  fp::DevBuf lt_buf, lr_buf, feat_buf, pose_stage, tok_mean;
  fp::DevBuf mask_buf, mask_stats, crop_stats;
  // fp_track: pinned host staging (frame in, pose out) so that the whole frame is ONE graph launch
  void* stage_rgb = nullptr;
  void* stage_depth = nullptr;
  float* stage_pose = nullptr;
  size_t stage_npix = 0;
  fp::DevBuf track_pose;
};

namespace fp {

static int ensure_capacity(fp_ctx* c, int N) {
  if (N <= c->cap_n) return 0;
  const size_t n = (size_t)N;
  int rc = 0;
  // the crop buffer is zeroed once: the 3-pixel border is never written afterwards
  const size_t m = 2 * n + 3;  // A + pad + B images
  rc |= dev_alloc(c->epoch, c->crops, m * kCropImg * 2, true);
  rc |= dev_alloc(c->epoch, c->act0, m * 80 * 80 * 64 * 2);
  rc |= dev_alloc(c->epoch, c->a1, m * 1600 * 128 * 2);
  rc |= dev_alloc(c->epoch, c->a2, m * 1600 * 128 * 2);
  rc |= dev_alloc(c->epoch, c->a3, m * 1600 * 128 * 2);
  rc |= dev_alloc(c->epoch, c->ab0, n * 1600 * 256 * 2);
  rc |= dev_alloc(c->epoch, c->ab1, n * 1600 * 256 * 2);
  rc |= dev_alloc(c->epoch, c->ab2, n * 1600 * 256 * 2);
  rc |= dev_alloc(c->epoch, c->c0, n * T * 512 * 2);
  rc |= dev_alloc(c->epoch, c->c1, n * T * 512 * 2);
  rc |= dev_alloc(c->epoch, c->c2, n * T * 512 * 2);
  rc |= dev_alloc(c->epoch, c->tok, n * T * 512 * 2);
  rc |= dev_alloc(c->epoch, c->qkv, n * T * 3072 * 2);
  rc |= dev_alloc(c->epoch, c->att, 2 * n * T * 512 * 2);
  // x 2: one set per decoder head (they may run concurrently, see run_refine_heads)
  rc |= dev_alloc(c->epoch, c->x1pre, 2 * n * T * 512 * 2);
  rc |= dev_alloc(c->epoch, c->x1, 2 * n * T * 512 * 2);
  rc |= dev_alloc(c->epoch, c->ff, 2 * n * T * 512 * 2);
  rc |= dev_alloc(c->epoch, c->x2pre, 2 * n * T * 512 * 2);
  rc |= dev_alloc(c->epoch, c->head_out, 2 * n * 3 * 4);
  rc |= dev_alloc(c->epoch, c->poses_a, n * 16 * 4);
  rc |= dev_alloc(c->epoch, c->poses_b, n * 16 * 4);
  rc |= dev_alloc(c->epoch, c->feats, n * 512 * 4);
  rc |= dev_alloc(c->epoch, c->lt_buf, n * 3 * 4);
  rc |= dev_alloc(c->epoch, c->lr_buf, n * 9 * 4);
  rc |= dev_alloc(c->epoch, c->feat_buf, n * 512 * 4);
  rc |= dev_alloc(c->epoch, c->pose_stage, n * 16 * 4);
  rc |= dev_alloc(c->epoch, c->tok_mean, n * 512 * 4);
  if (rc) return -2;
  c->cap_n = N;
  return 0;
}

static int ensure_tail(fp_ctx* c, int L) {
  if (L <= c->tail_cap) return 0;
  int rc = 0;
  rc |= dev_alloc(c->epoch, c->tail_qkv, (size_t)L * 1536 * 4);
  rc |= dev_alloc(c->epoch, c->tail_attn, (size_t)L * 512 * 4);
  rc |= dev_alloc(c->epoch, c->tail_proj, (size_t)L * 512 * 4);
  rc |= dev_alloc(c->epoch, c->scores, (size_t)L * 4);
  rc |= dev_alloc(c->epoch, c->best, 16);
  if (rc) return -2;
  c->tail_cap = L;
  return 0;
}

static GemmLayer mk(int kind, int n_img, int H, int W, int Cin, int Cout, const void* in, const __half* w,
                    const float* b, void* out, int relu, const void* res = nullptr, int out_ld = 0, int out_split = 0,
                    const float* post_add = nullptr) {
  GemmLayer L;
  L.kind = kind;
  L.n_img = n_img;
  L.Hin = H;
  L.Win = W;
  L.Cin = Cin;
  L.Cout = Cout;
  L.in = in;
  L.w = w;
  L.bias = b;
  L.res = res;
  L.res_ld = Cout;
  L.out = out;
  L.out_ld = out_ld ? out_ld : Cout;
  L.out_split = out_split;
  L.post_add = post_add;
  L.relu = relu;
  return L;
}

#define FP_TRY(expr)         \
  do {                       \
    int _rc = (expr);        \
    if (_rc) return _rc;     \
  } while (0)

// crops [2N][166][168][8] -> tokens [N][400][512] (+ positional embedding)
static int run_encoder(fp_ctx* c, const Net& net, const __half* crops, int N, cudaStream_t st) {
  char wn[32], bn[32];
  auto W = [&](int i) { snprintf(wn, sizeof wn, "enc.%d.w", i); return net.h(wn); };
  auto B = [&](int i) { snprintf(bn, sizeof bn, "enc.%d.b", i); return net.f(bn); };
  const int Np = b_img0_of(N);
  const int M = Np + N;
  FP_TRY(gemm_layer_launch(mk(LK_CONV7_S2, M, S, S, 8, 64, crops, W(0), B(0), c->act0.p, 1), st));
  FP_TRY(gemm_layer_launch(mk(LK_CONV3_S2, M, 80, 80, 64, 128, c->act0.p, W(1), B(1), c->a1.p, 1), st));
  FP_TRY(gemm_layer_launch(mk(LK_CONV3_S1, M, 40, 40, 128, 128, c->a1.p, W(2), B(2), c->a2.p, 1), st));
  FP_TRY(gemm_layer_launch(mk(LK_CONV3_S1, M, 40, 40, 128, 128, c->a2.p, W(3), B(3), c->a3.p, 1, c->a1.p), st));
  FP_TRY(gemm_layer_launch(mk(LK_CONV3_S1, M, 40, 40, 128, 128, c->a3.p, W(4), B(4), c->a2.p, 1), st));
  // last encodeA layer writes straight into the 256-channel concat buffer (refine_network.py:85)
  FP_TRY(gemm_layer_launch(mk(LK_CONV3_S1, M, 40, 40, 128, 128, c->a2.p, W(5), B(5), c->ab0.p, 1, c->a3.p, 256, Np), st));
  FP_TRY(gemm_layer_launch(mk(LK_CONV3_S1, N, 40, 40, 256, 256, c->ab0.p, W(6), B(6), c->ab1.p, 1), st));
  FP_TRY(gemm_layer_launch(mk(LK_CONV3_S1, N, 40, 40, 256, 256, c->ab1.p, W(7), B(7), c->ab2.p, 1, c->ab0.p), st));
  FP_TRY(gemm_layer_launch(mk(LK_CONV3_S1, N, 40, 40, 256, 256, c->ab2.p, W(8), B(8), c->ab1.p, 1), st));
  FP_TRY(gemm_layer_launch(mk(LK_CONV3_S1, N, 40, 40, 256, 256, c->ab1.p, W(9), B(9), c->ab0.p, 1, c->ab2.p), st));
  FP_TRY(gemm_layer_launch(mk(LK_CONV3_S2, N, 40, 40, 256, 512, c->ab0.p, W(10), B(10), c->c0.p, 1), st));
  FP_TRY(gemm_layer_launch(mk(LK_CONV3_S1, N, 20, 20, 512, 512, c->c0.p, W(11), B(11), c->c1.p, 1), st));
  FP_TRY(gemm_layer_launch(mk(LK_CONV3_S1, N, 20, 20, 512, 512, c->c1.p, W(12), B(12), c->c2.p, 1, c->c0.p), st));
  FP_TRY(gemm_layer_launch(mk(LK_CONV3_S1, N, 20, 20, 512, 512, c->c2.p, W(13), B(13), c->c1.p, 1), st));
  FP_TRY(gemm_layer_launch(
      mk(LK_CONV3_S1, N, 20, 20, 512, 512, c->c1.p, W(14), B(14), c->tok.p, 1, c->c2.p, 0, 0, net.f("pe")), st));
  return 0;
}

// tokens -> (trans, rot) raw network outputs, [2][N][3] fp32 in head_out
static int run_refine_heads(fp_ctx* c, const Net& net, int N, cudaStream_t st) {
  const int M = N * T;
  // both heads' in_proj as one GEMM: [M,512] x [3072,512]^T
  FP_TRY(gemm_layer_launch(mk(LK_LINEAR, 1, 1, M, 512, 3072, c->tok.p, net.h("heads.in_w"), net.f("heads.in_b"), c->qkv.p, 0), st));
  AttnParams ap;
  ap.qkv = reinterpret_cast<const __half*>(c->qkv.p);
  ap.ld = 3072;
  ap.q_off = 0;
  ap.k_off = 512;
  ap.v_off = 1024;
  ap.group_col_stride = 1536;
  ap.n_groups = 2;
  ap.out = reinterpret_cast<__half*>(c->att.p);
  ap.ld_out = 512;
  ap.out_group_stride = (size_t)M * 512;
  ap.B = N;
  ap.T = T;
  ap.n_heads = 4;
  ap.scale = 0.08838834764831845f;
  FP_TRY(attn_core_launch(ap, st));
  const bool fork = N <= c->fork_max_n;
  if (fork) {
    if (!c->side_stream) {
      FP_CUDA_OK(cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking));
      FP_CUDA_OK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
      FP_CUDA_OK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    }
    FP_CUDA_OK(cudaEventRecord(c->ev_fork, st));
    FP_CUDA_OK(cudaStreamWaitEvent(c->side_stream, c->ev_fork, 0));
  }
  for (int g = 0; g < 2; ++g) {
    cudaStream_t sg = (fork && g == 1) ? c->side_stream : st;
    char nm[48];
    auto H = [&](const char* s) { snprintf(nm, sizeof nm, "head%d.%s", g, s); return net.h(nm); };
    auto Fp = [&](const char* s) { snprintf(nm, sizeof nm, "head%d.%s", g, s); return net.f(nm); };
    const size_t off = (size_t)g * M * 512;
    const __half* att_g = reinterpret_cast<const __half*>(c->att.p) + off;
    __half* x1pre = reinterpret_cast<__half*>(c->x1pre.p) + off;
    __half* x1 = reinterpret_cast<__half*>(c->x1.p) + off;
    __half* ff = reinterpret_cast<__half*>(c->ff.p) + off;
    __half* x2pre = reinterpret_cast<__half*>(c->x2pre.p) + off;
    const __half* w;
    const float* b;
    w = H("out_w"); b = Fp("out_b");
    // the first kernel behind an event wait has a full (not programmatic) dependency
    if (fork && g == 1) pdl_skip_next();
    FP_TRY(gemm_layer_launch(mk(LK_LINEAR, 1, 1, M, 512, 512, att_g, w, b, x1pre, 0, c->tok.p), sg));
    const float* g1 = Fp("ln1_g");
    const float* b1 = Fp("ln1_b");
    FP_TRY(layernorm_launch(x1pre, x1, g1, b1, M, sg));
    w = H("ff1_w"); b = Fp("ff1_b");
    FP_TRY(gemm_layer_launch(mk(LK_LINEAR, 1, 1, M, 512, 512, x1, w, b, ff, 1), sg));
    w = H("ff2_w"); b = Fp("ff2_b");
    FP_TRY(gemm_layer_launch(mk(LK_LINEAR, 1, 1, M, 512, 512, ff, w, b, x2pre, 0, x1), sg));
    const float* g2 = Fp("ln2_g");
    const float* b2 = Fp("ln2_b");
    const float* fw = Fp("fin_w");
    const float* fb = Fp("fin_b");
    FP_TRY(head_final_launch(x2pre, g2, b2, fw, fb, reinterpret_cast<float*>(c->head_out.p) + (size_t)g * N * 3, N, T, 3, sg));
  }
  if (fork) {
    FP_CUDA_OK(cudaEventRecord(c->ev_join, c->side_stream));
    FP_CUDA_OK(cudaStreamWaitEvent(st, c->ev_join, 0));
    pdl_skip_next();  // the consumer of head_out joins two streams
  }
  return 0;
}

// tokens -> per-hypothesis 512-d features (score_network.py:72-74)
static int run_score_feats(fp_ctx* c, const Net& net, int N, float* feats, cudaStream_t st) {
  const int M = N * T;
  FP_TRY(gemm_layer_launch(mk(LK_LINEAR, 1, 1, M, 512, 1536, c->tok.p, net.h("att.in_w"), net.f("att.in_b"), c->qkv.p, 0), st));
  AttnParams ap;
  ap.qkv = reinterpret_cast<const __half*>(c->qkv.p);
  ap.ld = 1536;
  ap.q_off = 0;
  ap.k_off = 512;
  ap.v_off = 1024;
  ap.group_col_stride = 0;
  ap.n_groups = 1;
  ap.out = reinterpret_cast<__half*>(c->att.p);
  ap.ld_out = 512;
  ap.out_group_stride = 0;
  ap.B = N;
  ap.T = T;
  ap.n_heads = 4;
  ap.scale = 0.08838834764831845f;
  FP_TRY(attn_core_launch(ap, st));
  FP_TRY(token_mean_proj_launch(reinterpret_cast<const __half*>(c->att.p), net.f("att.out_w32"), net.f("att.out_b"),
                                reinterpret_cast<float*>(c->tok_mean.p), feats, N, T, st));
  return 0;
}

// external crop layout of the test hooks: [2N] images, A then B, contiguous
static int crops_import(fp_ctx* c, const void* ext, int N, cudaStream_t st) {
  const size_t img = kCropImg * 2;
  __half* dst = reinterpret_cast<__half*>(c->crops.p);
  const char* src = reinterpret_cast<const char*>(ext);
  FP_CUDA_OK(cudaMemcpyAsync(dst, src, (size_t)N * img, cudaMemcpyDeviceToDevice, st));
  FP_CUDA_OK(cudaMemcpyAsync(dst + (size_t)b_img0_of(N) * kCropImg, src + (size_t)N * img, (size_t)N * img,
                             cudaMemcpyDeviceToDevice, st));
  return 0;
}
static int crops_export(fp_ctx* c, void* ext, int N, cudaStream_t st) {
  const size_t img = kCropImg * 2;
  const __half* src = reinterpret_cast<const __half*>(c->crops.p);
  char* dst = reinterpret_cast<char*>(ext);
  FP_CUDA_OK(cudaMemcpyAsync(dst, src, (size_t)N * img, cudaMemcpyDeviceToDevice, st));
  FP_CUDA_OK(cudaMemcpyAsync(dst + (size_t)N * img, src + (size_t)b_img0_of(N) * kCropImg, (size_t)N * img,
                             cudaMemcpyDeviceToDevice, st));
  return 0;
}

static int make_crops(fp_ctx* c, const float* poses, int N, int mode, float* dbg, float* win, int* stats,
                      cudaStream_t st) {
  FP_REQUIRE(c->has_mesh, "no mesh: call fp_set_mesh first");
  FP_REQUIRE(c->has_frame, "no frame: call fp_set_frame first");
  // only launches below: this body is also what run_graphed captures (no allocation, no synchronisation)
  CropParams p;
  p.poses = poses;
  p.N = N;
  p.fx = c->K[0];
  p.fy = c->K[4];
  p.cx = c->K[2];
  p.cy = c->K[5];
  p.H = c->H;
  p.W = c->W;
  p.r3 = (float)((double)c->diameter * (double)c->crop_ratio[mode ? 1 : 0] / 2.0);
  p.inv_radius = 1.0f / (c->diameter / 2.0f);
  p.znear = 0.001f;  // Utils.py:161
  p.zfar = 100.f;
  p.mesh.vpos = reinterpret_cast<const float4*>(c->vpos.p);
  p.mesh.vnrm = reinterpret_cast<const float4*>(c->vnrm.p);
  p.mesh.vatt = reinterpret_cast<const float4*>(c->vatt.p);
  p.mesh.faces = reinterpret_cast<const int4*>(c->faces.p);
  p.mesh.meshlets = reinterpret_cast<const Meshlet*>(c->meshlets.p);
  p.mesh.ml_verts = reinterpret_cast<const int*>(c->ml_verts.p);
  p.mesh.ml_tris = reinterpret_cast<const uint2*>(c->ml_tris.p);
  p.mesh.n_meshlets = c->n_meshlets;
  p.mesh.V = c->V;
  p.mesh.F = c->F;
  p.mesh.front_sign = c->cull_backfaces ? c->front_sign : 0;
  p.mesh.bs_x = c->mesh_bs[0];
  p.mesh.bs_y = c->mesh_bs[1];
  p.mesh.bs_z = c->mesh_bs[2];
  p.mesh.bs_r = c->mesh_bs[3];
  p.has_tex = c->has_tex ? 1 : 0;
  p.tex = c->has_tex ? reinterpret_cast<const uchar4*>(c->tex.p) : nullptr;
  p.Ht = c->Ht;
  p.Wt = c->Wt;
  p.rgb = reinterpret_cast<const uchar4*>(c->rgba.p);
  p.xyz_map = reinterpret_cast<const float4*>(c->xyz.p);
  p.depth = c->depth_cur;
  p.mode = mode;
  p.crops = reinterpret_cast<__half*>(c->crops.p);
  p.b_img0 = b_img0_of(N);
  p.dbg = dbg;
  p.win_out = win;
  p.stats = stats;
  p.tile_override = c->crop_tile;
  return crop_launch(p, st);
}

// Runs `body(stream)` — a fixed sequence of kernel launches (and fixed-address copies) on ctx-owned buffers —
// through a cached CUDA graph: first sight of a key runs eagerly (sets function attributes), the second
// captures + instantiates, later calls replay.  Replay removes ~170 launch + 60 tensor-map-encode host
// calls per register(), which is what bounds track_one() and small per-GPU shards.  The body never allocates:
// callers size every workspace first, so a capture after an epoch bump (new mesh, new N) is safe.
template <class Body>
static int run_graphed(fp_ctx* c, int kind, int N, int iters, cudaStream_t st, Body body) {
  if (!c->use_graphs || g_prof_on) return body(st);
  const auto key = std::make_tuple(kind, N, iters);
  auto it = c->graphs.find(key);
  if (it == c->graphs.end()) {
    c->graphs[key] = fp_ctx::GraphEntry();  // seen once: next call captures
    return body(st);
  }
  fp_ctx::GraphEntry& g = it->second;
  if (g.exec == nullptr || g.epoch != c->epoch) {
    if (g.exec) {
      cudaGraphExecDestroy(g.exec);
      g.exec = nullptr;
    }
    if (!c->cap_stream) FP_CUDA_OK(cudaStreamCreateWithFlags(&c->cap_stream, cudaStreamNonBlocking));
    FP_CUDA_OK(cudaStreamBeginCapture(c->cap_stream, cudaStreamCaptureModeThreadLocal));
    set_capturing(true);  // nothing runs during capture: launches are counted per replay
    int rc = 0;
    try {
      rc = body(c->cap_stream);
    } catch (...) {
      rc = -3;
      set_last_error("exception while capturing the launch sequence");
    }
    set_capturing(false);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(c->cap_stream, &graph);
    if (rc != 0) {
      if (graph) cudaGraphDestroy(graph);
      return rc;
    }
    FP_CUDA_OK(ce);
    size_t n_nodes = 0;
    cudaGraphGetNodes(graph, nullptr, &n_nodes);
    size_t n_kernels = 0;
    {
      std::vector<cudaGraphNode_t> nodes(n_nodes);
      if (n_nodes && cudaGraphGetNodes(graph, nodes.data(), &n_nodes) == cudaSuccess)
        for (auto nd : nodes) {
          cudaGraphNodeType ty;
          if (cudaGraphNodeGetType(nd, &ty) == cudaSuccess && ty == cudaGraphNodeTypeKernel) ++n_kernels;
        }
    }
    const cudaError_t ie = cudaGraphInstantiate(&g.exec, graph, 0);
    cudaGraphDestroy(graph);
    FP_CUDA_OK(ie);
    g.epoch = c->epoch;
    c->graph_nodes[key] = (int)n_kernels;
  }
  FP_CUDA_OK(cudaGraphLaunch(g.exec, st));
  note_launches(c->graph_nodes[key]);
  return 0;
}

// RAII: make the context's device current for the duration of an entry point
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

static int set_frame_launches(fp_ctx* c, const unsigned char* rgb_dev, const float* depth_dev, int flags, float zfar,
                              cudaStream_t st) {
  const int H = c->H, W = c->W;
  const size_t npix = (size_t)H * W;
  static int fused = -1;  // FPOSE_FUSED_PREP=0: the four separate launches (A/B)
  if (fused < 0) {
    const char* e = getenv("FPOSE_FUSED_PREP");
    fused = (e && e[0] == '0') ? 0 : 1;
  }
  if ((flags & FP_FRAME_FILTER_DEPTH) && !fused) {
    FP_TRY(rgb_to_rgba_launch(rgb_dev, reinterpret_cast<uchar4*>(c->rgba.p), (int)npix, st));
    FP_TRY(erode_depth_launch(depth_dev, reinterpret_cast<float*>(c->depth_a.p), H, W, 2, 0.001f, 0.8f, 100.f, st));
    FP_TRY(bilateral_depth_launch(reinterpret_cast<const float*>(c->depth_a.p), reinterpret_cast<float*>(c->depth_b.p), H,
                                  W, 2, 100.f, 2.f, 100000.f, st));
    c->depth_cur = reinterpret_cast<const float*>(c->depth_b.p);
    return depth_to_xyz_launch(c->depth_cur, reinterpret_cast<float4*>(c->xyz.p), H, W, c->K[0], c->K[4], c->K[2], c->K[5], zfar, st);
  }
  if (flags & FP_FRAME_FILTER_DEPTH) {
    // estimater.py:173-174 erode_depth(radius=2), bilateral_filter_depth(radius=2); :214 depth2xyzmap: one launch
    c->depth_cur = reinterpret_cast<const float*>(c->depth_b.p);
    return frame_prep_launch(rgb_dev, depth_dev, reinterpret_cast<uchar4*>(c->rgba.p), reinterpret_cast<float*>(c->depth_b.p),
                             reinterpret_cast<float4*>(c->xyz.p), H, W, c->K[0], c->K[4], c->K[2], c->K[5], zfar, st);
  }
  FP_TRY(rgb_to_rgba_launch(rgb_dev, reinterpret_cast<uchar4*>(c->rgba.p), (int)npix, st));
  FP_CUDA_OK(cudaMemcpyAsync(c->depth_b.p, depth_dev, npix * 4, cudaMemcpyDeviceToDevice, st));
  c->depth_cur = reinterpret_cast<const float*>(c->depth_b.p);
  FP_TRY(depth_to_xyz_launch(c->depth_cur, reinterpret_cast<float4*>(c->xyz.p), H, W, c->K[0], c->K[4], c->K[2], c->K[5],
                             zfar, st));
  return 0;
}

// frame buffers + intrinsics; bumps the epoch when a by-value kernel parameter changes
static int prepare_frame(fp_ctx* c, const float* K, int H, int W, bool need_raw) {
  const size_t npix = (size_t)H * W;
  FP_TRY(dev_alloc(c->epoch, c->rgba, npix * 4));
  FP_TRY(dev_alloc(c->epoch, c->depth_a, npix * 4));
  FP_TRY(dev_alloc(c->epoch, c->depth_b, npix * 4));
  FP_TRY(dev_alloc(c->epoch, c->xyz, npix * 16));
  if (need_raw) {
    FP_TRY(dev_alloc(c->epoch, c->rgb_raw, npix * 3));
    FP_TRY(dev_alloc(c->epoch, c->depth_raw, npix * 4));
  }
  bool same = (c->H == H && c->W == W);
  for (int i = 0; i < 9; ++i) same = same && (c->K[i] == K[i]);
  if (!same) ++c->epoch;  // intrinsics / frame size are by-value kernel parameters
  for (int i = 0; i < 9; ++i) c->K[i] = K[i];
  c->H = H;
  c->W = W;
  return 0;
}

static int refine_body(fp_ctx* c, int N, int iterations, cudaStream_t s2) {
  float* cur = reinterpret_cast<float*>(c->poses_a.p);
  float* nxt = reinterpret_cast<float*>(c->poses_b.p);
  const float* ho = reinterpret_cast<const float*>(c->head_out.p);
  for (int it = 0; it < iterations; ++it) {
    FP_TRY(make_crops(c, cur, N, 0, nullptr, nullptr, nullptr, s2));
    FP_TRY(run_encoder(c, c->net[0], reinterpret_cast<const __half*>(c->crops.p), N, s2));
    FP_TRY(run_refine_heads(c, c->net[0], N, s2));
    const bool last = it == iterations - 1;
    FP_TRY(pose_update_launch(cur, ho, ho + (size_t)N * 3, nxt, last ? reinterpret_cast<float*>(c->lt_buf.p) : nullptr,
                              last ? reinterpret_cast<float*>(c->lr_buf.p) : nullptr, N, c->diameter / 2.0f,
                              c->rot_normalizer, s2));
    float* t = cur;
    cur = nxt;
    nxt = t;
  }
  return 0;
}

}  // namespace fp

using namespace fp;

// every entry point: exceptions never cross the C boundary, the context's device is current inside
#define FP_API_BEGIN try {
#define FP_API_END                                                        \
  }                                                                       \
  catch (const std::exception& e) {                                       \
    fp::set_last_error("%s: exception: %s", __func__, e.what());          \
    return -3;                                                            \
  }                                                                       \
  catch (...) {                                                           \
    fp::set_last_error("%s: unknown exception", __func__);                \
    return -3;                                                            \
  }

extern "C" {

int fp_create(fp_ctx** out) {
  FP_API_BEGIN
  if (!out) {
    set_last_error("fp_create: null output");
    return -1;
  }
  int dev = 0;
  FP_CUDA_OK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  FP_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
  FP_REQUIRE(prop.major == 10, "libfpose targets sm_100a (B200); device %d is sm_%d%d", dev, prop.major, prop.minor);
  fp_ctx* c = new fp_ctx();
  c->device = dev;
  const char* ng = getenv("FPOSE_NO_GRAPH");
  c->use_graphs = !(ng && ng[0] == '1');
  const char* nc = getenv("FPOSE_NO_CULL");
  c->cull_backfaces = !(nc && nc[0] == '1');
  if (const char* fk = getenv("FPOSE_FORK_MAX_N")) c->fork_max_n = atoi(fk);
  *out = c;
  return 0;
  FP_API_END
}

int fp_destroy(fp_ctx* c) {
  FP_API_BEGIN
  if (!c) return 0;
  DeviceGuard dg(c->device);
  cudaDeviceSynchronize();
  for (auto& net : c->net)
    for (auto& kv : net.t) cudaFree(kv.second.p);
  DevBuf* bufs[] = {&c->vpos, &c->vnrm, &c->vatt, &c->faces, &c->meshlets, &c->ml_verts, &c->ml_tris, &c->tex, &c->rgb_raw,
                    &c->rgba, &c->depth_raw, &c->depth_a, &c->depth_b, &c->xyz, &c->crops, &c->act0, &c->a1, &c->a2,
                    &c->a3, &c->ab0, &c->ab1, &c->ab2, &c->c0, &c->c1, &c->c2, &c->tok, &c->qkv, &c->att, &c->x1pre,
                    &c->x1, &c->ff, &c->x2pre, &c->head_out, &c->poses_a, &c->poses_b, &c->feats, &c->tail_qkv,
                    &c->tail_attn, &c->tail_proj, &c->scores, &c->best, &c->lt_buf, &c->lr_buf, &c->feat_buf,
                    &c->pose_stage, &c->mask_buf, &c->mask_stats, &c->crop_stats, &c->track_pose, &c->fold_v, &c->tail_counter, &c->tok_mean};
  for (DevBuf* b : bufs)
    if (b->p) cudaFree(b->p);
  for (auto& kv : c->graphs)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  if (c->cap_stream) cudaStreamDestroy(c->cap_stream);
  if (c->side_stream) cudaStreamDestroy(c->side_stream);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  if (c->stage_rgb) cudaFreeHost(c->stage_rgb);
  if (c->stage_depth) cudaFreeHost(c->stage_depth);
  if (c->stage_pose) cudaFreeHost(c->stage_pose);
  delete c;
  return 0;
  FP_API_END
}

int fp_set_config(fp_ctx* c, int which, float crop_ratio, float rot_normalizer) {
  FP_API_BEGIN
  FP_REQUIRE(c, "null ctx");
  FP_REQUIRE(which == 0 || which == 1, "fp_set_config: which must be 0 (refiner) or 1 (scorer)");
  FP_REQUIRE(crop_ratio > 0.f, "fp_set_config: crop_ratio must be positive");
  c->crop_ratio[which] = crop_ratio;
  if (which == 0) c->rot_normalizer = rot_normalizer;
  ++c->epoch;
  return 0;
  FP_API_END
}

int fp_load_network(fp_ctx* c, int which, const fp_tensor_t* tensors, int n) {
  FP_API_BEGIN
  FP_REQUIRE(c && tensors, "fp_load_network: null argument");
  FP_REQUIRE(which == 0 || which == 1, "fp_load_network: which must be 0 (refiner) or 1 (scorer)");
  DeviceGuard dg(c->device);
  Net& net = c->net[which];
  ++c->epoch;
  FP_CUDA_OK(cudaDeviceSynchronize());
  for (auto& kv : net.t) cudaFree(kv.second.p);
  net.t.clear();
  net.loaded = false;
  for (int i = 0; i < n; ++i) {
    const fp_tensor_t& t = tensors[i];
    FP_REQUIRE(t.name && t.data && t.numel > 0, "fp_load_network: bad tensor #%d", i);
    Tensor d;
    d.dtype = t.dtype;
    d.numel = t.numel;
    const size_t bytes = (size_t)t.numel * (t.dtype == 1 ? 2 : 4);
    FP_CUDA_OK(cudaMalloc(&d.p, bytes));
    FP_CUDA_OK(cudaMemcpy(d.p, t.data, bytes, cudaMemcpyHostToDevice));
    auto old = net.t.find(t.name);
    if (old != net.t.end()) cudaFree(old->second.p);
    net.t[t.name] = d;
  }
  // verify that everything the execution plan needs is present, with the right size
  std::vector<std::pair<std::string, long long>> need;
  const int cin[15] = {0, 64, 128, 128, 128, 128, 256, 256, 256, 256, 256, 512, 512, 512, 512};
  const int cout[15] = {64, 128, 128, 128, 128, 128, 256, 256, 256, 256, 512, 512, 512, 512, 512};
  for (int i = 0; i < 15; ++i) {
    const long long k = i == 0 ? 7 * 64 : 9LL * cin[i];
    need.push_back({"enc." + std::to_string(i) + ".w", k * cout[i]});
    need.push_back({"enc." + std::to_string(i) + ".b", cout[i]});
  }
  need.push_back({"pe", 400LL * 512});
  if (which == 0) {
    need.push_back({"heads.in_w", 3072LL * 512});
    need.push_back({"heads.in_b", 3072});
    for (int g = 0; g < 2; ++g) {
      const std::string h = "head" + std::to_string(g) + ".";
      for (const char* s : {"out_w", "ff1_w", "ff2_w"}) need.push_back({h + s, 512LL * 512});
      for (const char* s : {"out_b", "ff1_b", "ff2_b", "ln1_g", "ln1_b", "ln2_g", "ln2_b"}) need.push_back({h + s, 512});
      need.push_back({h + "fin_w", 3LL * 512});
      need.push_back({h + "fin_b", 3});
    }
  } else {
    need.push_back({"att.in_w", 1536LL * 512});
    need.push_back({"att.in_b", 1536});
    need.push_back({"att.out_w32", 512LL * 512});
    need.push_back({"att.out_b", 512});
    need.push_back({"cross.in_w", 1536LL * 512});
    need.push_back({"cross.in_b", 1536});
    need.push_back({"cross.out_w", 512LL * 512});
    need.push_back({"cross.out_b", 512});
    need.push_back({"lin.w", 512});
    need.push_back({"lin.b", 1});
  }
  for (auto& nd : need) {
    auto it = net.t.find(nd.first);
    FP_REQUIRE(it != net.t.end(), "fp_load_network: tensor '%s' missing", nd.first.c_str());
    FP_REQUIRE(it->second.numel == nd.second, "fp_load_network: tensor '%s' has %lld elements, expected %lld",
               nd.first.c_str(), it->second.numel, nd.second);
  }
  if (which == 1) {
    // score = linear(out_proj(a)) = (W_out^T w_lin) . a + (w_lin . b_out + b_lin): fold once, in fp64
    const float *wo = nullptr, *bo = nullptr, *wl = nullptr, *bl = nullptr;
    for (int i = 0; i < n; ++i) {
      const std::string nm = tensors[i].name;
      const float* d = reinterpret_cast<const float*>(tensors[i].data);
      if (tensors[i].dtype != 0) continue;
      if (nm == "cross.out_w") wo = d;
      else if (nm == "cross.out_b") bo = d;
      else if (nm == "lin.w") wl = d;
      else if (nm == "lin.b") bl = d;
    }
    FP_REQUIRE(wo && bo && wl && bl, "fp_load_network: the scorer tail tensors must be float32");
    std::vector<float> v(512);
    for (int i = 0; i < 512; ++i) {
      double acc = 0.0;
      for (int o = 0; o < 512; ++o) acc += (double)wl[o] * (double)wo[(size_t)o * 512 + i];
      v[i] = (float)acc;
    }
    double cc = (double)bl[0];
    for (int o = 0; o < 512; ++o) cc += (double)wl[o] * (double)bo[o];
    c->fold_c = (float)cc;
    FP_TRY(upload(c->epoch, c->fold_v, v));
    FP_TRY(dev_alloc(c->epoch, c->tail_counter, 16, /*zero=*/true));
  }
  net.loaded = true;
  return 0;
  FP_API_END
}

int fp_set_mesh(fp_ctx* c, int V, int F, const float* pos, const float* nrm, const float* uv, const float* vcol,
                const int* faces, const unsigned char* tex_rgb, int Ht, int Wt, float diameter) {
  FP_API_BEGIN
  FP_REQUIRE(c && pos && nrm && faces, "fp_set_mesh: null argument");
  FP_REQUIRE(V > 0 && F > 0 && diameter > 0.f, "fp_set_mesh: empty mesh");
  FP_REQUIRE((uv && tex_rgb && Ht > 0 && Wt > 0) || vcol, "fp_set_mesh: need (uv + texture) or vertex colours");
  for (int i = 0; i < 3 * F; ++i) FP_REQUIRE(faces[i] >= 0 && faces[i] < V, "fp_set_mesh: face index out of range");
  DeviceGuard dg(c->device);
  // graphs captured for the previous mesh may still be running on the caller's stream
  FP_CUDA_OK(cudaDeviceSynchronize());
  c->has_mesh = false;
  const bool has_tex = (uv && tex_rgb);
  MeshHost mh;
  FP_TRY(build_mesh_host(V, F, pos, nrm, has_tex ? uv : vcol, has_tex ? 2 : 3, faces, mh));
  FP_TRY(upload(c->epoch, c->vpos, mh.vpos));
  FP_TRY(upload(c->epoch, c->vnrm, mh.vnrm));
  FP_TRY(upload(c->epoch, c->vatt, mh.vatt));
  FP_TRY(upload(c->epoch, c->faces, mh.faces));
  FP_TRY(upload(c->epoch, c->meshlets, mh.meshlets));
  FP_TRY(upload(c->epoch, c->ml_verts, mh.ml_verts));
  FP_TRY(upload(c->epoch, c->ml_tris, mh.ml_tris));
  c->has_tex = has_tex;
  if (has_tex) {
    std::vector<unsigned char> rgba((size_t)Ht * Wt * 4);
    for (size_t i = 0; i < (size_t)Ht * Wt; ++i) {
      rgba[4 * i] = tex_rgb[3 * i];
      rgba[4 * i + 1] = tex_rgb[3 * i + 1];
      rgba[4 * i + 2] = tex_rgb[3 * i + 2];
      rgba[4 * i + 3] = 255;
    }
    FP_TRY(upload(c->epoch, c->tex, rgba));
    c->Ht = Ht;
    c->Wt = Wt;
  }
  c->V = V;
  c->F = F;
  c->n_meshlets = (int)mh.meshlets.size();
  c->front_sign = mh.front_sign;
  c->mesh_closed = mh.closed;
  for (int i = 0; i < 4; ++i) c->mesh_bs[i] = mh.bs[i];
  c->diameter = diameter;
  c->has_mesh = true;
  ++c->epoch;
  return 0;
  FP_API_END
}

int fp_set_crop_tile(fp_ctx* c, int tile) {
  FP_API_BEGIN
  FP_REQUIRE(c && (tile == 0 || tile == 16 || tile == 32 || tile == 80), "fp_set_crop_tile: tile must be 0 (automatic), 16, 32 or 80");
  c->crop_tile = tile;
  ++c->epoch;
  return 0;
  FP_API_END
}

int fp_mesh_info(fp_ctx* c, int* info) {
  FP_API_BEGIN
  FP_REQUIRE(c && info && c->has_mesh, "fp_mesh_info: no mesh");
  info[0] = c->n_meshlets;
  info[1] = c->mesh_closed;
  info[2] = c->cull_backfaces ? c->front_sign : 0;
  info[3] = c->V;
  info[4] = c->F;
  return 0;
  FP_API_END
}

int fp_set_frame(fp_ctx* c, const unsigned char* rgb, const float* depth, const float* K, int H, int W, int flags,
                 float zfar, void* stream) {
  FP_API_BEGIN
  FP_REQUIRE(c && rgb && depth && K, "fp_set_frame: null argument");
  FP_REQUIRE(H > 0 && W > 0, "fp_set_frame: empty frame");
  DeviceGuard dg(c->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t npix = (size_t)H * W;
  c->has_frame = false;
  const bool on_dev = (flags & FP_FRAME_ON_DEVICE) != 0;
  FP_TRY(prepare_frame(c, K, H, W, !on_dev));
  const unsigned char* rgb_dev = rgb;
  const float* depth_dev = depth;
  if (!on_dev) {
    FP_CUDA_OK(cudaMemcpyAsync(c->rgb_raw.p, rgb, npix * 3, cudaMemcpyHostToDevice, st));
    FP_CUDA_OK(cudaMemcpyAsync(c->depth_raw.p, depth, npix * 4, cudaMemcpyHostToDevice, st));
    rgb_dev = reinterpret_cast<const unsigned char*>(c->rgb_raw.p);
    depth_dev = reinterpret_cast<const float*>(c->depth_raw.p);
  }
  FP_TRY(set_frame_launches(c, rgb_dev, depth_dev, flags, zfar, st));
  c->has_frame = true;
  return 0;
  FP_API_END
}

int fp_set_xyz_map(fp_ctx* c, const float* xyz, void* stream) {
  FP_API_BEGIN
  FP_REQUIRE(c && xyz, "fp_set_xyz_map: null argument");
  FP_REQUIRE(c->has_frame, "fp_set_xyz_map: no frame (call fp_set_frame first)");
  DeviceGuard dg(c->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // [H][W][3] (host or device) -> the float4-per-pixel layout the crop kernel samples
  FP_CUDA_OK(cudaMemcpy2DAsync(c->xyz.p, 16, xyz, 12, 12, (size_t)c->H * c->W, cudaMemcpyDefault, st));
  return 0;
  FP_API_END
}

int fp_get_depth(fp_ctx* c, float* depth_out_dev, float* xyz_out_dev, void* stream) {
  FP_API_BEGIN
  FP_REQUIRE(c && c->has_frame, "fp_get_depth: no frame");
  DeviceGuard dg(c->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t npix = (size_t)c->H * c->W;
  if (depth_out_dev) FP_CUDA_OK(cudaMemcpyAsync(depth_out_dev, c->depth_cur, npix * 4, cudaMemcpyDeviceToDevice, st));
  // internal layout is float4 per pixel; the hook returns the reference's [H][W][3]
  if (xyz_out_dev)
    FP_CUDA_OK(cudaMemcpy2DAsync(xyz_out_dev, 12, c->xyz.p, 16, 12, npix, cudaMemcpyDeviceToDevice, st));
  return 0;
  FP_API_END
}

int fp_start_poses(fp_ctx* c, const unsigned char* mask, int mask_on_device, const float* rot_grid, int N, float* poses_out,
                   float* info_out, void* stream) {
  FP_API_BEGIN
  FP_REQUIRE(c && mask && rot_grid && poses_out && info_out && N >= 0, "fp_start_poses: bad argument");
  FP_REQUIRE(c->has_frame, "fp_start_poses: no frame (call fp_set_frame first)");
  DeviceGuard dg(c->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t npix = (size_t)c->H * c->W;
  const unsigned char* mdev = mask;
  if (!mask_on_device) {
    FP_TRY(dev_alloc(c->epoch, c->mask_buf, npix));
    FP_CUDA_OK(cudaMemcpyAsync(c->mask_buf.p, mask, npix, cudaMemcpyHostToDevice, st));
    mdev = reinterpret_cast<const unsigned char*>(c->mask_buf.p);
  }
  FP_TRY(dev_alloc(c->epoch, c->mask_stats, 64));
  return start_poses_launch(c->depth_cur, mdev, c->H, c->W, c->K[0], c->K[4], c->K[2], c->K[5], rot_grid, N,
                            reinterpret_cast<unsigned int*>(c->mask_stats.p), poses_out, info_out, st);
  FP_API_END
}

int fp_make_crops(fp_ctx* c, const float* poses, int N, int mode, void* crops_out, float* dbg_out, float* win_out,
                  void* stream) {
  FP_API_BEGIN
  FP_REQUIRE(c && poses && N >= 0, "fp_make_crops: bad argument");
  FP_REQUIRE(mode == 0 || mode == 1, "fp_make_crops: mode must be 0 (refiner) or 1 (scorer)");
  DeviceGuard dg(c->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (N == 0) return 0;
  FP_TRY(ensure_capacity(c, N));
  FP_TRY(make_crops(c, poses, N, mode, dbg_out, win_out, nullptr, st));
  if (crops_out) FP_TRY(crops_export(c, crops_out, N, st));
  return 0;
  FP_API_END
}

int fp_crop_stats(fp_ctx* c, const float* poses, int N, int mode, int* stats_out_host, void* stream) {
  FP_API_BEGIN
  FP_REQUIRE(c && poses && stats_out_host && N > 0, "fp_crop_stats: bad argument");
  DeviceGuard dg(c->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  FP_TRY(ensure_capacity(c, N));
  FP_TRY(dev_alloc(c->epoch, c->crop_stats, 16));
  FP_CUDA_OK(cudaMemsetAsync(c->crop_stats.p, 0, 16, st));
  FP_TRY(make_crops(c, poses, N, mode, nullptr, nullptr, reinterpret_cast<int*>(c->crop_stats.p), st));
  FP_CUDA_OK(cudaMemcpyAsync(stats_out_host, c->crop_stats.p, 16, cudaMemcpyDeviceToHost, st));
  FP_CUDA_OK(cudaStreamSynchronize(st));
  return 0;
  FP_API_END
}

int fp_op_refine_net(fp_ctx* c, const void* crops, int N, float* trans_out, float* rot_out, void* stream) {
  FP_API_BEGIN
  FP_REQUIRE(c && crops && trans_out && rot_out, "fp_op_refine_net: null argument");
  FP_REQUIRE(c->net[0].loaded, "refiner weights not loaded");
  DeviceGuard dg(c->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (N == 0) return 0;
  FP_TRY(ensure_capacity(c, N));
  FP_TRY(crops_import(c, crops, N, st));
  FP_TRY(run_encoder(c, c->net[0], reinterpret_cast<const __half*>(c->crops.p), N, st));
  FP_TRY(run_refine_heads(c, c->net[0], N, st));
  const float* ho = reinterpret_cast<const float*>(c->head_out.p);
  FP_CUDA_OK(cudaMemcpyAsync(trans_out, ho, (size_t)N * 12, cudaMemcpyDeviceToDevice, st));
  FP_CUDA_OK(cudaMemcpyAsync(rot_out, ho + (size_t)N * 3, (size_t)N * 12, cudaMemcpyDeviceToDevice, st));
  return 0;
  FP_API_END
}

int fp_op_score_feats(fp_ctx* c, const void* crops, int N, float* feats_out, void* stream) {
  FP_API_BEGIN
  FP_REQUIRE(c && crops && feats_out, "fp_op_score_feats: null argument");
  FP_REQUIRE(c->net[1].loaded, "scorer weights not loaded");
  DeviceGuard dg(c->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (N == 0) return 0;
  FP_TRY(ensure_capacity(c, N));
  FP_TRY(crops_import(c, crops, N, st));
  FP_TRY(run_encoder(c, c->net[1], reinterpret_cast<const __half*>(c->crops.p), N, st));
  FP_TRY(run_score_feats(c, c->net[1], N, feats_out, st));
  return 0;
  FP_API_END
}

int fp_op_tokens(fp_ctx* c, int which, const void* crops, int N, void* tokens_out, void* stream) {
  FP_API_BEGIN
  FP_REQUIRE(c && crops && tokens_out && (which == 0 || which == 1), "fp_op_tokens: bad argument");
  FP_REQUIRE(c->net[which].loaded, "weights not loaded");
  DeviceGuard dg(c->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (N == 0) return 0;
  FP_TRY(ensure_capacity(c, N));
  FP_TRY(crops_import(c, crops, N, st));
  FP_TRY(run_encoder(c, c->net[which], reinterpret_cast<const __half*>(c->crops.p), N, st));
  FP_CUDA_OK(cudaMemcpyAsync(tokens_out, c->tok.p, (size_t)N * T * 512 * 2, cudaMemcpyDeviceToDevice, st));
  return 0;
  FP_API_END
}

int fp_refine(fp_ctx* c, const float* poses_in, int N, int iterations, float* poses_out, float* last_trans,
              float* last_rot, void* stream) {
  FP_API_BEGIN
  FP_REQUIRE(c && poses_in && poses_out && N >= 0 && iterations >= 0, "fp_refine: bad argument");
  FP_REQUIRE(c->net[0].loaded, "refiner weights not loaded");
  FP_REQUIRE(c->has_mesh && c->has_frame, "fp_refine: needs fp_set_mesh and fp_set_frame first");
  DeviceGuard dg(c->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (N == 0) return 0;
  FP_TRY(ensure_capacity(c, N));
  float* pa = reinterpret_cast<float*>(c->poses_a.p);
  float* pb = reinterpret_cast<float*>(c->poses_b.p);
  FP_CUDA_OK(cudaMemcpyAsync(pa, poses_in, (size_t)N * 64, cudaMemcpyDeviceToDevice, st));
  FP_TRY(run_graphed(c, 0, N, iterations, st, [&](cudaStream_t s2) -> int { return refine_body(c, N, iterations, s2); }));
  const float* fin = (iterations % 2 == 0) ? pa : pb;
  FP_CUDA_OK(cudaMemcpyAsync(poses_out, fin, (size_t)N * 64, cudaMemcpyDeviceToDevice, st));
  if (iterations > 0) {
    if (last_trans) FP_CUDA_OK(cudaMemcpyAsync(last_trans, c->lt_buf.p, (size_t)N * 12, cudaMemcpyDeviceToDevice, st));
    if (last_rot) FP_CUDA_OK(cudaMemcpyAsync(last_rot, c->lr_buf.p, (size_t)N * 36, cudaMemcpyDeviceToDevice, st));
  }
  return 0;
  FP_API_END
}

int fp_score_features(fp_ctx* c, const float* poses, int N, float* feats_out, void* stream) {
  FP_API_BEGIN
  FP_REQUIRE(c && poses && feats_out && N >= 0, "fp_score_features: bad argument");
  FP_REQUIRE(c->net[1].loaded, "scorer weights not loaded");
  FP_REQUIRE(c->has_mesh && c->has_frame, "fp_score_features: needs fp_set_mesh and fp_set_frame first");
  DeviceGuard dg(c->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (N == 0) return 0;
  FP_TRY(ensure_capacity(c, N));
  float* ps = reinterpret_cast<float*>(c->pose_stage.p);
  float* fb = reinterpret_cast<float*>(c->feat_buf.p);
  FP_CUDA_OK(cudaMemcpyAsync(ps, poses, (size_t)N * 64, cudaMemcpyDeviceToDevice, st));
  auto body = [&](cudaStream_t s2) -> int {
    FP_TRY(make_crops(c, ps, N, 1, nullptr, nullptr, nullptr, s2));
    FP_TRY(run_encoder(c, c->net[1], reinterpret_cast<const __half*>(c->crops.p), N, s2));
    FP_TRY(run_score_feats(c, c->net[1], N, fb, s2));
    return 0;
  };
  FP_TRY(run_graphed(c, 1, N, 0, st, body));
  // feats_out may live on another GPU of the same process (peer access enabled by fp_group_create): the gather of
  // the sharded register is this copy, device to device over NVLink
  FP_CUDA_OK(cudaMemcpyAsync(feats_out, fb, (size_t)N * 2048, cudaMemcpyDefault, st));
  return 0;
  FP_API_END
}

int fp_score_tail(fp_ctx* c, const float* feats, int L, float* scores_out, int* best_out, void* stream) {
  FP_API_BEGIN
  FP_REQUIRE(c && feats && scores_out && L >= 0, "fp_score_tail: bad argument");
  FP_REQUIRE(c->net[1].loaded, "scorer weights not loaded");
  DeviceGuard dg(c->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (L == 0) return 0;
  FP_TRY(ensure_tail(c, L));
  const Net& net = c->net[1];
  ScoreTailParams p;
  p.feats = feats;
  p.L = L;
  p.w_in = net.f("cross.in_w");
  p.b_in = net.f("cross.in_b");
  p.fold_v = reinterpret_cast<const float*>(c->fold_v.p);
  p.fold_c = c->fold_c;
  p.offset = 100.f;
  p.qkv = reinterpret_cast<float*>(c->tail_qkv.p);
  p.scores = scores_out;
  p.best = best_out;
  p.counter = reinterpret_cast<unsigned int*>(c->tail_counter.p);
  return score_tail_launch(p, st);
  FP_API_END
}

int fp_score(fp_ctx* c, const float* poses, int N, float* scores_out, int* best_out, void* stream) {
  FP_API_BEGIN
  FP_REQUIRE(c && poses && scores_out && N >= 0, "fp_score: bad argument");
  if (N == 0) return 0;
  DeviceGuard dg(c->device);
  FP_TRY(ensure_capacity(c, N));
  FP_TRY(fp_score_features(c, poses, N, reinterpret_cast<float*>(c->feats.p), stream));
  return fp_score_tail(c, reinterpret_cast<const float*>(c->feats.p), N, scores_out, best_out, stream);
  FP_API_END
}

int fp_register(fp_ctx* c, const float* poses_host, int N, int iterations, float* poses_out_host, float* scores_out_host,
                int* best_out_host, void* stream) {
  FP_API_BEGIN
  FP_REQUIRE(c && poses_host && poses_out_host && scores_out_host && best_out_host && N > 0, "fp_register: bad argument");
  DeviceGuard dg(c->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  FP_TRY(ensure_capacity(c, N));
  FP_TRY(ensure_tail(c, N));
  // poses_b is the loop's ping-pong partner; stage the input in `feats`' neighbour: use tail_proj as scratch
  float* pin = reinterpret_cast<float*>(c->tail_proj.p);  // >= N*512 floats
  float* pout = pin + (size_t)N * 16;
  FP_CUDA_OK(cudaMemcpyAsync(pin, poses_host, (size_t)N * 64, cudaMemcpyHostToDevice, st));
  FP_TRY(fp_refine(c, pin, N, iterations, pout, nullptr, nullptr, stream));
  // keep the refined poses out of the tail's scratch: copy to poses_b's idle half (poses_a/b are free now)
  float* refined = reinterpret_cast<float*>(c->poses_b.p);
  FP_CUDA_OK(cudaMemcpyAsync(refined, pout, (size_t)N * 64, cudaMemcpyDeviceToDevice, st));
  FP_TRY(fp_score(c, refined, N, reinterpret_cast<float*>(c->scores.p), reinterpret_cast<int*>(c->best.p), stream));
  FP_CUDA_OK(cudaMemcpyAsync(poses_out_host, refined, (size_t)N * 64, cudaMemcpyDeviceToHost, st));
  FP_CUDA_OK(cudaMemcpyAsync(scores_out_host, c->scores.p, (size_t)N * 4, cudaMemcpyDeviceToHost, st));
  FP_CUDA_OK(cudaMemcpyAsync(best_out_host, c->best.p, 4, cudaMemcpyDeviceToHost, st));
  FP_CUDA_OK(cudaStreamSynchronize(st));
  return 0;
  FP_API_END
}

int fp_track(fp_ctx* c, const unsigned char* rgb_host, const float* depth_host, const float* K, int H, int W,
             const float* pose_in_dev, int iterations, float* pose_out_dev, float* pose_out_host, void* stream) {
  FP_API_BEGIN
  FP_REQUIRE(c && rgb_host && depth_host && K && H > 0 && W > 0 && iterations >= 0, "fp_track: bad argument");
  FP_REQUIRE(c->net[0].loaded, "refiner weights not loaded");
  FP_REQUIRE(c->has_mesh, "fp_track: no mesh");
  DeviceGuard dg(c->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t npix = (size_t)H * W;
  FP_TRY(ensure_capacity(c, 1));
  FP_TRY(prepare_frame(c, K, H, W, true));
  FP_TRY(dev_alloc(c->epoch, c->track_pose, 64));
  if (c->stage_npix < npix) {
    if (c->stage_rgb) cudaFreeHost(c->stage_rgb);
    if (c->stage_depth) cudaFreeHost(c->stage_depth);
    c->stage_rgb = c->stage_depth = nullptr;
    c->stage_npix = 0;
    FP_CUDA_OK(cudaMallocHost(&c->stage_rgb, npix * 3));
    FP_CUDA_OK(cudaMallocHost(&c->stage_depth, npix * 4));
    c->stage_npix = npix;
    ++c->epoch;  // the graph's copy nodes hold these addresses
  }
  if (!c->stage_pose) FP_CUDA_OK(cudaMallocHost(reinterpret_cast<void**>(&c->stage_pose), 64));
  if (pose_in_dev) {
    FP_CUDA_OK(cudaMemcpyAsync(c->track_pose.p, pose_in_dev, 64, cudaMemcpyDeviceToDevice, st));
  } else {
    FP_REQUIRE(c->track_valid, "fp_track: no previous pose in this context: pass pose_in");
  }
  // the previous frame's graph has finished (fp_track synchronises), so the staging buffers are free.  The two
  // uploads are issued as soon as their staging copy is done — the depth DMA runs under the host's rgb copy, the
  // rgb DMA under the graph launch — instead of being nodes of the graph (measured: -40 us per frame)
  memcpy(c->stage_depth, depth_host, npix * 4);
  FP_CUDA_OK(cudaMemcpyAsync(c->depth_raw.p, c->stage_depth, npix * 4, cudaMemcpyHostToDevice, st));
  memcpy(c->stage_rgb, rgb_host, npix * 3);
  FP_CUDA_OK(cudaMemcpyAsync(c->rgb_raw.p, c->stage_rgb, npix * 3, cudaMemcpyHostToDevice, st));
  c->has_frame = false;
  float* pa = reinterpret_cast<float*>(c->poses_a.p);
  float* pb = reinterpret_cast<float*>(c->poses_b.p);
  auto body = [&](cudaStream_t s2) -> int {
    // estimater.py:250-268 in one launch sequence: erode + bilateral, depth2xyzmap_batch(zfar=inf), K refiner passes
    FP_TRY(set_frame_launches(c, reinterpret_cast<const unsigned char*>(c->rgb_raw.p),
                              reinterpret_cast<const float*>(c->depth_raw.p), FP_FRAME_FILTER_DEPTH, INFINITY, s2));
    FP_CUDA_OK(cudaMemcpyAsync(pa, c->track_pose.p, 64, cudaMemcpyDeviceToDevice, s2));
    c->has_frame = true;
    FP_TRY(refine_body(c, 1, iterations, s2));
    const float* fin = (iterations % 2 == 0) ? pa : pb;
    FP_CUDA_OK(cudaMemcpyAsync(c->track_pose.p, fin, 64, cudaMemcpyDeviceToDevice, s2));
    FP_CUDA_OK(cudaMemcpyAsync(c->stage_pose, fin, 64, cudaMemcpyDeviceToHost, s2));
    return 0;
  };
  FP_TRY(run_graphed(c, 2, 1, iterations, st, body));
  c->has_frame = true;
  c->track_valid = true;
  if (pose_out_dev) FP_CUDA_OK(cudaMemcpyAsync(pose_out_dev, c->track_pose.p, 64, cudaMemcpyDeviceToDevice, st));
  FP_CUDA_OK(cudaStreamSynchronize(st));
  if (pose_out_host) memcpy(pose_out_host, c->stage_pose, 64);
  return 0;
  FP_API_END
}

int fp_op_depth_filter(const float* depth_dev, float* out_dev, int H, int W, int which, void* stream) {
  FP_API_BEGIN
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  FP_REQUIRE(depth_dev && out_dev && H > 0 && W > 0, "fp_op_depth_filter: bad argument");
  if (which == 0) return erode_depth_launch(depth_dev, out_dev, H, W, 2, 0.001f, 0.8f, 100.f, st);
  return bilateral_depth_launch(depth_dev, out_dev, H, W, 2, 100.f, 2.f, 100000.f, st);
  FP_API_END
}

int fp_op_pose_update(const float* poses_in, const float* trans, const float* rot, float* poses_out, int N,
                      float mesh_diameter, float rot_normalizer, void* stream) {
  FP_API_BEGIN
  FP_REQUIRE(poses_in && trans && rot && poses_out, "fp_op_pose_update: null argument");
  return pose_update_launch(poses_in, trans, rot, poses_out, nullptr, nullptr, N, mesh_diameter / 2.0f, rot_normalizer,
                            reinterpret_cast<cudaStream_t>(stream));
  FP_API_END
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// fp_group: ONE process (one host thread) driving several GPUs — the reference's process model (run_demo.py is one
// script).  One fp_ctx per device, each with its own stream; register() shards the hypothesis list contiguously,
// every device filters the frame, derives the start poses and refines / featurises its slice; the only exchange is
// the per-hypothesis feature rows (+ refined poses), written by each device DIRECTLY into device 0's gather buffer
// over NVLink peer memory (cudaMemcpyAsync device-to-device on the producing device's stream: no host staging, no
// collective library); device 0 waits on one event per peer and runs the cross-hypothesis tail once.
// ------------------------------------------------------------------------------------------------
struct fp_group {
  std::vector<fp_ctx*> ctx;
  std::vector<cudaStream_t> stream;
  std::vector<cudaEvent_t> done;
  std::vector<fp::DevBuf> grid, start, info, refined;  // per device: rot grid [N][16], start poses, info[4], refined slice
  fp::DevBuf feats_all, poses_all, scores, best;       // on device 0
  void* pin_rgb = nullptr;
  void* pin_depth = nullptr;
  void* pin_mask = nullptr;
  void* pin_grid = nullptr;
  size_t pin_npix = 0, pin_grid_n = 0;
  unsigned long long epoch = 0;
};

extern "C" {

int fp_group_destroy(fp_group* g) {
  FP_API_BEGIN
  if (!g) return 0;
  for (size_t i = 0; i < g->ctx.size(); ++i) {
    DeviceGuard dg(g->ctx[i]->device);
    cudaDeviceSynchronize();
    for (fp::DevBuf* b : {&g->grid[i], &g->start[i], &g->info[i], &g->refined[i]})
      if (b->p) cudaFree(b->p);
    if (i == 0)
      for (fp::DevBuf* b : {&g->feats_all, &g->poses_all, &g->scores, &g->best})
        if (b->p) cudaFree(b->p);
    if (g->stream[i]) cudaStreamDestroy(g->stream[i]);
    if (g->done[i]) cudaEventDestroy(g->done[i]);
    fp_destroy(g->ctx[i]);
  }
  for (void* p : {g->pin_rgb, g->pin_depth, g->pin_mask, g->pin_grid})
    if (p) cudaFreeHost(p);
  delete g;
  return 0;
  FP_API_END
}

int fp_group_create(int ndev, const int* dev_ids, fp_group** out) {
  FP_API_BEGIN
  FP_REQUIRE(out && ndev > 0, "fp_group_create: bad argument");
  int visible = 0;
  FP_CUDA_OK(cudaGetDeviceCount(&visible));
  fp_group* g = new fp_group();
  int prev = 0;
  cudaGetDevice(&prev);
  for (int i = 0; i < ndev; ++i) {
    const int dev = dev_ids ? dev_ids[i] : i;
    if (dev < 0 || dev >= visible) {
      fp_group_destroy(g);
      set_last_error("fp_group_create: device %d not visible (%d devices)", dev, visible);
      cudaSetDevice(prev);
      return -1;
    }
    cudaSetDevice(dev);
    fp_ctx* c = nullptr;
    const int rc = fp_create(&c);
    if (rc) {
      fp_group_destroy(g);
      cudaSetDevice(prev);
      return rc;
    }
    cudaStream_t s = nullptr;
    cudaEvent_t e = nullptr;
    cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    g->ctx.push_back(c);
    g->stream.push_back(s);
    g->done.push_back(e);
    g->grid.emplace_back();
    g->start.emplace_back();
    g->info.emplace_back();
    g->refined.emplace_back();
    if (i > 0) {
      // peers write their feature rows into device 0's buffer: map device 0's memory into this device
      int can = 0;
      cudaDeviceCanAccessPeer(&can, dev, g->ctx[0]->device);
      if (can) {
        const cudaError_t pe = cudaDeviceEnablePeerAccess(g->ctx[0]->device, 0);
        if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) {
          set_last_error("fp_group_create: cudaDeviceEnablePeerAccess(%d -> %d): %s", dev, g->ctx[0]->device, cudaGetErrorString(pe));
          fp_group_destroy(g);
          cudaSetDevice(prev);
          return -2;
        }
        cudaGetLastError();
      }
    }
  }
  cudaSetDevice(prev);
  *out = g;
  return 0;
  FP_API_END
}

int fp_group_size(fp_group* g) { return g ? (int)g->ctx.size() : 0; }

fp_ctx* fp_group_ctx(fp_group* g, int i) { return (g && i >= 0 && i < (int)g->ctx.size()) ? g->ctx[i] : nullptr; }

int fp_group_load_network(fp_group* g, int which, const fp_tensor_t* tensors, int n) {
  FP_API_BEGIN
  FP_REQUIRE(g, "fp_group_load_network: null group");
  for (fp_ctx* c : g->ctx) FP_TRY(fp_load_network(c, which, tensors, n));
  return 0;
  FP_API_END
}

int fp_group_set_config(fp_group* g, int which, float crop_ratio, float rot_normalizer) {
  FP_API_BEGIN
  FP_REQUIRE(g, "fp_group_set_config: null group");
  for (fp_ctx* c : g->ctx) FP_TRY(fp_set_config(c, which, crop_ratio, rot_normalizer));
  return 0;
  FP_API_END
}

int fp_group_set_mesh(fp_group* g, int V, int F, const float* pos, const float* nrm, const float* uv, const float* vcol,
                      const int* faces, const unsigned char* tex_rgb, int Ht, int Wt, float diameter) {
  FP_API_BEGIN
  FP_REQUIRE(g, "fp_group_set_mesh: null group");
  for (fp_ctx* c : g->ctx) FP_TRY(fp_set_mesh(c, V, F, pos, nrm, uv, vcol, faces, tex_rgb, Ht, Wt, diameter));
  return 0;
  FP_API_END
}

int fp_group_register(fp_group* g, const unsigned char* rgb_host, const float* depth_host, const float* K, int H, int W,
                      const unsigned char* mask_host, const float* rot_grid_host, int N, int iterations,
                      float* poses_out_host, float* scores_out_host, int* best_out_host, float* info_out_host) {
  FP_API_BEGIN
  FP_REQUIRE(g && rgb_host && depth_host && K && mask_host && rot_grid_host && poses_out_host && scores_out_host &&
                 best_out_host && N > 0 && H > 0 && W > 0,
             "fp_group_register: bad argument");
  const int G = (int)g->ctx.size();
  const size_t npix = (size_t)H * W;
  // pinned staging, filled once, read by every device
  if (g->pin_npix < npix) {
    for (void** p : {&g->pin_rgb, &g->pin_depth, &g->pin_mask}) {
      if (*p) cudaFreeHost(*p);
      *p = nullptr;
    }
    g->pin_npix = 0;
    FP_CUDA_OK(cudaHostAlloc(&g->pin_rgb, npix * 3, cudaHostAllocPortable));
    FP_CUDA_OK(cudaHostAlloc(&g->pin_depth, npix * 4, cudaHostAllocPortable));
    FP_CUDA_OK(cudaHostAlloc(&g->pin_mask, npix, cudaHostAllocPortable));
    g->pin_npix = npix;
  }
  if (g->pin_grid_n < (size_t)N) {
    if (g->pin_grid) cudaFreeHost(g->pin_grid);
    g->pin_grid = nullptr;
    g->pin_grid_n = 0;
    FP_CUDA_OK(cudaHostAlloc(&g->pin_grid, (size_t)N * 64, cudaHostAllocPortable));
    g->pin_grid_n = N;
  }
  memcpy(g->pin_rgb, rgb_host, npix * 3);
  memcpy(g->pin_depth, depth_host, npix * 4);
  memcpy(g->pin_mask, mask_host, npix);
  memcpy(g->pin_grid, rot_grid_host, (size_t)N * 64);
  fp_ctx* c0 = g->ctx[0];
  {
    DeviceGuard dg(c0->device);
    FP_TRY(dev_alloc(g->epoch, g->feats_all, (size_t)N * 2048));
    FP_TRY(dev_alloc(g->epoch, g->poses_all, (size_t)N * 64));
    FP_TRY(dev_alloc(g->epoch, g->scores, (size_t)N * 4));
    FP_TRY(dev_alloc(g->epoch, g->best, 16));
  }
  const int base = N / G, rem = N % G;
  for (int i = 0; i < G; ++i) {
    fp_ctx* c = g->ctx[i];
    DeviceGuard dg(c->device);
    cudaStream_t st = g->stream[i];
    const int lo = i * base + (i < rem ? i : rem), n = base + (i < rem ? 1 : 0);
    FP_TRY(dev_alloc(g->epoch, g->grid[i], (size_t)N * 64));
    FP_TRY(dev_alloc(g->epoch, g->start[i], (size_t)N * 64));
    FP_TRY(dev_alloc(g->epoch, g->info[i], 16));
    FP_TRY(dev_alloc(g->epoch, g->refined[i], (size_t)(n > 0 ? n : 1) * 64));
    FP_TRY(fp_set_frame(c, reinterpret_cast<const unsigned char*>(g->pin_rgb), reinterpret_cast<const float*>(g->pin_depth),
                        K, H, W, FP_FRAME_FILTER_DEPTH, INFINITY, st));
    FP_CUDA_OK(cudaMemcpyAsync(g->grid[i].p, g->pin_grid, (size_t)N * 64, cudaMemcpyHostToDevice, st));
    FP_TRY(fp_start_poses(c, reinterpret_cast<const unsigned char*>(g->pin_mask), 0, reinterpret_cast<const float*>(g->grid[i].p),
                          N, reinterpret_cast<float*>(g->start[i].p), reinterpret_cast<float*>(g->info[i].p), st));
    if (n > 0) {
      float* refined = reinterpret_cast<float*>(g->refined[i].p);
      FP_TRY(fp_refine(c, reinterpret_cast<const float*>(g->start[i].p) + (size_t)lo * 16, n, iterations, refined, nullptr,
                       nullptr, st));
      // the gather: feature rows and refined poses land in device 0's buffers, straight over peer memory
      FP_TRY(fp_score_features(c, refined, n, reinterpret_cast<float*>(g->feats_all.p) + (size_t)lo * 512, st));
      FP_CUDA_OK(cudaMemcpyAsync(reinterpret_cast<float*>(g->poses_all.p) + (size_t)lo * 16, refined, (size_t)n * 64,
                                 cudaMemcpyDefault, st));
    }
    FP_CUDA_OK(cudaEventRecord(g->done[i], st));
  }
  {
    DeviceGuard dg(c0->device);
    cudaStream_t s0 = g->stream[0];
    for (int i = 1; i < G; ++i) FP_CUDA_OK(cudaStreamWaitEvent(s0, g->done[i], 0));
    FP_TRY(fp_score_tail(c0, reinterpret_cast<const float*>(g->feats_all.p), N, reinterpret_cast<float*>(g->scores.p),
                         reinterpret_cast<int*>(g->best.p), s0));
    FP_CUDA_OK(cudaMemcpyAsync(poses_out_host, g->poses_all.p, (size_t)N * 64, cudaMemcpyDeviceToHost, s0));
    FP_CUDA_OK(cudaMemcpyAsync(scores_out_host, g->scores.p, (size_t)N * 4, cudaMemcpyDeviceToHost, s0));
    FP_CUDA_OK(cudaMemcpyAsync(best_out_host, g->best.p, 4, cudaMemcpyDeviceToHost, s0));
    if (info_out_host) FP_CUDA_OK(cudaMemcpyAsync(info_out_host, g->info[0].p, 16, cudaMemcpyDeviceToHost, s0));
    FP_CUDA_OK(cudaStreamSynchronize(s0));
  }
  return 0;
  FP_API_END
}

}  // extern "C"
