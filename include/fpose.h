/* fpose.h — C ABI of libfpose.so, the B200-native (sm_100a) render-and-compare hot path behind
 * NVlabs/FoundationPose's Python surfaces.
 *
 * Every entry point is `extern "C"`, takes plain pointers and sizes, returns 0 on success or a
 * negative error code (fp_last_error() returns a thread-local message); nothing throws or aborts.
 * All `void* stream` arguments are a cudaStream_t (pass torch.cuda.current_stream().cuda_stream);
 * all work is enqueued on that stream and is asynchronous unless stated otherwise.
 * Device pointers are owned by the caller; the library owns only what it allocates inside an
 * fp_ctx (packed weights, mesh copy, frame copy, workspaces).
 *
 * Each declaration cites the reference interface it replaces (paths relative to the
 * FoundationPose repository root).
 */
#ifndef FPOSE_H_
#define FPOSE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------ */
/* errors / counters                                                                          */
/* ------------------------------------------------------------------------------------------ */
const char* fp_last_error(void);
/* Number of CUDA kernels this library has launched so far in this process (bench.py's
 * `gpu_launches`). */
unsigned long long fp_launch_count(void);

/* ------------------------------------------------------------------------------------------ */
/* single operators (parity-test hooks; the product path below calls the same code)           */
/* ------------------------------------------------------------------------------------------ */

/* kinds of dense layer the tcgen05 implicit-GEMM kernel executes */
#define FP_LAYER_LINEAR 0   /* torch.nn.Linear / in_proj / out_proj (refine_network.py:56-70)      */
#define FP_LAYER_CONV3_S1 1 /* 3x3 s1 p1 conv of ResnetBasicBlock (network_modules.py:73-111)       */
#define FP_LAYER_CONV3_S2 2 /* 3x3 s2 p1 ConvBNReLU (refine_network.py:37, :45)                      */
#define FP_LAYER_CONV7_S2 3 /* 7x7 s2 p3 stem ConvBNReLU (refine_network.py:36)                      */

typedef struct fp_gemm_layer {
  int kind;
  int n_img;          /* images in the batch (LINEAR: 1)                                          */
  int Hin, Win;       /* un-padded input size (LINEAR: Hin = 1, Win = number of rows M)           */
  int Cin;            /* input channels (LINEAR: K; CONV7_S2: 8 = 6 real + 2 zero)                */
  int Cout;           /* output channels, multiple of 64                                          */
  const void* in;     /* fp16; NHWC.  CONV7_S2: [n][Hin+6][Win+8][8] with image at offset (3,3)
                         and a zero border                                                        */
  const void* w;      /* fp16 [Cout][taps*Cin] (tap-major, channel-minor); CONV7_S2: [Cout][7][64]
                         = per filter row 7 taps x 8 ch + 8 zeros                                 */
  const float* bias;  /* fp32 [Cout] (BatchNorm folded in)                                        */
  const void* res;    /* optional fp16 residual, same indexing as the output, leading dim res_ld  */
  int res_ld;
  void* out;          /* fp16 NHWC output                                                         */
  int out_ld;         /* elements between consecutive output pixels                               */
  int out_split;      /* >0: image n goes to image n % out_split at channel (n / out_split)*Cout  */
  const float* post_add; /* optional fp32 [Ho*Wo][Cout], added after the activation               */
  int relu;
} fp_gemm_layer_t;

/* Runs one layer: out = act(in (*) w + bias [+ res]) [+ post_add]. */
int fp_op_gemm_layer(const fp_gemm_layer_t* layer, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FPOSE_H_ */
