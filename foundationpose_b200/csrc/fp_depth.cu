// fp_depth.cu — per-frame depth pre-processing (once per register()/track_one() call).
//   erode_depth_kernel      restates Utils.py:359-384 (Warp kernel `erode_depth_kernel`)
//   bilateral_depth_kernel  restates Utils.py:304-343 (Warp kernel `bilateral_filter_depth_kernel`)
// 5x5 stencils over a 480x640 fp32 image: HBM-bound, 2 x 1.2 MB of traffic each; one thread per pixel,
// rows of a warp are contiguous so global accesses coalesce and the stencil reuse is served by L1.
#include "fp_depth.cuh"

#include "fp_common.cuh"
#include "fp_gemm.cuh"

namespace fp {

__global__ void erode_depth_kernel(const float* __restrict__ depth, float* __restrict__ out, int H, int W, int radius,
                                   float diff_thres, float ratio_thres, float zfar) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  const int h = blockIdx.y * blockDim.y + threadIdx.y;
  if (w >= W || h >= H) return;
  const float d_ori = depth[h * W + w];
  // NB: the reference writes 0 for an invalid centre and then *falls through* (Utils.py:366-384);
  // the final store below decides the value, exactly as there.
  float bad = 0.f, total = 0.f;
  for (int u = w - radius; u <= w + radius; ++u) {
    if (u < 0 || u >= W) continue;
    for (int v = h - radius; v <= h + radius; ++v) {
      if (v < 0 || v >= H) continue;
      const float cur = depth[v * W + u];
      total += 1.f;
      if (cur < 0.001f || cur >= zfar || fabsf(cur - d_ori) > diff_thres) bad += 1.f;
    }
  }
  out[h * W + w] = (bad / total > ratio_thres) ? 0.f : d_ori;
}

__global__ void bilateral_depth_kernel(const float* __restrict__ depth, float* __restrict__ out, int H, int W,
                                       int radius, float zfar, float sigmaD, float sigmaR) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  const int h = blockIdx.y * blockDim.y + threadIdx.y;
  if (w >= W || h >= H) return;
  float mean_depth = 0.f;
  int num_valid = 0;
  for (int u = w - radius; u <= w + radius; ++u) {
    if (u < 0 || u >= W) continue;
    for (int v = h - radius; v <= h + radius; ++v) {
      if (v < 0 || v >= H) continue;
      const float cur = depth[v * W + u];
      if (cur >= 0.001f && cur < zfar) {
        ++num_valid;
        mean_depth += cur;
      }
    }
  }
  float result = 0.f;
  if (num_valid > 0) {
    mean_depth /= (float)num_valid;
    const float dc = depth[h * W + w];
    float sum_w = 0.f, sum = 0.f;
    for (int u = w - radius; u <= w + radius; ++u) {
      if (u < 0 || u >= W) continue;
      for (int v = h - radius; v <= h + radius; ++v) {
        if (v < 0 || v >= H) continue;
        const float cur = depth[v * W + u];
        if (cur >= 0.001f && cur < zfar && fabsf(cur - mean_depth) < 0.01f) {
          const float wgt = expf(-(float)((u - w) * (u - w) + (h - v) * (h - v)) / (2.f * sigmaD * sigmaD) -
                                 (dc - cur) * (dc - cur) / (2.f * sigmaR * sigmaR));
          sum_w += wgt;
          sum += wgt * cur;
        }
      }
    }
    if (sum_w > 0.f) result = sum / sum_w;
  }
  out[h * W + w] = result;
}

int erode_depth_launch(const float* depth, float* out, int H, int W, int radius, float diff_thres, float ratio_thres,
                       float zfar, cudaStream_t stream) {
  dim3 block(32, 8), grid((W + 31) / 32, (H + 7) / 8);
  erode_depth_kernel<<<grid, block, 0, stream>>>(depth, out, H, W, radius, diff_thres, ratio_thres, zfar);
  ++g_launch_count;
  FP_CUDA_OK(cudaGetLastError());
  return 0;
}

int bilateral_depth_launch(const float* depth, float* out, int H, int W, int radius, float zfar, float sigmaD,
                           float sigmaR, cudaStream_t stream) {
  dim3 block(32, 8), grid((W + 31) / 32, (H + 7) / 8);
  bilateral_depth_kernel<<<grid, block, 0, stream>>>(depth, out, H, W, radius, zfar, sigmaD, sigmaR);
  ++g_launch_count;
  FP_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace fp
