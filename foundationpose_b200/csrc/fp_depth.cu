// fp_depth.cu — per-frame depth pre-processing (once per register()/track_one() call).
//   erode_depth_kernel      restates Utils.py:359-384 (Warp kernel `erode_depth_kernel`)
//   bilateral_depth_kernel  restates Utils.py:304-343 (Warp kernel `bilateral_filter_depth_kernel`)
//   frame_prep_kernel       the two filters + depth2xyzmap + the rgb repack fused into one launch per frame
// 5x5 stencils over a 480x640 fp32 image: one thread per pixel; the stand-alone kernels (operator hooks, parity tests)
// read through L1, the fused kernel from a shared-memory tile.
#include "fp_depth.cuh"

#include "fp_common.cuh"
#include "fp_gemm.cuh"

namespace fp {

// Per-pixel bodies, shared by the stand-alone kernels (global-memory source) and the fused frame kernel (shared-memory
// tile source): the loops run column-outer / row-inner like the Warp kernels, so the float sums are formed in the
// reference's order whichever source is used.
struct GlobalSrc {
  const float* p;
  int W;
  __device__ __forceinline__ float operator()(int v, int u) const { return p[v * W + u]; }
};
struct TileSrc {  // rows [v0, ...), columns [u0, ...) of the image, `pitch` floats per row
  const float* s;
  int v0, u0, pitch;
  __device__ __forceinline__ float operator()(int v, int u) const { return s[(v - v0) * pitch + (u - u0)]; }
};

template <class Src>
__device__ __forceinline__ float erode_px(const Src& src, int w, int h, int H, int W, int radius, float diff_thres,
                                          float ratio_thres, float zfar) {
  const float d_ori = src(h, w);
  // NB: the reference writes 0 for an invalid centre and then *falls through* (Utils.py:366-384);
  // the final select below decides the value, exactly as there.
  float bad = 0.f, total = 0.f;
  for (int u = w - radius; u <= w + radius; ++u) {
    if (u < 0 || u >= W) continue;
    for (int v = h - radius; v <= h + radius; ++v) {
      if (v < 0 || v >= H) continue;
      const float cur = src(v, u);
      total += 1.f;
      if (cur < 0.001f || cur >= zfar || fabsf(cur - d_ori) > diff_thres) bad += 1.f;
    }
  }
  return (bad / total > ratio_thres) ? 0.f : d_ori;
}

template <class Src>
__device__ __forceinline__ float bilateral_px(const Src& src, int w, int h, int H, int W, int radius, float zfar,
                                              float sigmaD, float sigmaR) {
  float mean_depth = 0.f;
  int num_valid = 0;
  for (int u = w - radius; u <= w + radius; ++u) {
    if (u < 0 || u >= W) continue;
    for (int v = h - radius; v <= h + radius; ++v) {
      if (v < 0 || v >= H) continue;
      const float cur = src(v, u);
      if (cur >= 0.001f && cur < zfar) {
        ++num_valid;
        mean_depth += cur;
      }
    }
  }
  float result = 0.f;
  if (num_valid > 0) {
    mean_depth /= (float)num_valid;
    const float dc = src(h, w);
    float sum_w = 0.f, sum = 0.f;
    for (int u = w - radius; u <= w + radius; ++u) {
      if (u < 0 || u >= W) continue;
      for (int v = h - radius; v <= h + radius; ++v) {
        if (v < 0 || v >= H) continue;
        const float cur = src(v, u);
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
  return result;
}

__global__ void erode_depth_kernel(const float* __restrict__ depth, float* __restrict__ out, int H, int W, int radius,
                                   float diff_thres, float ratio_thres, float zfar) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  const int h = blockIdx.y * blockDim.y + threadIdx.y;
  if (w >= W || h >= H) return;
  out[h * W + w] = erode_px(GlobalSrc{depth, W}, w, h, H, W, radius, diff_thres, ratio_thres, zfar);
}

__global__ void bilateral_depth_kernel(const float* __restrict__ depth, float* __restrict__ out, int H, int W,
                                       int radius, float zfar, float sigmaD, float sigmaR) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  const int h = blockIdx.y * blockDim.y + threadIdx.y;
  if (w >= W || h >= H) return;
  out[h * W + w] = bilateral_px(GlobalSrc{depth, W}, w, h, H, W, radius, zfar, sigmaD, sigmaR);
}

// The whole per-frame preparation of register() / track_one() in ONE launch (estimater.py:173-174, :214 / :258-262):
// erode (radius 2) -> bilateral (radius 2) -> back-projection, plus the rgb -> rgba repack, per 32 x 8-pixel block.
// The block stages its raw depth tile with a 4-pixel halo in shared memory, erodes the tile with a 2-pixel halo into
// a second shared tile (halo pixels are recomputed by the neighbouring blocks, identically) and filters from there:
// one read of the depth image instead of ~50 per pixel from L1 / L2, three launches fewer per frame (47 -> ~12 us of
// a 0.97 ms tracked frame).  Same per-pixel code as the stand-alone kernels above: bit-identical.
constexpr int kFpW = 32, kFpH = 8, kFpR = 2;
__global__ void __launch_bounds__(kFpW* kFpH)
    frame_prep_kernel(const unsigned char* __restrict__ rgb, const float* __restrict__ depth, uchar4* __restrict__ rgba,
                      float* __restrict__ depth_out, float4* __restrict__ xyz, int H, int W, float fx, float fy, float cx,
                      float cy, float zfar_xyz) {
  constexpr int RW = kFpW + 4 * kFpR, RH = kFpH + 4 * kFpR;  // raw tile 40 x 16
  constexpr int EW = kFpW + 2 * kFpR, EH = kFpH + 2 * kFpR;  // eroded tile 36 x 12
  __shared__ float raw[RH * RW];
  __shared__ float er[EH * EW];
  const int w0 = blockIdx.x * kFpW, h0 = blockIdx.y * kFpH;
  const int tid = threadIdx.y * kFpW + threadIdx.x;
  for (int i = tid; i < RH * RW; i += kFpW * kFpH) {
    const int v = h0 - 2 * kFpR + i / RW, u = w0 - 2 * kFpR + i % RW;
    raw[i] = (v >= 0 && v < H && u >= 0 && u < W) ? depth[v * W + u] : 0.f;  // out-of-image cells are never read
  }
  __syncthreads();
  const TileSrc rsrc{raw, h0 - 2 * kFpR, w0 - 2 * kFpR, RW};
  for (int i = tid; i < EH * EW; i += kFpW * kFpH) {
    const int v = h0 - kFpR + i / EW, u = w0 - kFpR + i % EW;
    er[i] = (v >= 0 && v < H && u >= 0 && u < W) ? erode_px(rsrc, u, v, H, W, kFpR, 0.001f, 0.8f, 100.f) : 0.f;
  }
  __syncthreads();
  const int w = w0 + threadIdx.x, h = h0 + threadIdx.y;
  if (w >= W || h >= H) return;
  const float z = bilateral_px(TileSrc{er, h0 - kFpR, w0 - kFpR, EW}, w, h, H, W, kFpR, 100.f, 2.f, 100000.f);
  const int i = h * W + w;
  depth_out[i] = z;
  float X = 0.f, Y = 0.f, Z = 0.f;
  if (!(z < 0.001f) && !(z > zfar_xyz)) {  // depth2xyzmap(_batch): Utils.py:399-438
    X = ((float)w - cx) * z / fx;
    Y = ((float)h - cy) * z / fy;
    Z = z;
  }
  xyz[i] = make_float4(X, Y, Z, 0.f);
  rgba[i] = make_uchar4(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2], 255);
}

int frame_prep_launch(const unsigned char* rgb, const float* depth, uchar4* rgba, float* depth_out, float4* xyz, int H, int W,
                      float fx, float fy, float cx, float cy, float zfar_xyz, cudaStream_t stream) {
  dim3 block(kFpW, kFpH), grid((W + kFpW - 1) / kFpW, (H + kFpH - 1) / kFpH);
  frame_prep_kernel<<<grid, block, 0, stream>>>(rgb, depth, rgba, depth_out, xyz, H, W, fx, fy, cx, cy, zfar_xyz);
  note_launches(1);
  FP_CUDA_OK(cudaGetLastError());
  return 0;
}

int erode_depth_launch(const float* depth, float* out, int H, int W, int radius, float diff_thres, float ratio_thres,
                       float zfar, cudaStream_t stream) {
  dim3 block(32, 8), grid((W + 31) / 32, (H + 7) / 8);
  erode_depth_kernel<<<grid, block, 0, stream>>>(depth, out, H, W, radius, diff_thres, ratio_thres, zfar);
  note_launches(1);
  FP_CUDA_OK(cudaGetLastError());
  return 0;
}

int bilateral_depth_launch(const float* depth, float* out, int H, int W, int radius, float zfar, float sigmaD,
                           float sigmaR, cudaStream_t stream) {
  dim3 block(32, 8), grid((W + 31) / 32, (H + 7) / 8);
  bilateral_depth_kernel<<<grid, block, 0, stream>>>(depth, out, H, W, radius, zfar, sigmaD, sigmaR);
  note_launches(1);
  FP_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace fp

// ------------------------------------------------------------------------------------------------
// start poses on the device: guess_translation (estimater.py:137-156) + rot_grid with that translation
// (estimater.py:127-134, :203-209).  The median of the masked valid depths is an exact radix
// select over the float bit patterns (depths are positive, so the patterns are ordered), np.median's
// mean of the two middle elements for an even count included.
// ------------------------------------------------------------------------------------------------
namespace fp {

__device__ unsigned int radix_select(const float* __restrict__ depth, const unsigned char* __restrict__ mask, int W,
                                     int u0, int v0, int bw, int bh, unsigned int k, unsigned int* hist /*smem[256]*/,
                                     unsigned int* sh /*smem[2]*/) {
  // returns the bit pattern of the k-th smallest (0-based) valid masked depth; only the mask's bounding box
  // (u0, v0, bw x bh) is scanned
  unsigned int prefix = 0, prefix_mask = 0;
  const int nbox = bw * bh;
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int q = threadIdx.x; q < nbox; q += blockDim.x) {
      const int i = (v0 + q / bw) * W + u0 + q % bw;
      const float d = depth[i];
      if (mask[i] && d >= 0.001f) {
        const unsigned int b = __float_as_uint(d);
        if ((b & prefix_mask) == prefix) atomicAdd(&hist[(b >> shift) & 255u], 1u);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int acc = 0, bin = 0;
      for (; bin < 256; ++bin) {
        if (acc + hist[bin] > k) break;
        acc += hist[bin];
      }
      sh[0] = bin;
      sh[1] = k - acc;
    }
    __syncthreads();
    prefix |= sh[0] << shift;
    prefix_mask |= 255u << shift;
    k = sh[1];
    __syncthreads();
  }
  return prefix;
}

// pass 1 (grid-wide): mask bounding box and pixel counts.  All six statistics are max / sum reductions over values
// that start at 0 (the box minima are stored mirrored), so one 24-byte memset initialises them.
__global__ void __launch_bounds__(256) mask_stats_kernel(const float* __restrict__ depth,
                                                         const unsigned char* __restrict__ mask, int H, int W,
                                                         unsigned int* __restrict__ stats) {
  const int npix = H * W;
  unsigned int mu0 = 0, u1 = 0, mv0 = 0, v1 = 0, n_mask = 0, n_valid = 0;  // mu0 = W - 1 - umin, mv0 = H - 1 - vmin
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
    if (mask[i]) {
      const int v = i / W, u = i - v * W;
      mu0 = max(mu0, (unsigned)(W - 1 - u) + 1u);  // +1: 0 stays "no pixel"
      u1 = max(u1, (unsigned)u + 1u);
      mv0 = max(mv0, (unsigned)(H - 1 - v) + 1u);
      v1 = max(v1, (unsigned)v + 1u);
      ++n_mask;
      if (depth[i] >= 0.001f) ++n_valid;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mu0 = max(mu0, __shfl_xor_sync(0xffffffffu, mu0, o));
    u1 = max(u1, __shfl_xor_sync(0xffffffffu, u1, o));
    mv0 = max(mv0, __shfl_xor_sync(0xffffffffu, mv0, o));
    v1 = max(v1, __shfl_xor_sync(0xffffffffu, v1, o));
    n_mask += __shfl_xor_sync(0xffffffffu, n_mask, o);
    n_valid += __shfl_xor_sync(0xffffffffu, n_valid, o);
  }
  if ((threadIdx.x & 31) == 0 && n_mask) {
    atomicMax(&stats[0], mu0);
    atomicMax(&stats[1], u1);
    atomicMax(&stats[2], mv0);
    atomicMax(&stats[3], v1);
    atomicAdd(&stats[4], n_mask);
    atomicAdd(&stats[5], n_valid);
  }
}

// pass 2 (one CTA): exact median over the bounding box, translation, start poses
__global__ void __launch_bounds__(1024) start_poses_kernel(const float* __restrict__ depth,
                                                           const unsigned char* __restrict__ mask, int H, int W, float fx,
                                                           float fy, float cx, float cy, const float* __restrict__ rot_grid,
                                                           int N, const unsigned int* __restrict__ stats,
                                                           float* __restrict__ poses_out, float* __restrict__ info) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned int sh[2];
  __shared__ float tvec[3];
  const unsigned int nm = stats[4], nv = stats[5];
  // umin, umax, vmin, vmax
  const int bb[4] = {W - (int)stats[0], (int)stats[1] - 1, H - (int)stats[2], (int)stats[3] - 1};
  float zc = 0.f;
  if (nm > 0 && nv > 0) {  // uniform branch
    const int u0 = bb[0], v0 = bb[2], bw = bb[1] - bb[0] + 1, bh = bb[3] - bb[2] + 1;
    const unsigned int lo = radix_select(depth, mask, W, u0, v0, bw, bh, (nv - 1) / 2, hist, sh);
    const unsigned int hi = (nv & 1u) ? lo : radix_select(depth, mask, W, u0, v0, bw, bh, nv / 2, hist, sh);
    zc = (nv & 1u) ? __uint_as_float(lo) : (__uint_as_float(lo) + __uint_as_float(hi)) * 0.5f;
  }
  if (threadIdx.x == 0) {
    double t[3] = {0.0, 0.0, 0.0};
    if (nm > 0 && nv > 0) {
      const double uc = (bb[0] + bb[1]) / 2.0, vc = (bb[2] + bb[3]) / 2.0;
      // np.linalg.inv(K) @ [uc, vc, 1] * zc  for K = [[fx,0,cx],[0,fy,cy],[0,0,1]]
      t[0] = (uc - (double)cx) / (double)fx * (double)zc;
      t[1] = (vc - (double)cy) / (double)fy * (double)zc;
      t[2] = (double)zc;
    }
    tvec[0] = (float)t[0]; tvec[1] = (float)t[1]; tvec[2] = (float)t[2];
    info[0] = tvec[0]; info[1] = tvec[1]; info[2] = tvec[2]; info[3] = (float)nv;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < N * 16; i += blockDim.x) {
    const int e = i & 15;
    float v = rot_grid[i];
    if (e == 3) v = tvec[0];
    else if (e == 7) v = tvec[1];
    else if (e == 11) v = tvec[2];
    poses_out[i] = v;
  }
}

int start_poses_launch(const float* depth, const unsigned char* mask, int H, int W, float fx, float fy, float cx, float cy,
                       const float* rot_grid, int N, unsigned int* stats /*6 words of device scratch*/, float* poses_out,
                       float* info, cudaStream_t stream) {
  FP_CUDA_OK(cudaMemsetAsync(stats, 0, 6 * sizeof(unsigned int), stream));
  mask_stats_kernel<<<148, 256, 0, stream>>>(depth, mask, H, W, stats);
  start_poses_kernel<<<1, 1024, 0, stream>>>(depth, mask, H, W, fx, fy, cx, cy, rot_grid, N, stats, poses_out, info);
  note_launches(2);
  FP_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace fp
