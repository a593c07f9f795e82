// fp_depth.cuh — launchers of the depth pre-processing kernels (fp_depth.cu).
#pragma once
#include <cuda_runtime.h>

namespace fp {
int erode_depth_launch(const float* depth, float* out, int H, int W, int radius, float diff_thres, float ratio_thres,
                       float zfar, cudaStream_t stream);
int bilateral_depth_launch(const float* depth, float* out, int H, int W, int radius, float zfar, float sigmaD,
                           float sigmaR, cudaStream_t stream);
// erode(2) -> bilateral(2) -> back-projection (invalid: z < 0.001 or z > zfar_xyz) + rgb -> rgba, one launch
int frame_prep_launch(const unsigned char* rgb, const float* depth, uchar4* rgba, float* depth_out, float4* xyz, int H, int W,
                      float fx, float fy, float cx, float cy, float zfar_xyz, cudaStream_t stream);
int start_poses_launch(const float* depth, const unsigned char* mask, int H, int W, float fx, float fy, float cx, float cy,
                       const float* rot_grid, int N, unsigned int* stats, float* poses_out, float* info,
                       cudaStream_t stream);
}  // namespace fp
