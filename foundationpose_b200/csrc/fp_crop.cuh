// fp_crop.cuh — parameters of the fused crop producer (fp_crop.cu).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace fp {

struct CropParams {
  const float* poses;  // [N][16] row-major ob_in_cam
  int N;
  float fx, fy, cx, cy;
  int H, W;
  float r3;          // mesh_diameter * crop_ratio / 2 (Utils.py:603)
  float inv_radius;  // 1 / (mesh_diameter / 2)       (h5_dataset.py:96)
  float znear;
  // mesh (device)
  const float* vpos;   // [V][3]
  const float* vnrm;   // [V][3]
  const float* vuv;    // [V][2] (v already flipped, Utils.py:117) or null
  const float* vcol;   // [V][3] in 0..1 or null
  const int* faces;    // [F][3]
  int V, F;
  const uchar4* tex;   // [Ht][Wt] RGBA8 or null
  int Ht, Wt;
  // frame (device)
  const uchar4* rgb;     // [H][W] RGBA8
  const float* xyz_map;  // [H][W][3]   (mode 0)
  const float* depth;    // [H][W]      (mode 1)
  int mode;              // 0 = refiner crops, 1 = scorer crops
  // outputs
  __half* crops;   // [b_img0 + N][166][2][84][8] fp16 (rows x {even, odd columns} x column pairs x 8 channels, the
                   // "EO" layout of fp_stem.cu): images 0..N-1 = rendered (A), b_img0..b_img0+N-1 = observed (B)
  int b_img0;      // first B image (N rounded up to the conv tile's image count, see fp_api.cu)
  // workspaces (device): per-hypothesis pre-transformed vertices and crop windows
  struct VtxA* vtx_a;  // [N][V] 16 B
  struct VtxB* vtx_b;  // [N][V] 16 B
  float* win_buf;      // [N][8]
  float* tab_buf;      // [N][6][160] per-axis resampling tables
  unsigned long long* zbuf;  // [N][160*160] depth/triangle keys
  float* dbg;      // optional [N][2][160][160][6] fp32 copy of the normalised crops
  float* win_out;  // optional [N][4] = (left, top, sx, sy)
};

int crop_launch(const CropParams& p, cudaStream_t stream);
int rgb_to_rgba_launch(const unsigned char* rgb, uchar4* out, int npix, cudaStream_t stream);
int depth_to_xyz_launch(const float* depth, float* xyz, int H, int W, float fx, float fy, float cx, float cy,
                        float zfar, cudaStream_t stream);

}  // namespace fp
