// umma_probe.cu — micro-experiment: does a 128B-swizzled K-major tcgen05 operand descriptor accept a start
// address that is NOT aligned to the 1024-byte swizzle period, and 8-row groups whose stride (SBO) is not a
// multiple of 1024 bytes?  (Needed for reading shifted 3x3-convolution taps out of ONE shared-memory patch.)
//
//   X [256 rows][64 fp16] is TMA-loaded (SWIZZLE_128B) to a 1024-aligned buffer; W [128][64] likewise.
//   D[128 x 128] = A x W^T, A row m = X[off + (m / 8) * group_rows + (m % 8)]
//   descriptor: start = X + off * 128 B, SBO = group_rows * 128 B, base_offset field = mode ? (start >> 7) & 7 : 0
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I foundationpose_b200/csrc \
//              tools/umma_probe.cu -o tools/umma_probe -lcuda
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "fp_common.cuh"

namespace fp {
void set_last_error(const char*, ...) {}
bool pdl_enabled() { return false; }
}  // namespace fp
using namespace fp;

__global__ void __launch_bounds__(128, 1)
    probe_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, float* out, int off,
                 int group_rows, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* xs = smem;               // 256 rows x 128 B = 32 KB
  uint8_t* ws = smem + 32768;       // 128 rows x 128 B = 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 49152);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bars[0], 32768 + 16384);
    tma_load_2d(&map_x, &bars[0], xs, 0, 0);
    tma_load_2d(&map_w, &bars[0], ws, 0, 0);
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint32_t a_addr = smem_u32(xs) + off * 128;
    uint64_t da = 0;
    da |= (uint64_t)((a_addr & 0x3FFFF) >> 4);
    da |= (uint64_t)((group_rows * 128) >> 4) << 32;
    da |= (uint64_t)1 << 46;
    if (mode) da |= (uint64_t)((a_addr >> 7) & 7) << 49;
    da |= (uint64_t)2 << 61;
    const uint64_t db = umma_desc_sw128(smem_u32(ws));
    constexpr uint32_t idesc = umma_idesc_f16(128, 128);
    for (int k = 0; k < 4; ++k) umma_f16(tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, k > 0);
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c = 0; c < 128; c += 32) {
    uint32_t v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[row * 128 + c + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 128);
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int XR = 256;
  std::vector<__half> hx(XR * 64), hw(128 * 64);
  std::vector<float> fx(XR * 64), fw(128 * 64);
  srand(1);
  for (int i = 0; i < XR * 64; ++i) {
    fx[i] = (float)((rand() % 7) - 3);
    hx[i] = __float2half(fx[i]);
  }
  for (int i = 0; i < 128 * 64; ++i) {
    fw[i] = (float)((rand() % 5) - 2);
    hw[i] = __float2half(fw[i]);
  }
  __half *dx, *dw;
  float* dout;
  cudaMalloc(&dx, hx.size() * 2);
  cudaMalloc(&dw, hw.size() * 2);
  cudaMalloc(&dout, 128 * 128 * 4);
  cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice);
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fnp;
  CUtensorMap mx, mw;
  {
    cuuint64_t d[2] = {64, (cuuint64_t)XR}, s[1] = {128};
    cuuint32_t b[2] = {64, 256}, e[2] = {1, 1};
    CUresult r = enc(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dx, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cuuint64_t d2[2] = {64, 128};
    cuuint32_t b2[2] = {64, 128};
    CUresult r2 = enc(&mw, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dw, d2, s, b2, e, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS || r2 != CUDA_SUCCESS) {
      printf("encode failed %d %d\n", (int)r, (int)r2);
      return 1;
    }
  }
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 51200 + 1024);
  std::vector<float> hout(128 * 128);
  const int cfgs[][2] = {{0, 8}, {1, 8}, {2, 8}, {3, 8}, {5, 8}, {8, 8}, {0, 10}, {1, 10}, {2, 10}, {11, 10}, {12, 10}, {22, 10}, {0, 12}, {3, 14}};
  for (auto& c : cfgs) {
    for (int mode = 0; mode < 2; ++mode) {
      const int off = c[0], gr = c[1];
      cudaMemset(dout, 0, 128 * 128 * 4);
      probe_kernel<<<1, 128, 51200 + 1024>>>(mx, mw, dout, off, gr, mode);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("off %2d group_rows %2d base_offset_mode %d : CUDA error %s\n", off, gr, mode, cudaGetErrorString(e));
        return 2;
      }
      cudaMemcpy(hout.data(), dout, 128 * 128 * 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int m = 0; m < 128; ++m) {
        const int xr = off + (m / 8) * gr + (m % 8);
        for (int n = 0; n < 128; ++n) {
          float ref = 0.f;
          for (int k = 0; k < 64; ++k) ref += fx[xr * 64 + k] * fw[n * 64 + k];
          if (ref != hout[m * 128 + n]) ++bad;
        }
      }
      printf("off %2d group_rows %2d base_offset_mode %d : %s (%d / 16384 wrong)\n", off, gr, mode, bad ? "MISMATCH" : "exact", bad);
    }
  }
  return 0;
}
