// fp_crop.cu — fused "pose -> 160x160 network inputs" producer: for every pose hypothesis, one CTA
//   1. derives the crop window from the pose              (Utils.py:577-621 compute_crop_window_tf_batch, 'box_3d')
//   2. rasterises the textured mesh straight into the crop (Utils.py:133-219 nvdiffrast_render with bbox2d,
//      dr.rasterize / dr.interpolate / dr.texture; predict_pose_refine.py:44-53)
//   3. resamples the observed frame into the same window  (predict_pose_refine.py:63,72 / predict_score.py:89-90,
//      kornia warp_perspective -> F.grid_sample; h5_dataset.py:158-161 depth round trip for the scorer)
//   4. normalises both crops                              (h5_dataset.py:79-127 refiner, :137-179 scorer)
// and writes the two 6-channel crops as fp16 8-channel images with the 3-pixel zero border and the even/odd
// column split the 7x7 stem convolution reads (fp_stem.cu, LK_CONV7_S2).  Nothing full-frame and nothing fp32 is materialised in HBM.
//
// Raster: integer/fp32 load-store work (no tensor cores).  Vertices are snapped to 1/256 pixel and
// coverage is decided by exact 64-bit integer edge functions with a top-left tie rule (watertight);
// depth test = largest interpolated 1/Z, ties -> lowest triangle id, resolved by one 64-bit atomicMax
// per covered pixel on a z-buffer that lives entirely in shared memory (160*160*8 B = 200 KB).
#include "fp_crop.cuh"

#include "fp_common.cuh"
#include "fp_gemm.cuh"

namespace fp {

constexpr int S = 160;               // crop size (cfg.input_resize)

struct Window {
  float left, top, sx, sy;           // tf_to_crop = [[sx,0,-left*sx],[0,sy,-top*sy],[0,0,1]]
  float umin, vmin, rsx, rsy;        // render window origin and raster scale (pixels of crop per image pixel)
};

// Utils.py:602-621 + :584-598, fp32 with the reference's operation order (no FMA contraction so the
// rounded window edges are reproducible bit-for-bit by the oracle).
__device__ __forceinline__ void crop_window(const float* __restrict__ pose, float fx, float fy, float cx,
                                            float cy, float r3, Window& w) {
  const float tx = pose[3], ty = pose[7], tz = pose[11];
  float u[5], v[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const float ox = (k == 1) ? r3 : (k == 2 ? -r3 : 0.f);
    const float oy = (k == 3) ? r3 : (k == 4 ? -r3 : 0.f);
    const float px = __fadd_rn(tx, ox), py = __fadd_rn(ty, oy), pz = tz;
    const float x = __fadd_rn(__fmul_rn(fx, px), __fmul_rn(cx, pz));
    const float y = __fadd_rn(__fmul_rn(fy, py), __fmul_rn(cy, pz));
    u[k] = __fdiv_rn(x, pz);
    v[k] = __fdiv_rn(y, pz);
  }
  float radius = 0.f;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    radius = fmaxf(radius, fabsf(__fsub_rn(u[k], u[0])));
    radius = fmaxf(radius, fabsf(__fsub_rn(v[k], v[0])));
  }
  const float left = rintf(__fsub_rn(u[0], radius)), right = rintf(__fadd_rn(u[0], radius));
  const float top = rintf(__fsub_rn(v[0], radius)), bottom = rintf(__fadd_rn(v[0], radius));
  w.left = left;
  w.top = top;
  // Utils.py:594-595 `out_size[0] / (right - left)` is int / Tensor = Tensor.__rtruediv__ = reciprocal() * 160 in
  // torch: two roundings, not one division (tests/golden/geometry_golden.npz pins the bits)
  w.sx = __fmul_rn(__frcp_rn(__fsub_rn(right, left)), (float)S);
  w.sy = __fmul_rn(__frcp_rn(__fsub_rn(bottom, top)), (float)S);
  // predict_pose_refine.py:44-45: render window = crop corners (0,0)-(159,159) mapped back to the image
  w.umin = left;
  w.vmin = top;
  const float umax = __fadd_rn(left, __fdiv_rn(159.f, w.sx));
  const float vmax = __fadd_rn(top, __fdiv_rn(159.f, w.sy));
  w.rsx = __fdiv_rn((float)S, __fsub_rn(umax, w.umin));
  w.rsy = __fdiv_rn((float)S, __fsub_rn(vmax, w.vmin));
}

struct VtxScreen {
  int xi, yi;     // 1/256-pixel fixed point, crop raster space (y down)
  float iz;       // 1 / camera Z
  float X, Y, Z;  // camera-space position
};

__device__ __forceinline__ void xform_vertex(const float* __restrict__ P /*pose 4x4 row-major, smem*/,
                                             const float* __restrict__ vp, const Window& w, float fx, float fy,
                                             float cx, float cy, VtxScreen& o) {
  const float x = __ldg(vp), y = __ldg(vp + 1), z = __ldg(vp + 2);
  o.X = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(P[0], x), __fmul_rn(P[1], y)), __fmul_rn(P[2], z)), P[3]);
  o.Y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(P[4], x), __fmul_rn(P[5], y)), __fmul_rn(P[6], z)), P[7]);
  o.Z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(P[8], x), __fmul_rn(P[9], y)), __fmul_rn(P[10], z)), P[11]);
  o.iz = __frcp_rn(o.Z);
  const float u = __fadd_rn(__fmul_rn(__fmul_rn(fx, o.X), o.iz), cx);
  const float v = __fadd_rn(__fmul_rn(__fmul_rn(fy, o.Y), o.iz), cy);
  float px = __fmul_rn(__fsub_rn(u, w.umin), w.rsx);
  float py = __fmul_rn(__fsub_rn(v, w.vmin), w.rsy);
  px = fminf(fmaxf(px, -30000.f), 30000.f);
  py = fminf(fmaxf(py, -30000.f), 30000.f);
  o.xi = __float2int_rn(__fmul_rn(px, 256.f));
  o.yi = __float2int_rn(__fmul_rn(py, 256.f));
}

struct TriSetup {
  long long area2;
  int x0, y0, x1, y1, x2, y2;  // after orientation fix (area2 > 0)
  int swapped;                 // vertices 1 and 2 were exchanged
};

__device__ __forceinline__ bool tri_setup(const VtxScreen& a, const VtxScreen& b, const VtxScreen& c, TriSetup& t) {
  t.x0 = a.xi; t.y0 = a.yi; t.x1 = b.xi; t.y1 = b.yi; t.x2 = c.xi; t.y2 = c.yi;
  t.swapped = 0;
  long long area2 = (long long)(t.x1 - t.x0) * (t.y2 - t.y0) - (long long)(t.y1 - t.y0) * (t.x2 - t.x0);
  if (area2 == 0) return false;
  if (area2 < 0) {
    int tx = t.x1, ty = t.y1;
    t.x1 = t.x2; t.y1 = t.y2; t.x2 = tx; t.y2 = ty;
    t.swapped = 1;
    area2 = -area2;
  }
  t.area2 = area2;
  return true;
}

// edge function of edge (xa,ya)->(xb,yb) at pixel centre (px,py); all in 1/256 px
__device__ __forceinline__ long long edge_fn(int xa, int ya, int xb, int yb, int px, int py) {
  return (long long)(xb - xa) * (py - ya) - (long long)(yb - ya) * (px - xa);
}
// tie rule: a pixel centre exactly on an edge belongs to the triangle for which the (oriented) edge
// direction satisfies dy > 0 || (dy == 0 && dx > 0): exactly one of the two triangles sharing it.
__device__ __forceinline__ bool edge_ok(long long e, int dx, int dy) {
  return e > 0 || (e == 0 && (dy > 0 || (dy == 0 && dx > 0)));
}

// barycentric weights (screen space) of the *original* vertex order a, b, c; false if outside
__device__ __forceinline__ bool tri_cover(const TriSetup& t, int px, int py, float& b0, float& b1, float& b2) {
  const long long e0 = edge_fn(t.x1, t.y1, t.x2, t.y2, px, py);  // weight of vertex 0
  const long long e1 = edge_fn(t.x2, t.y2, t.x0, t.y0, px, py);  // weight of (oriented) vertex 1
  const long long e2 = t.area2 - e0 - e1;
  if (!edge_ok(e0, t.x2 - t.x1, t.y2 - t.y1) || !edge_ok(e1, t.x0 - t.x2, t.y0 - t.y2) ||
      !edge_ok(e2, t.x1 - t.x0, t.y1 - t.y0))
    return false;
  const float fa = __ll2float_rn(t.area2);
  b0 = __fdiv_rn(__ll2float_rn(e0), fa);
  const float w1 = __fdiv_rn(__ll2float_rn(e1), fa);
  const float w2 = __fdiv_rn(__ll2float_rn(e2), fa);
  b1 = t.swapped ? w2 : w1;
  b2 = t.swapped ? w1 : w2;
  return true;
}

__device__ __forceinline__ float inv_depth(float b0, float b1, float b2, float iz0, float iz1, float iz2) {
  return __fadd_rn(__fadd_rn(__fmul_rn(b0, iz0), __fmul_rn(b1, iz1)), __fmul_rn(b2, iz2));
}

// kornia.warp_perspective(..., align_corners=False) coordinate chain: destination
// pixel index d, affine map x = d * inv_scale + offset into a source of `size` pixels, then the
// (size-1)-normalisation followed by grid_sample's align_corners=False un-normalisation.
__device__ __forceinline__ float kornia_src_coord(float x_src, int size) {
  const float xn = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, x_src), (float)(size - 1)), 1.f);
  return __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(xn, 1.f), (float)size), 1.f), 0.5f);
}

__device__ __forceinline__ void normalise_xyz(float x, float y, float z, const float* t, float inv_radius, float tau,
                                              float& ox, float& oy, float& oz) {
  // h5_dataset.py:93-99 / :151-156
  const bool inv = z < tau;
  ox = (x - t[0]) * inv_radius;
  oy = (y - t[1]) * inv_radius;
  oz = (z - t[2]) * inv_radius;
  if (inv || fabsf(ox) >= 2.f) ox = 0.f;
  if (inv || fabsf(oy) >= 2.f) oy = 0.f;
  if (inv || fabsf(oz) >= 2.f) oz = 0.f;
}

// ------------------------------------------------------------------------------------------------
// pass 0: one thread per (hypothesis, vertex): camera transform, projection into the crop raster,
// 1/256-px snap, per-vertex diffuse term.  16 + 16 bytes per vertex, written once, read ~6x (L2).
// ------------------------------------------------------------------------------------------------
constexpr int kTabWords = 6 * S;  // per hypothesis: colf, rowf (float), coln, rown, colz, rowz (int)

struct __align__(16) VtxA {  // what the z-buffer pass needs
  int xi, yi;
  float iz, Z;
};
struct __align__(16) VtxB {  // what shading additionally needs
  float X, Y, dif, pad;
};

__global__ void __launch_bounds__(256) vertex_kernel(const CropParams p) {
  __shared__ float sP[16];
  __shared__ Window sW;
  pdl_trigger();
  pdl_wait();  // poses come from the previous iteration's pose update; vtx / tab buffers are still read by its shade
  const int n = blockIdx.y;
  if (threadIdx.x < 16) sP[threadIdx.x] = p.poses[(size_t)n * 16 + threadIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) {
    Window w;
    crop_window(sP, p.fx, p.fy, p.cx, p.cy, p.r3, w);
    sW = w;
    if (blockIdx.x == 0) {
      float* wb = p.win_buf + (size_t)n * 8;
      wb[0] = w.left; wb[1] = w.top; wb[2] = w.sx; wb[3] = w.sy;
      wb[4] = w.umin; wb[5] = w.vmin; wb[6] = w.rsx; wb[7] = w.rsy;
      if (p.win_out) {
        p.win_out[n * 4 + 0] = w.left;
        p.win_out[n * 4 + 1] = w.top;
        p.win_out[n * 4 + 2] = w.sx;
        p.win_out[n * 4 + 3] = w.sy;
      }
    }
  }
  __syncthreads();
  if (blockIdx.x == 0) {
    // per-axis tables of the observed-crop resampling (every quantity is separable in x and y): S columns, S rows
    for (int tI = threadIdx.x; tI < 2 * S; tI += blockDim.x) {
      const bool is_row = tI >= S;
      const int d = is_row ? tI - S : tI;
      const float sc = is_row ? sW.sy : sW.sx, org = is_row ? sW.top : sW.left;
      const int size = is_row ? p.H : p.W;
      const float xs = __fadd_rn(__fdiv_rn((float)d, sc), org);
      const float ix = kornia_src_coord(xs, size);
      int un = (int)rintf(ix);
      if (un < 0 || un >= size) un = -1;
      int uz = -1;
      if (un >= 0) {
        // scorer: depth crop -> full-res (nearest) -> back-project -> crop (nearest), h5_dataset.py:158-161
        const float xc = __fadd_rn(__fmul_rn(sc, (float)un), __fmul_rn(-org, sc));
        const int jc = (int)rintf(kornia_src_coord(xc, S));
        if (jc >= 0 && jc < S) {
          const float xs2 = __fadd_rn(__fdiv_rn((float)jc, sc), org);
          const int u2 = (int)rintf(kornia_src_coord(xs2, size));
          if (u2 >= 0 && u2 < size) uz = u2;
        }
      }
      float* tb = p.tab_buf + (size_t)n * kTabWords;
      tb[(is_row ? 1 : 0) * S + d] = ix;
      reinterpret_cast<int*>(tb)[(2 + (is_row ? 1 : 0)) * S + d] = un;
      reinterpret_cast<int*>(tb)[(4 + (is_row ? 1 : 0)) * S + d] = uz;
    }
  }
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= p.V) return;
  VtxScreen o;
  xform_vertex(sP, p.vpos + 3 * v, sW, p.fx, p.fy, p.cx, p.cy, o);
  // diffuse = clip(normalize(R n) . (0,0,-1), 0, 1)   (Utils.py:203-207)
  const float nx = __ldg(p.vnrm + 3 * v), ny = __ldg(p.vnrm + 3 * v + 1), nz = __ldg(p.vnrm + 3 * v + 2);
  const float cxn = sP[0] * nx + sP[1] * ny + sP[2] * nz;
  const float cyn = sP[4] * nx + sP[5] * ny + sP[6] * nz;
  const float czn = sP[8] * nx + sP[9] * ny + sP[10] * nz;
  const float len = fmaxf(sqrtf(cxn * cxn + cyn * cyn + czn * czn), 1e-12f);
  VtxA a;
  a.xi = o.xi; a.yi = o.yi; a.iz = o.iz; a.Z = o.Z;
  VtxB b;
  b.X = o.X; b.Y = o.Y; b.dif = fminf(fmaxf(-czn / len, 0.f), 1.f); b.pad = 0.f;
  p.vtx_a[(size_t)n * p.V + v] = a;
  p.vtx_b[(size_t)n * p.V + v] = b;
}

// 32-bit edge functions relative to vertex 0 when the triangle is small enough (all deltas < 2^15, i.e.
// < 128 px: products < 2^30); the integers are the same as the 64-bit ones, so coverage is unchanged.
struct TriSetup32 {
  int area2;
  int x0, y0;               // vertex 0 (absolute, 1/256 px)
  int ax, ay, bx, by;       // oriented vertices 1 and 2 relative to vertex 0
  int swapped;
};

__device__ __forceinline__ bool tri_small(int x0, int y0, int x1, int y1, int x2, int y2) {
  const int m = max(max(abs(x1 - x0), abs(y1 - y0)), max(abs(x2 - x0), abs(y2 - y0)));
  return m < 16384;
}
__device__ __forceinline__ bool tri_setup32(int x0, int y0, int x1, int y1, int x2, int y2, TriSetup32& t) {
  t.x0 = x0; t.y0 = y0;
  t.ax = x1 - x0; t.ay = y1 - y0; t.bx = x2 - x0; t.by = y2 - y0;
  t.swapped = 0;
  int area2 = t.ax * t.by - t.ay * t.bx;
  if (area2 == 0) return false;
  if (area2 < 0) {
    int tx = t.ax, ty = t.ay;
    t.ax = t.bx; t.ay = t.by; t.bx = tx; t.by = ty;
    t.swapped = 1;
    area2 = -area2;
  }
  t.area2 = area2;
  return true;
}
__device__ __forceinline__ bool edge_ok32(int e, int dx, int dy) {
  return e > 0 || (e == 0 && (dy > 0 || (dy == 0 && dx > 0)));
}
// pixel centre (px, py) absolute; must lie inside the triangle's bounding box (deltas < 2^15)
__device__ __forceinline__ bool tri_cover32(const TriSetup32& t, int px, int py, float& b0, float& b1, float& b2) {
  const int qx = px - t.x0, qy = py - t.y0;
  // e0: edge v1->v2 (weight of v0); e1: edge v2->v0 (weight of v1); e2 = area2 - e0 - e1
  const int e0 = (t.bx - t.ax) * (qy - t.ay) - (t.by - t.ay) * (qx - t.ax);
  const int e1 = (-t.bx) * (qy - t.by) - (-t.by) * (qx - t.bx);
  const int e2 = t.area2 - e0 - e1;
  if (!edge_ok32(e0, t.bx - t.ax, t.by - t.ay) || !edge_ok32(e1, -t.bx, -t.by) || !edge_ok32(e2, t.ax, t.ay)) return false;
  const float fa = __int2float_rn(t.area2);
  b0 = __fdiv_rn(__int2float_rn(e0), fa);
  const float w1 = __fdiv_rn(__int2float_rn(e1), fa);
  const float w2 = __fdiv_rn(__int2float_rn(e2), fa);
  b1 = t.swapped ? w2 : w1;
  b2 = t.swapped ? w1 : w2;
  return true;
}

// rare path: triangles spanning >= 128 px use the 64-bit edge functions (same integers, same coverage)
__device__ __noinline__ void raster_big_tri(const VtxA a, const VtxA b, const VtxA c, int f, int j0, int j1, int r0,
                                            int r1, unsigned long long* zbuf) {
  VtxScreen sa, sb, sc;
  sa.xi = a.xi; sa.yi = a.yi; sb.xi = b.xi; sb.yi = b.yi; sc.xi = c.xi; sc.yi = c.yi;
  TriSetup t;
  if (!tri_setup(sa, sb, sc, t)) return;
  for (int r = r0; r <= r1; ++r)
    for (int j = j0; j <= j1; ++j) {
      float b0, b1, b2;
      if (!tri_cover(t, j * 256 + 128, r * 256 + 128, b0, b1, b2)) continue;
      const float iz = inv_depth(b0, b1, b2, a.iz, b.iz, c.iz);
      const unsigned long long key =
          ((unsigned long long)__float_as_uint(iz) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)f);
      atomicMax(&zbuf[r * S + j], key);
    }
}
__device__ __noinline__ void bary_big_tri(const VtxA a, const VtxA b, const VtxA c, int px, int py, float* bw) {
  VtxScreen sa, sb, sc;
  sa.xi = a.xi; sa.yi = a.yi; sb.xi = b.xi; sb.yi = b.yi; sc.xi = c.xi; sc.yi = c.yi;
  TriSetup t;
  tri_setup(sa, sb, sc, t);
  float b0 = 0.f, b1 = 0.f, b2 = 0.f;
  tri_cover(t, px, py, b0, b1, b2);
  bw[0] = b0; bw[1] = b1; bw[2] = b2;
}

// ------------------------------------------------------------------------------------------------
// pass 1: one thread per (hypothesis, triangle): coverage + depth test into the per-hypothesis z-buffer
// (64-bit keys in global memory = L2: 200 KB per hypothesis), so the work spreads over the whole GPU for
// any batch size (one pose in track_one, ~32 per GPU when sharded, 252 in register).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) raster_kernel(const CropParams p) {
  pdl_trigger();
  pdl_wait();
  const int n = blockIdx.y;
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= p.F) return;
  const VtxA* va = p.vtx_a + (size_t)n * p.V;
  unsigned long long* zbuf = p.zbuf + (size_t)n * S * S;
  const int i0 = __ldg(p.faces + 3 * f), i1 = __ldg(p.faces + 3 * f + 1), i2 = __ldg(p.faces + 3 * f + 2);
  const VtxA a = va[i0], b = va[i1], c = va[i2];
  if (!(a.Z > p.znear && b.Z > p.znear && c.Z > p.znear)) return;  // no near-plane clipping
  const int minx = min(a.xi, min(b.xi, c.xi)), maxx = max(a.xi, max(b.xi, c.xi));
  const int miny = min(a.yi, min(b.yi, c.yi)), maxy = max(a.yi, max(b.yi, c.yi));
  const int j0 = max((minx + 127) >> 8, 0), j1 = min((maxx - 128) >> 8, S - 1);
  const int r0 = max((miny + 127) >> 8, 0), r1 = min((maxy - 128) >> 8, S - 1);
  if (j0 > j1 || r0 > r1) return;
  if (!tri_small(a.xi, a.yi, b.xi, b.yi, c.xi, c.yi)) {
    raster_big_tri(a, b, c, f, j0, j1, r0, r1, zbuf);
    return;
  }
  TriSetup32 t32;
  if (!tri_setup32(a.xi, a.yi, b.xi, b.yi, c.xi, c.yi, t32)) return;
  for (int r = r0; r <= r1; ++r)
    for (int j = j0; j <= j1; ++j) {
      float b0, b1, b2;
      if (!tri_cover32(t32, j * 256 + 128, r * 256 + 128, b0, b1, b2)) continue;
      const float iz = inv_depth(b0, b1, b2, a.iz, b.iz, c.iz);
      const unsigned long long key =
          ((unsigned long long)__float_as_uint(iz) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)f);
      atomicMax(&zbuf[r * S + j], key);
    }
}

// ------------------------------------------------------------------------------------------------
// pass 2: one thread per (hypothesis, crop pixel): shade the winning triangle (A), resample the observed
// frame (B), normalise both, write the two fp16 NHWC(8) pixels.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) shade_kernel(const CropParams p) {
  __shared__ float sP[16];
  pdl_trigger();
  pdl_wait();
  const int n = blockIdx.y;
  if (threadIdx.x < 16) sP[threadIdx.x] = p.poses[(size_t)n * 16 + threadIdx.x];
  __syncthreads();
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= S * S) return;
  const float* tb = p.tab_buf + (size_t)n * kTabWords;
  const int* tbi = reinterpret_cast<const int*>(tb);
  const VtxA* va = p.vtx_a + (size_t)n * p.V;
  const VtxB* vb = p.vtx_b + (size_t)n * p.V;
  const float inv_radius = p.inv_radius;
  const float tvec[3] = {sP[3], sP[7], sP[11]};
  const float tau = p.mode == 0 ? 0.001f : 0.1f;
  const size_t img_stride = (size_t)(S + 6) * (S + 8) * 8;
  __half* outA = p.crops + (size_t)n * img_stride;
  __half* outB = p.crops + (size_t)(p.b_img0 + n) * img_stride;
  {
    const int r = pix / S, j = pix - r * S;
    // ---- A: rendered crop
    float ar = 0.f, ag = 0.f, ab = 0.f, ax = 0.f, ay = 0.f, az = 0.f;
    const unsigned long long key = p.zbuf[(size_t)n * S * S + pix];
    if (key != 0ull) {
      p.zbuf[(size_t)n * S * S + pix] = 0ull;  // leave the z-buffer clean for the next launch (no separate clear pass)
      const int f = (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull));
      const int i0 = __ldg(p.faces + 3 * f), i1 = __ldg(p.faces + 3 * f + 1), i2 = __ldg(p.faces + 3 * f + 2);
      const VtxA a = va[i0], b = va[i1], c = va[i2];
      const VtxB a2 = vb[i0], b2v = vb[i1], c2 = vb[i2];
      float b0, b1, b2;
      if (tri_small(a.xi, a.yi, b.xi, b.yi, c.xi, c.yi)) {
        TriSetup32 t32;
        tri_setup32(a.xi, a.yi, b.xi, b.yi, c.xi, c.yi, t32);
        tri_cover32(t32, j * 256 + 128, r * 256 + 128, b0, b1, b2);
      } else {
        float bw[3];
        bary_big_tri(a, b, c, j * 256 + 128, r * 256 + 128, bw);
        b0 = bw[0]; b1 = bw[1]; b2 = bw[2];
      }
      const float iz = inv_depth(b0, b1, b2, a.iz, b.iz, c.iz);
      // perspective-correct weights (nvdiffrast: barycentrics computed in clip space)
      const float z = 1.f / iz;
      const float w0 = b0 * a.iz * z, w1 = b1 * b.iz * z, w2 = b2 * c.iz * z;
      const float X = w0 * a2.X + w1 * b2v.X + w2 * c2.X;
      const float Y = w0 * a2.Y + w1 * b2v.Y + w2 * c2.Y;
      const float Z = w0 * a.Z + w1 * b.Z + w2 * c.Z;
      const float diffuse = w0 * a2.dif + w1 * b2v.dif + w2 * c2.dif;
      float cr, cg, cb;
      if (p.tex) {
        const float tu = w0 * __ldg(p.vuv + 2 * i0) + w1 * __ldg(p.vuv + 2 * i1) + w2 * __ldg(p.vuv + 2 * i2);
        const float tv = w0 * __ldg(p.vuv + 2 * i0 + 1) + w1 * __ldg(p.vuv + 2 * i1 + 1) + w2 * __ldg(p.vuv + 2 * i2 + 1);
        // dr.texture(filter_mode='linear', boundary 'wrap'): texel centres at +0.5
        const float xx = tu * p.Wt - 0.5f, yy = tv * p.Ht - 0.5f;
        const float xf = floorf(xx), yf = floorf(yy);
        const float ax1 = xx - xf, ay1 = yy - yf;
        int x0 = (int)xf % p.Wt, y0 = (int)yf % p.Ht;
        if (x0 < 0) x0 += p.Wt;
        if (y0 < 0) y0 += p.Ht;
        const int x1 = (x0 + 1 == p.Wt) ? 0 : x0 + 1, y1 = (y0 + 1 == p.Ht) ? 0 : y0 + 1;
        const uchar4 t00 = __ldg(p.tex + (size_t)y0 * p.Wt + x0), t01 = __ldg(p.tex + (size_t)y0 * p.Wt + x1);
        const uchar4 t10 = __ldg(p.tex + (size_t)y1 * p.Wt + x0), t11 = __ldg(p.tex + (size_t)y1 * p.Wt + x1);
        const float w00 = (1.f - ax1) * (1.f - ay1), w01 = ax1 * (1.f - ay1), w10 = (1.f - ax1) * ay1, w11 = ax1 * ay1;
        const float k255 = 1.f / 255.f;
        cr = (w00 * t00.x + w01 * t01.x + w10 * t10.x + w11 * t11.x) * k255;
        cg = (w00 * t00.y + w01 * t01.y + w10 * t10.y + w11 * t11.y) * k255;
        cb = (w00 * t00.z + w01 * t01.z + w10 * t10.z + w11 * t11.z) * k255;
      } else {
        cr = w0 * __ldg(p.vcol + 3 * i0) + w1 * __ldg(p.vcol + 3 * i1) + w2 * __ldg(p.vcol + 3 * i2);
        cg = w0 * __ldg(p.vcol + 3 * i0 + 1) + w1 * __ldg(p.vcol + 3 * i1 + 1) + w2 * __ldg(p.vcol + 3 * i2 + 1);
        cb = w0 * __ldg(p.vcol + 3 * i0 + 2) + w1 * __ldg(p.vcol + 3 * i1 + 2) + w2 * __ldg(p.vcol + 3 * i2 + 2);
      }
      // color*w_ambient + diffuse*color*w_diffuse, clip(0,1)   (Utils.py:211-213)
      ar = fminf(fmaxf(cr * 0.8f + diffuse * cr * 0.5f, 0.f), 1.f);
      ag = fminf(fmaxf(cg * 0.8f + diffuse * cg * 0.5f, 0.f), 1.f);
      ab = fminf(fmaxf(cb * 0.8f + diffuse * cb * 0.5f, 0.f), 1.f);
      normalise_xyz(X, Y, Z, tvec, inv_radius, tau, ax, ay, az);
    }
    // ---- B: observed crop
    float br = 0.f, bg = 0.f, bb = 0.f, bx = 0.f, by = 0.f, bz = 0.f;
    {
      const float ix = __ldg(tb + j), iy = __ldg(tb + S + r);
      // bilinear rgb, zeros padding
      const float fx0 = floorf(ix), fy0 = floorf(iy);
      const int x0 = (int)fx0, y0 = (int)fy0;
      const float wx1 = ix - fx0, wy1 = iy - fy0;
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const int xq = x0 + dx, yq = y0 + dy;
          if (xq >= 0 && xq < p.W && yq >= 0 && yq < p.H) {
            const float wgt = (dx ? wx1 : 1.f - wx1) * (dy ? wy1 : 1.f - wy1);
            const uchar4 c4 = __ldg(p.rgb + (size_t)yq * p.W + xq);
            br += wgt * c4.x;
            bg += wgt * c4.y;
            bb += wgt * c4.z;
          }
        }
      br *= (1.f / 255.f);
      bg *= (1.f / 255.f);
      bb *= (1.f / 255.f);
      // nearest geometry
      const int un = __ldg(tbi + 2 * S + j), vn = __ldg(tbi + 3 * S + r);
      float X = 0.f, Y = 0.f, Z = 0.f;
      if (un >= 0 && vn >= 0) {
        if (p.mode == 0) {
          // refiner: xyz_map (depth2xyzmap, Utils.py:399-438) sampled nearest
          const float* q = p.xyz_map + ((size_t)vn * p.W + un) * 3;
          X = __ldg(q);
          Y = __ldg(q + 1);
          Z = __ldg(q + 2);
        } else {
          const int u2 = __ldg(tbi + 4 * S + j), v2 = __ldg(tbi + 5 * S + r);
          float zz = 0.f;
          if (u2 >= 0 && v2 >= 0) zz = __ldg(p.depth + (size_t)v2 * p.W + u2);
          if (zz >= 0.001f) {  // depth2xyzmap_batch(zfar=inf): invalid z<0.001 -> 0
            X = ((float)un - p.cx) * zz / p.fx;
            Y = ((float)vn - p.cy) * zz / p.fy;
            Z = zz;
          }
        }
      }
      normalise_xyz(X, Y, Z, tvec, inv_radius, tau, bx, by, bz);
    }
    // even / odd padded columns live in separate half-rows ("EO" layout, fp_stem.cu): pixel (row, col) ->
    // [row][col & 1][col >> 1][8]
    const int pc = j + 3;
    const size_t off = (((size_t)(r + 3) * 2 + (pc & 1)) * ((S + 8) / 2) + (pc >> 1)) * 8;
    *reinterpret_cast<uint4*>(outA + off) = make_uint4(pack_half2(ar, ag), pack_half2(ab, ax), pack_half2(ay, az), 0u);
    *reinterpret_cast<uint4*>(outB + off) = make_uint4(pack_half2(br, bg), pack_half2(bb, bx), pack_half2(by, bz), 0u);
    if (p.dbg) {
      float* d = p.dbg + (((size_t)n * 2 + 0) * S * S + pix) * 6;
      d[0] = ar; d[1] = ag; d[2] = ab; d[3] = ax; d[4] = ay; d[5] = az;
      d = p.dbg + (((size_t)n * 2 + 1) * S * S + pix) * 6;
      d[0] = br; d[1] = bg; d[2] = bb; d[3] = bx; d[4] = by; d[5] = bz;
    }
  }
}

int crop_launch(const CropParams& p, cudaStream_t stream) {
  if (p.N == 0) return 0;
  // algorithmic bytes: the two 6-channel fp16 crops each hypothesis produces
  prof_mark_begin(1, (double)p.N * 2.0 * 6.0 * S * S * 2.0, stream);
  // the z-buffer is all-zero here: allocated zeroed, and shade_kernel clears every cell it consumes
  FP_CUDA_OK(launch_pdl(vertex_kernel, dim3((p.V + 255) / 256, p.N), dim3(256), 0, stream, 1, p));
  FP_CUDA_OK(launch_pdl(raster_kernel, dim3((p.F + 255) / 256, p.N), dim3(256), 0, stream, 1, p));
  FP_CUDA_OK(launch_pdl(shade_kernel, dim3((S * S + 255) / 256, p.N), dim3(256), 0, stream, 1, p));
  prof_mark_end(stream);
  g_launch_count += 3;
  FP_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// frame preparation: uint8 RGB -> uchar4, depth -> xyz map (Utils.py:399-438)
// ------------------------------------------------------------------------------------------------
__global__ void rgb_to_rgba_kernel(const unsigned char* __restrict__ rgb, uchar4* __restrict__ out, int npix) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < npix) out[i] = make_uchar4(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2], 255);
}

__global__ void depth_to_xyz_kernel(const float* __restrict__ depth, float* __restrict__ xyz, int H, int W, float fx,
                                    float fy, float cx, float cy, float zfar) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * W) return;
  const int v = i / W, u = i - v * W;
  const float z = depth[i];
  float X = 0.f, Y = 0.f, Z = 0.f;
  if (!(z < 0.001f) && !(z > zfar)) {
    X = ((float)u - cx) * z / fx;
    Y = ((float)v - cy) * z / fy;
    Z = z;
  }
  xyz[3 * i] = X;
  xyz[3 * i + 1] = Y;
  xyz[3 * i + 2] = Z;
}

int rgb_to_rgba_launch(const unsigned char* rgb, uchar4* out, int npix, cudaStream_t stream) {
  rgb_to_rgba_kernel<<<(npix + 255) / 256, 256, 0, stream>>>(rgb, out, npix);
  ++g_launch_count;
  FP_CUDA_OK(cudaGetLastError());
  return 0;
}

int depth_to_xyz_launch(const float* depth, float* xyz, int H, int W, float fx, float fy, float cx, float cy,
                        float zfar, cudaStream_t stream) {
  depth_to_xyz_kernel<<<(H * W + 255) / 256, 256, 0, stream>>>(depth, xyz, H, W, fx, fy, cx, cy, zfar);
  ++g_launch_count;
  FP_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace fp
