// eigen_shim.h — just enough of Eigen's interface to compile two functions of the reference's C++ helper library AS THEY
// ARE (TEST INFRASTRUCTURE, see oracle/build_ref.py):
//     cluster_poses              /root/reference/mycpp/src/app/pybind_api.cpp:24-68
//     Utils::rotationGeodesicDistance   /root/reference/mycpp/src/Utils.cpp:21-26
// Eigen is not in this image, so the reference's library cannot be built by its own recipe; those two function bodies use
// only: Matrix4f / Matrix3f / Vector3f values, .block(i, j, r, c), operator* and operator-, .transpose(), .trace(),
// .norm(), and std::vector with Eigen::aligned_allocator.  This header provides exactly that, in single precision with
// plain sequential sums (Eigen may vectorise the 4x4 product; the results can differ in the last bit, which only matters
// for a pose lying exactly on the 30-degree threshold).
#pragma once
#include <cmath>
#include <cstddef>
#include <memory>
#include <vector>

namespace Eigen {

template <class T>
using aligned_allocator = std::allocator<T>;

template <int R, int C>
struct Mat {
  float v[R][C];
  Mat() {
    for (int i = 0; i < R; ++i)
      for (int j = 0; j < C; ++j) v[i][j] = 0.f;
  }
  float& operator()(int i, int j) { return v[i][j]; }
  float operator()(int i, int j) const { return v[i][j]; }

  // a run-time block: converts to whichever fixed-size matrix the caller asks for (sizes are checked by construction
  // in the two call sites: (0,3,3,1) -> Vector3f, (0,0,3,3) -> Matrix3f)
  struct Block {
    const Mat* m;
    int i0, j0, r, c;
    template <int R2, int C2>
    operator Mat<R2, C2>() const {
      Mat<R2, C2> o;
      for (int i = 0; i < R2; ++i)
        for (int j = 0; j < C2; ++j) o.v[i][j] = m->v[i0 + i][j0 + j];
      return o;
    }
  };
  Block block(int i0, int j0, int r, int c) const { return Block{this, i0, j0, r, c}; }

  Mat<C, R> transpose() const {
    Mat<C, R> o;
    for (int i = 0; i < R; ++i)
      for (int j = 0; j < C; ++j) o.v[j][i] = v[i][j];
    return o;
  }
  float trace() const {
    float s = 0.f;
    for (int i = 0; i < (R < C ? R : C); ++i) s += v[i][i];
    return s;
  }
  float norm() const {
    float s = 0.f;
    for (int i = 0; i < R; ++i)
      for (int j = 0; j < C; ++j) s += v[i][j] * v[i][j];
    return std::sqrt(s);
  }
};

template <int R, int K, int C>
Mat<R, C> operator*(const Mat<R, K>& a, const Mat<K, C>& b) {
  Mat<R, C> o;
  for (int i = 0; i < R; ++i)
    for (int j = 0; j < C; ++j) {
      float s = 0.f;
      for (int k = 0; k < K; ++k) s += a.v[i][k] * b.v[k][j];
      o.v[i][j] = s;
    }
  return o;
}
template <int R, int C>
Mat<R, C> operator-(const Mat<R, C>& a, const Mat<R, C>& b) {
  Mat<R, C> o;
  for (int i = 0; i < R; ++i)
    for (int j = 0; j < C; ++j) o.v[i][j] = a.v[i][j] - b.v[i][j];
  return o;
}

using Matrix4f = Mat<4, 4>;
using Matrix3f = Mat<3, 3>;
using Vector3f = Mat<3, 1>;

}  // namespace Eigen

using vectorMatrix4f = std::vector<Eigen::Matrix4f, Eigen::aligned_allocator<Eigen::Matrix4f>>;  // mycpp/include/Utils.h:39
