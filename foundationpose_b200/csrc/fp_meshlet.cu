// fp_meshlet.cu — host-side mesh preparation for the tiled crop producer (init time, fp_set_mesh):
//   * packs the vertex attributes of Utils.py:104-130 make_mesh_tensors into 16-byte records;
//   * decides whether back faces can be culled: the mesh must be CLOSED and CONSISTENTLY ORIENTED (every directed edge
//     of the position-welded mesh occurs exactly once and its reverse exactly once); the sign of the enclosed volume
//     tells which screen-space winding is front-facing.  nvdiffrast (Utils.py:182) renders both sides, so an open mesh
//     keeps both sides here too;
//   * builds meshlets: faces are sorted by (normal octant, Morton code of the centroid) and chunked greedily into
//     groups of <= 64 triangles / <= 64 unique vertices, each with a bounding sphere and a normal cone.
// Pure host code (no kernels); compiled with the rest of the library.
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <numeric>
#include <unordered_map>

#include "fp_common.cuh"
#include "fp_crop.cuh"

namespace fp {

namespace {

struct V3 {
  double x, y, z;
};
inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline double norm(V3 a) { return sqrt(dot(a, a)); }

inline uint32_t part1by2(uint32_t x) {  // spread the low 10 bits
  x &= 0x3ff;
  x = (x | (x << 16)) & 0x30000ff;
  x = (x | (x << 8)) & 0x300f00f;
  x = (x | (x << 4)) & 0x30c30c3;
  x = (x | (x << 2)) & 0x9249249;
  return x;
}

struct PosKey {
  uint32_t a, b, c;
  bool operator==(const PosKey& o) const { return a == o.a && b == o.b && c == o.c; }
};
struct PosHash {
  size_t operator()(const PosKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (uint32_t w : {k.a, k.b, k.c}) {
      h ^= w;
      h *= 1099511628211ull;
    }
    return (size_t)h;
  }
};

}  // namespace

int build_mesh_host(int V, int F, const float* pos, const float* nrm, const float* att, int n_att, const int* faces,
                    MeshHost& out) {
  out.vpos.resize(V);
  out.vnrm.resize(V);
  out.vatt.resize(V);
  for (int v = 0; v < V; ++v) {
    out.vpos[v] = make_float4(pos[3 * v], pos[3 * v + 1], pos[3 * v + 2], 0.f);
    out.vnrm[v] = make_float4(nrm[3 * v], nrm[3 * v + 1], nrm[3 * v + 2], 0.f);
    out.vatt[v] = n_att == 2 ? make_float4(att[2 * v], att[2 * v + 1], 0.f, 0.f)
                             : make_float4(att[3 * v], att[3 * v + 1], att[3 * v + 2], 0.f);
  }
  out.faces.resize(F);
  for (int f = 0; f < F; ++f) out.faces[f] = make_int4(faces[3 * f], faces[3 * f + 1], faces[3 * f + 2], 0);

  auto P = [&](int v) { return V3{pos[3 * v], pos[3 * v + 1], pos[3 * v + 2]}; };

  // ---- closedness / orientation on the position-welded mesh (texture seams duplicate vertices)
  std::vector<int> canon(V);
  {
    std::unordered_map<PosKey, int, PosHash> seen;
    seen.reserve((size_t)V * 2);
    for (int v = 0; v < V; ++v) {
      float q[3] = {pos[3 * v] + 0.f, pos[3 * v + 1] + 0.f, pos[3 * v + 2] + 0.f};  // +0 folds -0 into +0
      PosKey k;
      memcpy(&k.a, q, 4);
      memcpy(&k.b, q + 1, 4);
      memcpy(&k.c, q + 2, 4);
      auto it = seen.find(k);
      if (it == seen.end()) {
        seen.emplace(k, v);
        canon[v] = v;
      } else {
        canon[v] = it->second;
      }
    }
  }
  bool closed = true;
  {
    std::unordered_map<uint64_t, int> edges;  // directed edge (a, b) -> count
    edges.reserve((size_t)F * 4);
    for (int f = 0; f < F && closed; ++f) {
      const int i[3] = {canon[faces[3 * f]], canon[faces[3 * f + 1]], canon[faces[3 * f + 2]]};
      if (i[0] == i[1] || i[1] == i[2] || i[0] == i[2]) continue;  // degenerate after welding: covers nothing
      for (int e = 0; e < 3; ++e) {
        const uint64_t key = ((uint64_t)(uint32_t)i[e] << 32) | (uint32_t)i[(e + 1) % 3];
        if (++edges[key] > 1) closed = false;  // an edge used twice in the same direction: inconsistent winding
      }
    }
    if (closed)
      for (auto& kv : edges) {
        const uint64_t rev = (kv.first << 32) | (kv.first >> 32);
        auto it = edges.find(rev);
        if (it == edges.end() || it->second != 1) {
          closed = false;
          break;
        }
      }
  }
  double vol6 = 0.0;
  for (int f = 0; f < F; ++f) vol6 += dot(P(faces[3 * f]), cross(P(faces[3 * f + 1]), P(faces[3 * f + 2])));
  out.closed = closed ? 1 : 0;
  // camera looks down +z with x right / y down (OpenCV): a triangle whose outward normal faces the camera has a
  // NEGATIVE signed screen area (x1-x0)(y2-y0) - (y1-y0)(x2-x0); inside-out meshes (negative volume) flip that
  out.front_sign = !closed || vol6 == 0.0 ? 0 : (vol6 > 0.0 ? -1 : 1);

  // ---- sort faces: (normal octant, Morton code of the centroid)
  V3 lo = {1e300, 1e300, 1e300}, hi = {-1e300, -1e300, -1e300};
  for (int v = 0; v < V; ++v) {
    const V3 p = P(v);
    lo = {std::min(lo.x, p.x), std::min(lo.y, p.y), std::min(lo.z, p.z)};
    hi = {std::max(hi.x, p.x), std::max(hi.y, p.y), std::max(hi.z, p.z)};
  }
  const double ext = std::max({hi.x - lo.x, hi.y - lo.y, hi.z - lo.z, 1e-30});
  {
    const V3 c = {(lo.x + hi.x) / 2, (lo.y + hi.y) / 2, (lo.z + hi.z) / 2};
    double r = 0;
    for (int v = 0; v < V; ++v) r = std::max(r, norm(sub(P(v), c)));
    out.bs[0] = (float)c.x;
    out.bs[1] = (float)c.y;
    out.bs[2] = (float)c.z;
    out.bs[3] = (float)(r * 1.0001 + 1e-9);
  }
  std::vector<V3> fn(F);
  std::vector<uint64_t> key(F);
  for (int f = 0; f < F; ++f) {
    const V3 a = P(faces[3 * f]), b = P(faces[3 * f + 1]), c = P(faces[3 * f + 2]);
    V3 n = cross(sub(b, a), sub(c, a));
    const double l = norm(n);
    fn[f] = l > 0 ? V3{n.x / l, n.y / l, n.z / l} : V3{0, 0, 0};
    const V3 ctr = {(a.x + b.x + c.x) / 3, (a.y + b.y + c.y) / 3, (a.z + b.z + c.z) / 3};
    const uint32_t qx = (uint32_t)std::min(1023.0, (ctr.x - lo.x) / ext * 1023.0);
    const uint32_t qy = (uint32_t)std::min(1023.0, (ctr.y - lo.y) / ext * 1023.0);
    const uint32_t qz = (uint32_t)std::min(1023.0, (ctr.z - lo.z) / ext * 1023.0);
    const uint64_t morton = part1by2(qx) | (part1by2(qy) << 1) | (part1by2(qz) << 2);
    const uint64_t oct = out.front_sign ? (uint64_t)((fn[f].x < 0) | ((fn[f].y < 0) << 1) | ((fn[f].z < 0) << 2)) : 0;
    key[f] = (oct << 32) | morton;
  }
  std::vector<int> order(F);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key[a] < key[b]; });

  // ---- greedy chunking
  out.meshlets.clear();
  out.ml_verts.clear();
  out.ml_tris.clear();
  std::vector<int> slot_of(V, -1);
  std::vector<int> cur_verts;
  std::vector<int> cur_faces;
  std::vector<uint2> cur_tris;
  auto flush = [&]() {
    if (cur_faces.empty()) return;
    Meshlet m;
    m.vert_off = (int)out.ml_verts.size();
    m.n_verts = (int)cur_verts.size();
    m.tri_off = (int)out.ml_tris.size();
    m.n_tris = (int)cur_faces.size();
    V3 c = {0, 0, 0};
    for (int v : cur_verts) {
      const V3 p = P(v);
      c = {c.x + p.x, c.y + p.y, c.z + p.z};
    }
    c = {c.x / m.n_verts, c.y / m.n_verts, c.z / m.n_verts};
    double r = 0;
    for (int v : cur_verts) r = std::max(r, norm(sub(P(v), c)));
    m.cx = (float)c.x;
    m.cy = (float)c.y;
    m.cz = (float)c.z;
    m.r = (float)(r * 1.0001 + 1e-9);
    V3 ax = {0, 0, 0};
    bool degenerate = false;
    for (int f : cur_faces) {
      if (dot(fn[f], fn[f]) == 0) degenerate = true;
      ax = {ax.x + fn[f].x, ax.y + fn[f].y, ax.z + fn[f].z};
    }
    const double al = norm(ax);
    m.ax = m.ay = m.az = 0.f;
    m.cutoff = -2.f;
    if (out.front_sign && !degenerate && al > 1e-9) {
      ax = {ax.x / al, ax.y / al, ax.z / al};
      double mn = 1.0;
      for (int f : cur_faces) mn = std::min(mn, dot(ax, fn[f]));
      m.ax = (float)ax.x;
      m.ay = (float)ax.y;
      m.az = (float)ax.z;
      m.cutoff = (float)mn;
    }
    for (int v : cur_verts) {
      out.ml_verts.push_back(v);
      slot_of[v] = -1;
    }
    for (const uint2& t : cur_tris) out.ml_tris.push_back(t);
    out.meshlets.push_back(m);
    cur_verts.clear();
    cur_faces.clear();
    cur_tris.clear();
  };
  uint64_t cur_oct = ~0ull;
  for (int k = 0; k < F; ++k) {
    const int f = order[k];
    const int i[3] = {faces[3 * f], faces[3 * f + 1], faces[3 * f + 2]};
    const uint64_t oct = key[f] >> 32;
    int fresh = 0;
    for (int e = 0; e < 3; ++e) {
      bool dup = false;
      for (int e2 = 0; e2 < e; ++e2) dup = dup || i[e2] == i[e];
      if (slot_of[i[e]] < 0 && !dup) ++fresh;
    }
    if (oct != cur_oct || (int)cur_faces.size() == kMeshletTris || (int)cur_verts.size() + fresh > kMeshletVerts) flush();
    cur_oct = oct;
    uint32_t packed = 0;
    for (int e = 0; e < 3; ++e) {
      if (slot_of[i[e]] < 0) {
        slot_of[i[e]] = (int)cur_verts.size();
        cur_verts.push_back(i[e]);
      }
      packed |= (uint32_t)slot_of[i[e]] << (8 * e);
    }
    cur_tris.push_back(make_uint2(packed, (unsigned)f));
    cur_faces.push_back(f);
  }
  flush();
  return 0;
}

}  // namespace fp
