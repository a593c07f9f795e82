"""Object-initialisation helpers of `FoundationPose.reset_object` (estimater.py:43-76) without open3d / sklearn:
SURVEY.md §8(f) N2.  Host-side numpy, init-time only (never on the per-frame path).

  mesh_diameter      compute_mesh_diameter (Utils.py:559-574).  The reference takes the largest pairwise distance of
                     a RANDOM 10 000-vertex subset (non-deterministic); the maximum over all vertices is attained on
                     the convex hull, so it is computed exactly there — deterministic and an upper bound of any subset.
  voxel_down_sample  open3d PointCloud.voxel_down_sample as used at estimater.py:59-64: points (and normals) falling
                     into the same cubic voxel are averaged; voxel origin = min bound - voxel_size / 2.
"""
import numpy as np


def _max_pairwise(p, chunk=1024):
    """max |a - b| over all pairs: |a|^2 + |b|^2 - 2 a.b by blocked GEMM, then the winning pair re-evaluated directly
    (so the result carries no cancellation error)."""
    p = p - p.mean(axis=0)
    n2 = (p * p).sum(-1)
    best, arg = -1.0, (0, 0)
    for i in range(0, len(p), chunk):
        d2 = n2[i:i + chunk, None] + n2[None, :] - 2.0 * (p[i:i + chunk] @ p.T)
        k = int(d2.argmax())
        if d2.flat[k] > best:
            best = float(d2.flat[k])
            arg = (i + k // len(p), k % len(p))
    return float(np.linalg.norm(p[arg[0]] - p[arg[1]]))


def mesh_diameter(vertices):
    """Largest distance between two vertices (exact).  Pruning: with c the centroid and r_i = |v_i - c|, the pair
    (a, b) that attains the maximum d* satisfies r_a + r_b >= d* >= L for any lower bound L (here: a double sweep),
    hence r_a >= L - max r; only those candidates are searched exhaustively."""
    v = np.asarray(vertices, dtype=np.float64).reshape(-1, 3)
    if len(v) < 2:
        return 0.0
    v = v - v.mean(axis=0)
    r = np.linalg.norm(v, axis=1)
    a = int(r.argmax())
    b = int(np.linalg.norm(v - v[a], axis=1).argmax())
    lower = float(np.linalg.norm(v - v[b], axis=1).max())
    cand = v[r >= lower - r.max() - 1e-12]
    if len(cand) > 4096:  # near-spherical cloud: the hull cannot help either; search the candidates in blocks
        return max(lower, _max_pairwise(cand))
    return max(lower, _max_pairwise(cand)) if len(cand) >= 2 else lower


def voxel_down_sample(points, voxel_size, normals=None):
    """-> (points', normals' or None): per-voxel means, voxels ordered by their (ix, iy, iz) index."""
    p = np.asarray(points, dtype=np.float64).reshape(-1, 3)
    if voxel_size <= 0:
        raise ValueError("voxel_size must be positive")
    if len(p) == 0:
        return p.copy(), (None if normals is None else np.asarray(normals, dtype=np.float64).reshape(-1, 3).copy())
    origin = p.min(axis=0) - voxel_size * 0.5
    idx = np.floor((p - origin) / voxel_size).astype(np.int64)
    _, inv, counts = np.unique(idx, axis=0, return_inverse=True, return_counts=True)
    inv = inv.reshape(-1)
    out_p = np.zeros((len(counts), 3))
    np.add.at(out_p, inv, p)
    out_p /= counts[:, None]
    out_n = None
    if normals is not None:
        n = np.asarray(normals, dtype=np.float64).reshape(-1, 3)
        out_n = np.zeros((len(counts), 3))
        np.add.at(out_n, inv, n)
        out_n /= counts[:, None]
    return out_p, out_n
