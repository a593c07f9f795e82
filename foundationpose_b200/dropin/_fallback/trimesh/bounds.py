"""Minimum-volume oriented bounding box (trimesh.bounds.oriented_bounds): for every face of the convex hull take its
normal as one box axis, solve the 2-D minimum-area rectangle of the projected hull (one side collinear with a hull
edge), keep the smallest volume."""
import numpy as np


def _min_rect_2d(pts):
    """pts (n,2) -> (area, rotation angle, (min, max) along the rotated axes)."""
    from scipy.spatial import ConvexHull

    try:
        hull = pts[ConvexHull(pts).vertices]
    except Exception:
        hull = pts
    edges = np.roll(hull, -1, axis=0) - hull
    ang = np.unique(np.mod(np.arctan2(edges[:, 1], edges[:, 0]), np.pi / 2))
    best = None
    for a in ang:
        c, s = np.cos(a), np.sin(a)
        R = np.array([[c, s], [-s, c]])
        q = hull @ R.T
        lo, hi = q.min(0), q.max(0)
        area = float(np.prod(hi - lo))
        if best is None or area < best[0]:
            best = (area, a, lo, hi)
    return best


def oriented_bounds(obj, **kwargs):
    """-> (to_origin (4,4), extents (3,)): `to_origin` moves the mesh so that its oriented box is centred at the origin
    and axis-aligned (the convention run_demo.py:35-36 relies on)."""
    from scipy.spatial import ConvexHull

    pts = np.asarray(getattr(obj, "vertices", obj), dtype=np.float64)
    hull = ConvexHull(pts)
    hp = pts[hull.vertices]
    normals = np.unique(np.round(hull.equations[:, :3], 9), axis=0)
    normals = normals[normals @ np.array([1.0, 1e-3, 1e-6]) >= 0]  # n and -n give the same box
    best = None
    for n in normals:
        n = n / np.linalg.norm(n)
        a = np.cross(n, [1.0, 0, 0] if abs(n[0]) < 0.9 else [0, 1.0, 0])
        a /= np.linalg.norm(a)
        b = np.cross(n, a)
        h = hp @ n
        area, ang, lo, hi = _min_rect_2d(np.stack([hp @ a, hp @ b], 1))
        vol = area * (h.max() - h.min())
        if best is None or vol < best[0]:
            c, s = np.cos(ang), np.sin(ang)
            u, v = c * a + s * b, -s * a + c * b
            best = (vol, np.stack([u, v, n]), np.array([lo[0], lo[1], h.min()]), np.array([hi[0], hi[1], h.max()]))
    _, R, lo, hi = best
    if np.linalg.det(R) < 0:
        R[2], lo[2], hi[2] = -R[2], -hi[2], -lo[2]
    to_origin = np.eye(4)
    to_origin[:3, :3] = R
    to_origin[:3, 3] = -(lo + hi) / 2
    return to_origin, hi - lo
