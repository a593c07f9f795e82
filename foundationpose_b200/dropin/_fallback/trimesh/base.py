import numpy as np

from . import visual as _visual


class Trimesh:
    def __init__(self, vertices=None, faces=None, vertex_normals=None, visual=None, process=False, **kwargs):
        self.vertices = np.asarray(vertices, dtype=np.float64).reshape(-1, 3)
        self.faces = np.asarray(faces, dtype=np.int64).reshape(-1, 3)
        self._vn = None if vertex_normals is None else np.asarray(vertex_normals, dtype=np.float64).reshape(-1, 3)
        self._vn_for = self.vertices.shape
        self.visual = visual if visual is not None else _visual.ColorVisuals(n_vertices=len(self.vertices))

    # ---- geometry
    @property
    def face_normals(self):
        v = self.vertices[self.faces]
        n = np.cross(v[:, 1] - v[:, 0], v[:, 2] - v[:, 0])
        l = np.linalg.norm(n, axis=1, keepdims=True)
        return n / np.maximum(l, 1e-300)

    @property
    def vertex_normals(self):
        """File normals when the OBJ had them, else the angle-weighted mean of the adjacent face normals."""
        if self._vn is not None and len(self._vn) == len(self.vertices):
            return self._vn
        v = self.vertices[self.faces]
        fn = self.face_normals
        out = np.zeros_like(self.vertices)
        for k in range(3):
            a = v[:, (k + 1) % 3] - v[:, k]
            b = v[:, (k + 2) % 3] - v[:, k]
            cosang = (a * b).sum(1) / np.maximum(np.linalg.norm(a, axis=1) * np.linalg.norm(b, axis=1), 1e-300)
            np.add.at(out, self.faces[:, k], fn * np.arccos(np.clip(cosang, -1, 1))[:, None])
        return out / np.maximum(np.linalg.norm(out, axis=1, keepdims=True), 1e-300)

    @vertex_normals.setter
    def vertex_normals(self, v):
        self._vn = np.asarray(v, dtype=np.float64).reshape(-1, 3)

    @property
    def bounds(self):
        return np.stack([self.vertices.min(0), self.vertices.max(0)])

    @property
    def extents(self):
        return self.vertices.max(0) - self.vertices.min(0)

    @property
    def centroid(self):
        return self.vertices.mean(0)

    def copy(self):
        return Trimesh(self.vertices.copy(), self.faces.copy(), None if self._vn is None else self._vn.copy(), self.visual.copy())

    def apply_transform(self, tf):
        tf = np.asarray(tf, dtype=np.float64)
        self.vertices = self.vertices @ tf[:3, :3].T + tf[:3, 3]
        if self._vn is not None:
            n = self._vn @ np.linalg.inv(tf[:3, :3])  # (M^-T n)^T = n^T M^-1
            self._vn = n / np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-300)
        return self

    def to_mesh(self):
        return self

    def export(self, path):
        from .exchange import export_obj, export_ply

        if str(path).lower().endswith(".ply"):
            return export_ply(self, path)
        return export_obj(self, path)
