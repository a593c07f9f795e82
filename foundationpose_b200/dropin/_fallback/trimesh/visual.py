"""TextureVisuals / ColorVisuals with the attributes Utils.py:104-130 make_mesh_tensors reads."""
import numpy as np


class SimpleMaterial:
    def __init__(self, image=None):
        self.image = image  # PIL.Image


class TextureVisuals:
    kind = "texture"

    def __init__(self, uv=None, image=None, material=None):
        self.uv = None if uv is None else np.asarray(uv, dtype=np.float64)
        self.material = material if material is not None else SimpleMaterial(image)
        self.vertex_colors = None

    def copy(self):
        return TextureVisuals(None if self.uv is None else self.uv.copy(), material=SimpleMaterial(self.material.image))


class ColorVisuals:
    kind = "vertex"

    def __init__(self, vertex_colors=None, n_vertices=0):
        if vertex_colors is None:
            vertex_colors = np.tile(np.array([[102, 102, 102, 255]], dtype=np.uint8), (n_vertices, 1))
        vc = np.asarray(vertex_colors)
        if vc.dtype != np.uint8:
            vc = np.clip(np.rint(vc * (255.0 if vc.max() <= 1.0 else 1.0)), 0, 255).astype(np.uint8)
        if vc.shape[1] == 3:
            vc = np.concatenate([vc, np.full((len(vc), 1), 255, dtype=np.uint8)], 1)
        self.vertex_colors = vc
        self.uv = None
        self.material = None

    def copy(self):
        return ColorVisuals(self.vertex_colors.copy())


class texture:  # `trimesh.visual.texture.TextureVisuals` (Utils.py:661)
    TextureVisuals = TextureVisuals
    SimpleMaterial = SimpleMaterial
