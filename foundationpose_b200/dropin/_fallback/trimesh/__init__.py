"""Minimal stand-in for trimesh, used ONLY when the real package is not installed (it is a pure init-time / IO
dependency of the reference: SURVEY.md Appendix C).  Covers what the reference's drivers and estimator touch:

    trimesh.load(path)                      Wavefront OBJ (+ MTL `map_Kd` texture) or Stanford PLY (BOP models) -> Trimesh
    Trimesh(vertices, faces, ...)           .vertices .faces .vertex_normals .visual .copy() .export() .apply_transform()
    mesh.visual.uv / .material.image        TextureVisuals (PIL image), or mesh.visual.vertex_colors (ColorVisuals)
    trimesh.bounds.oriented_bounds(mesh)    minimum-volume oriented box (run_demo.py:35)
    trimesh.creation.icosphere(...)         Utils.py:485-489
    trimesh.primitives.Box(extents, transform)  run_ycb_video.py:93, run_linemod.py:100

Not a general mesh library: no repair, no merging beyond the (position, uv) de-duplication an OBJ needs.
"""
import os

import numpy as np

from . import bounds, creation, primitives, visual  # noqa: F401
from .base import Trimesh  # noqa: F401
from .exchange import load, load_obj, load_ply  # noqa: F401

__version__ = "0.0-fpose-b200-standin"
