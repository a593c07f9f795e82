import numpy as np

from .base import Trimesh


def icosphere(subdivisions=3, radius=1.0, **kwargs):
    """Unit icosahedron, each triangle split in four per subdivision, vertices pushed to the sphere."""
    from foundationpose_b200.synth import icosphere as _ico

    v, f = _ico(subdivisions)
    return Trimesh(v * radius, f, v.copy())


def box(extents=(1, 1, 1), transform=None):
    e = np.asarray(extents, dtype=np.float64) / 2
    v = np.array([[x, y, z] for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)], dtype=np.float64) * e
    f = np.array([[0, 1, 3], [0, 3, 2], [4, 6, 7], [4, 7, 5], [0, 4, 5], [0, 5, 1], [2, 3, 7], [2, 7, 6], [0, 2, 6], [0, 6, 4],
                  [1, 5, 7], [1, 7, 3]], dtype=np.int64)
    m = Trimesh(v, f)
    if transform is not None:
        m.apply_transform(transform)
    return m
