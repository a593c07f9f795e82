from . import creation


def Box(extents=(1, 1, 1), transform=None, **kwargs):
    """trimesh.primitives.Box(extents=..., transform=...) as the dataset drivers build their placeholder mesh."""
    return creation.box(extents, transform)
