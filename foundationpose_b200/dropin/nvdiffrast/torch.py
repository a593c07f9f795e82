"""`import nvdiffrast.torch as dr` as the reference does (Utils.py:17).  The drivers only construct a raster context
(run_demo.py:41, run_ycb_video.py:47,92, run_linemod.py:54,99, estimater.py:100,168) and hand it to FoundationPose, which
ignores it: the tiled rasteriser of libfpose.so needs no context."""


class RasterizeCudaContext:
    def __init__(self, device=None):
        self.device = device


class RasterizeGLContext(RasterizeCudaContext):
    def __init__(self, output_db=True, mode="automatic", device=None):
        super().__init__(device)


def _unsupported(name):
    def f(*a, **k):
        raise NotImplementedError(f"nvdiffrast.torch.{name} is not provided: rendering happens inside libfpose.so "
                                  "(use foundationpose_b200.engine.Engine.make_crops)")
    f.__name__ = name
    return f


rasterize = _unsupported("rasterize")
interpolate = _unsupported("interpolate")
texture = _unsupported("texture")
antialias = _unsupported("antialias")
