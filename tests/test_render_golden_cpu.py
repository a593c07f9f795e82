"""CPU: the oracle rasteriser (oracle/raster.py::render_crop — the rule the CUDA crop producer is tested against) vs the
reference's own unmodified `nvdiffrast_render` (Utils.py:133-219) + `projection_matrix_from_intrinsics` +
`make_mesh_tensors`, executed by tools/make_golden_render.py with nvdiffrast's three primitives supplied by an independent
float64 implementation of their published semantics (tools/nvdiffrast_semantics.py: clip-space input, exact pixel-centre
sampling, no snapping).  Pins everything the reference's function does AROUND those primitives — OpenGL projection with
znear / zfar, the clip-space crop to the bbox2d window, attribute routing, uv flip + wrap texture, the Lambert term,
0.8 / 0.5 weights, clipping, masking, vertical flips — for textured and vertex-coloured meshes, centred and clipped windows,
the camera INSIDE the object (triangles crossing the near plane z = 1 mm) and an open mesh seen from its back faces.

Measured: coverage identical in all six cases (48 370 covered pixels); camera-space xyz within 5e-5 m at 99.9 % of the
pixels (a handful on triangle edges up to 9e-4: the oracle snaps vertices to 1/256 px, the independent rasteriser does
not); colour within 7e-4 at 99 % (texture-edge pixels up to 0.06)."""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_render_crop_matches_the_reference_nvdiffrast_render_glue():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_golden_render as gen  # cases() only: nothing of the reference is read here

    from oracle import pipeline, raster

    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "render_golden.npz")))
    K = g["K"]
    total = 0
    for name, mesh, colours, pose, win in gen.cases():
        assert np.array_equal(pose, g[f"{name}.pose"]) and np.array_equal(win, g[f"{name}.window"])
        assert bool(g[f"{name}.depth_equals_xyz_z"])
        if colours is not None:
            mesh = types.SimpleNamespace(vertices=mesh.vertices, faces=mesh.faces, vertex_normals=mesh.vertex_normals,
                                         visual=types.SimpleNamespace(vertex_colors=colours, uv=None, image=None))
        rgb, xyz, _ = raster.render_crop(pose, pipeline.mesh_tensors(mesh), K, tuple(win))
        want_rgb, want_xyz = g[f"{name}.color"], g[f"{name}.xyz"]
        cov, want_cov = xyz[..., 2] > 0, want_xyz[..., 2] > 0
        assert (cov != want_cov).mean() <= 1e-3, name                         # measured: 0 pixels
        assert (rgb[~cov] == 0).all() and (want_rgb[~want_cov] == 0).all()   # background = 0 on both sides
        both = cov & want_cov
        total += int(both.sum())
        dx, dc = np.abs(xyz - want_xyz)[both], np.abs(rgb - want_rgb)[both]
        print(f"{name}: {int(both.sum())} px, xyz max {dx.max():.2e} p99.9 {np.quantile(dx, 0.999):.2e}, rgb max {dc.max():.2e} p99 {np.quantile(dc, 0.99):.2e}")
        assert np.quantile(dx, 0.999) < 1e-4 and dx.max() < 2e-3, name
        assert np.quantile(dc, 0.99) < 2e-3 and dc.max() < 0.1, name
    assert total > 40000
