"""Golden vectors for the RENDER GLUE, produced by the reference's own unmodified `nvdiffrast_render` (Utils.py:133-219),
`projection_matrix_from_intrinsics` (:752-802), `make_mesh_tensors` (:104-130), `transform_pts` / `transform_dirs` /
`to_homo_torch` and the module constant `glcam_in_cvcam` (:68) — extracted with `ast` and executed on the CPU.

The only substitution is the `dr` module: nvdiffrast's `rasterize` / `interpolate` / `texture` (absent here) are provided
by tools/nvdiffrast_semantics.py, an INDEPENDENT implementation of their published semantics (float64, exact pixel-centre
sampling in clip space, no snapping).  Everything else — the OpenGL projection with znear 0.001 /
zfar 100, the clip-space crop to the `bbox2d` window, which attribute is interpolated how, `diffuse = clip(n . -light)`,
`color*0.8 + diffuse*color*0.5`, clip, coverage mask, the vertical flips, `extra['xyz_map']` — is the reference's code.

    python tools/make_golden_render.py      # needs /root/reference; writes tests/golden/render_golden.npz

tests/test_render_golden_cpu.py holds oracle.raster.render_crop (and through it the convention the CUDA kernel follows)
to these vectors: coverage, colour, camera-space xyz.
"""
import ast
import logging
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
REF = os.environ.get("FPOSE_REFERENCE", "/root/reference")
H, W, S = 480, 640, 160


def extract_assign(path, name):
    tree = ast.parse(open(path).read())
    node = next(n for n in tree.body if isinstance(n, ast.Assign) and any(isinstance(t, ast.Name) and t.id == name for t in n.targets))
    return compile(ast.fix_missing_locations(ast.Module(body=[node], type_ignores=[])), f"{path}:{name}", "exec")


def cases():
    """(name, mesh, pose, window) — textured and vertex-coloured meshes, centred / off-centre / partly outside windows."""
    from foundationpose_b200 import synth

    rng = np.random.default_rng(5)
    tex_mesh = synth.make_mesh(2, tex_size=64)
    col_mesh = synth.make_mesh(2, tex_size=64)
    vc = rng.integers(0, 256, size=(len(col_mesh.vertices), 4)).astype(np.uint8)
    out = []
    for i, (mesh, colours, t, win) in enumerate((
            (tex_mesh, None, [0.01, -0.01, 0.6], (203.0, 104.0, 462.36877, 362.375)),
            (tex_mesh, None, [0.05, 0.03, 0.45], (230.0, 130.0, 530.0, 430.0)),
            (col_mesh, vc, [-0.02, 0.02, 0.7], (180.0, 140.0, 420.0, 380.0)),
            (tex_mesh, None, [0.12, 0.0, 0.5], (300.0, 120.0, 540.0, 360.0)),   # object partly outside the window
    )):
        pose = np.eye(4)
        pose[:3, :3] = synth.random_rotation(20 + i)
        pose[:3, 3] = t
        out.append((f"case{i}", mesh, colours, pose.astype(np.float32), np.asarray(win, dtype=np.float32)))
    # camera INSIDE the ellipsoid: triangles cross the near plane z = 1 mm (clip-space path; tests/test_raster_gpu.py)
    from oracle import geometry

    def window_of(pose, mesh):
        win, _ = geometry.crop_window(pose[None].astype(np.float32), synth.DEFAULT_K, synth.mesh_diameter(mesh.vertices))
        return np.array([float(v[0]) for v in geometry.render_window(win)], dtype=np.float32)

    pose = np.eye(4)
    pose[:3, :3] = synth.random_rotation(9)
    pose[:3, 3] = [0.004, -0.003, 0.03]
    out.append(("near_plane", tex_mesh, None, pose.astype(np.float32), window_of(pose, tex_mesh)))
    # a bowl seen through its opening: the visible surface consists of BACK faces (nvdiffrast does not cull)
    bowl = synth.make_mesh(2, tex_size=64)
    bowl.faces = bowl.faces[bowl.vertices[bowl.faces].mean(1)[:, 2] < 0.04]
    pose = np.eye(4)
    pose[:3, :3] = np.diag([1.0, -1.0, -1.0]) @ synth.random_rotation(5)[:3, :3]
    pose[:3, 3] = [0.0, 0.0, 0.55]
    out.append(("open_mesh", bowl, None, pose.astype(np.float32), window_of(pose, bowl)))
    return out


def reference_mesh_tensors(ref, mesh, colours):
    from PIL import Image

    from make_golden_geometry import _TextureVisuals

    if colours is None:
        vis = _TextureVisuals()
        vis.uv = mesh.visual.uv
        vis.material = types.SimpleNamespace(image=Image.fromarray(mesh.visual.image))
    else:
        vis = types.SimpleNamespace(vertex_colors=colours)
    tm = types.SimpleNamespace(vertices=mesh.vertices, faces=mesh.faces, vertex_normals=mesh.vertex_normals, visual=vis)
    return ref["make_mesh_tensors"](tm, device="cpu")


def main():
    import nvdiffrast_semantics as dr
    from make_golden_flow import _TorchProxy
    from make_golden_geometry import extract, load_reference_functions

    torch.Tensor.cuda = lambda self, *a, **k: self
    ref = load_reference_functions()  # make_mesh_tensors, projection_matrix_from_intrinsics (+ the geometry functions)
    ns = {"np": np, "torch": _TorchProxy("torch"), "logging": logging, "F": torch.nn.functional, "dr": dr,
          "make_mesh_tensors": ref["make_mesh_tensors"], "projection_matrix_from_intrinsics": ref["projection_matrix_from_intrinsics"]}
    exec(extract_assign(os.path.join(REF, "Utils.py"), "glcam_in_cvcam"), ns)
    for name in ("nvdiffrast_render", "transform_pts", "transform_dirs", "to_homo_torch"):
        exec(extract(os.path.join(REF, "Utils.py"), name), ns)
    from foundationpose_b200 import synth

    K = synth.DEFAULT_K.copy()
    out = {"K": K}
    for name, mesh, colours, pose, win in cases():
        mt = reference_mesh_tensors(ref, mesh, colours)
        extra = {}
        color, depth, normal = ns["nvdiffrast_render"](K=K, H=H, W=W, ob_in_cams=torch.from_numpy(pose)[None], context="cuda", get_normal=False,
                                                      glctx="ctx", mesh_tensors=mt, output_size=(S, S), bbox2d=torch.from_numpy(win)[None], use_light=True,
                                                      extra=extra)
        out[f"{name}.pose"], out[f"{name}.window"] = pose, win
        out[f"{name}.color"] = color[0].numpy()
        out[f"{name}.xyz"] = extra["xyz_map"][0].numpy()
        out[f"{name}.depth_equals_xyz_z"] = np.array(bool(torch.equal(depth[0], extra["xyz_map"][0][..., 2])))
        if colours is not None:
            out[f"{name}.vertex_colors"] = colours
        print(name, "covered pixels", int((extra["xyz_map"][0][..., 2] > 0).sum()))
    dst = os.path.join(ROOT, "tests", "golden", "render_golden.npz")
    np.savez_compressed(dst, **out)
    print(f"wrote {dst}: {len(out)} entries, {os.path.getsize(dst) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
