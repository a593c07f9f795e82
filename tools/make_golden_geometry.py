"""Golden vectors for the geometry around the networks, produced by the REFERENCE'S OWN CODE.

`Utils.py`, `estimater.py` and `learning/datasets/h5_dataset.py` cannot be imported (their module-level imports need
nvdiffrast, pytorch3d, kornia, warp, open3d ...), but several hot-path functions in them are pure numpy / torch.
This script parses those files with `ast`, extracts the unmodified source of exactly these functions, executes them
on the CPU (`Tensor.cuda()` and `torch.set_default_tensor_type` neutralised, nothing else touched) on seeded inputs,
and writes inputs + outputs to tests/golden/geometry_golden.npz:

  compute_crop_window_tf_batch (box_3d)     Utils.py:577-621            -> oracle.geometry.crop_window
  depth2xyzmap / depth2xyzmap_batch         Utils.py:399-438            -> oracle.geometry.depth2xyzmap
  egocentric_delta_pose_to_pose             Utils.py:848-855            -> oracle.geometry.pose_update (pose composition)
  FoundationPose.guess_translation          estimater.py:137-156        -> hypotheses.guess_translation, start_poses_kernel
  PairH5Dataset.transform_depth_to_xyzmap   h5_dataset.py:79-114        -> oracle.geometry.normalise_xyz (tau = 0.001)
  TripletH5Dataset.transform_depth_to_xyzmap h5_dataset.py:137-170      -> oracle.geometry.normalise_xyz (tau = 0.1)
  erode_depth_kernel / bilateral_filter_depth_kernel (Warp kernel bodies run as plain Python, Utils.py:304-384)
                                                                         -> oracle.geometry.erode_depth / bilateral_filter_depth
  make_mesh_tensors                          Utils.py:104-130            -> estimater.make_mesh_tensors (product host code)
  projection_matrix_from_intrinsics + the bbox2d crop of nvdiffrast_render (Utils.py:159-181, 752-802)
                                                                         -> the pixel mapping oracle/raster.py assumes

Run here (needs /root/reference):  python tools/make_golden_geometry.py
tests/test_oracle_golden.py holds the oracle (and the product's host code) to these vectors.
"""
import ast
import logging
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("FPOSE_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)


def extract(path, name, cls=None):
    """Source-exact FunctionDef `name` (inside class `cls` if given) of a reference file, compiled on its own."""
    src = open(path).read()
    tree = ast.parse(src)
    body = tree.body
    if cls is not None:
        body = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls).body
    fn = next((n for n in body if isinstance(n, ast.FunctionDef) and n.name == name), None)
    if fn is None and cls is None:  # e.g. the Warp kernels, nested under `if wp is not None:`
        fn = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == name)
    fn.decorator_list = []
    for a in list(fn.args.args) + list(fn.args.kwonlyargs):  # annotations name types we do not import
        a.annotation = None
    fn.returns = None
    mod = ast.Module(body=[fn], type_ignores=[])
    return compile(ast.fix_missing_locations(mod), f"{path}:{name}", "exec")


class _TorchProxy(types.ModuleType):
    """`torch` with the one global side effect of compute_crop_window_tf_batch removed."""

    def __getattr__(self, k):
        return getattr(torch, k)

    @staticmethod
    def set_default_tensor_type(*a, **k):
        return None


class _TextureVisuals:
    """Stands in for trimesh.visual.texture.TextureVisuals in make_mesh_tensors' isinstance check."""


_TRIMESH = types.SimpleNamespace(visual=types.SimpleNamespace(texture=types.SimpleNamespace(TextureVisuals=_TextureVisuals)))


class _WarpStub:
    """Just enough of `warp` to run a kernel body as ordinary Python, one thread at a time (Warp computes in fp32, this
    runs in fp64: the goldens are compared with 2e-6)."""
    _tid = (0, 0)
    exp = staticmethod(__import__("math").exp)

    @staticmethod
    def tid():
        return _WarpStub._tid

    @staticmethod
    def launch(kernel, H, W, *args):
        for h in range(H):
            for w in range(W):
                _WarpStub._tid = (h, w)
                kernel(*args)


def load_reference_functions():
    ns = {"np": np, "torch": _TorchProxy("torch"), "logging": logging, "kornia": None, "F": torch.nn.functional, "wp": _WarpStub, "trimesh": _TRIMESH}
    for name in ("compute_crop_window_tf_batch", "depth2xyzmap", "depth2xyzmap_batch", "egocentric_delta_pose_to_pose",
                 "projection_matrix_from_intrinsics"):
        exec(extract(os.path.join(REF, "Utils.py"), name), ns)
    exec(extract(os.path.join(REF, "Utils.py"), "make_mesh_tensors"), ns)
    for name in ("erode_depth_kernel", "bilateral_filter_depth_kernel"):  # Warp kernel bodies, run as plain Python
        exec(extract(os.path.join(REF, "Utils.py"), name), ns)
    exec(extract(os.path.join(REF, "estimater.py"), "guess_translation", cls="FoundationPose"), ns)
    ns_t = dict(ns)
    exec(extract(os.path.join(REF, "learning/datasets/h5_dataset.py"), "transform_depth_to_xyzmap", cls="TripletH5Dataset"), ns_t)
    ns["transform_depth_to_xyzmap_scorer"] = ns_t["transform_depth_to_xyzmap"]
    exec(extract(os.path.join(REF, "learning/datasets/h5_dataset.py"), "transform_depth_to_xyzmap", cls="PairH5Dataset"), ns)
    return ns


def main():
    from foundationpose_b200 import hypotheses, synth

    torch.Tensor.cuda = lambda self, *a, **k: self  # the reference hard-codes .cuda(); run it where we are
    ref = load_reference_functions()
    rng = np.random.default_rng(7)
    out = {}

    # ---- G1: crop windows of 252 seeded hypotheses
    K = synth.DEFAULT_K.copy()
    poses = hypotheses.make_rotation_grid().astype(np.float32)
    poses[:, :3, 3] = (np.array([0.02, -0.01, 0.6]) + rng.normal(0, [0.08, 0.06, 0.12], size=(len(poses), 3))).astype(np.float32)
    d = 0.19
    tf = ref["compute_crop_window_tf_batch"](pts=None, H=480, W=640, poses=torch.from_numpy(poses), K=torch.from_numpy(K).float(),
                                            crop_ratio=1.2, out_size=(160, 160), method="box_3d", mesh_diameter=d)
    out.update(cw_poses=poses, cw_K=K.astype(np.float32), cw_diameter=np.float32(d), cw_tf=tf.numpy().astype(np.float32))

    # ---- depth2xyzmap (numpy) and depth2xyzmap_batch (torch, zfar)
    depth = rng.uniform(0.0, 2.0, size=(48, 64)).astype(np.float32)
    depth[rng.random(depth.shape) < 0.2] = 0.0
    depth[rng.random(depth.shape) < 0.05] = 0.0005
    Ks = np.array([[61.5, 0, 32.0], [0, 61.5, 24.0], [0, 0, 1]], dtype=np.float32)
    xyz = ref["depth2xyzmap"](depth, Ks)
    xyz_b = ref["depth2xyzmap_batch"](torch.from_numpy(depth)[None].clone(), torch.from_numpy(Ks)[None], zfar=1.5)
    xyz_inf = ref["depth2xyzmap_batch"](torch.from_numpy(depth)[None].clone(), torch.from_numpy(Ks)[None], zfar=np.inf)
    out.update(dx_depth=depth, dx_K=Ks, dx_xyz=xyz.astype(np.float32), dx_xyz_zfar15=xyz_b[0].numpy(), dx_xyz_inf=xyz_inf[0].numpy())

    # ---- pose composition of the refiner update
    A = torch.from_numpy(poses[:16].copy())
    td = torch.from_numpy(rng.normal(0, 0.02, size=(16, 3)).astype(np.float32))
    rv = rng.normal(0, 0.2, size=(16, 3))
    Rd = []
    for v in rv:  # any rotation matrices (Rodrigues in float64, then float32)
        th = np.linalg.norm(v)
        k = v / th
        Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        Rd.append(np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx)
    Rd = torch.from_numpy(np.stack(Rd).astype(np.float32))
    B = ref["egocentric_delta_pose_to_pose"](A, trans_delta=td, rot_mat_delta=Rd)
    out.update(pu_A=A.numpy(), pu_trans_delta=td.numpy(), pu_rot_delta=Rd.numpy(), pu_B=B.numpy())

    # ---- guess_translation: odd / even valid counts, empty mask, no valid depth
    mesh = synth.make_mesh(3)
    pose = np.eye(4)
    pose[:3, :3] = synth.random_rotation(0)
    pose[:3, 3] = [0.02, -0.01, 0.6]
    _, dep, mask = synth.make_scene(mesh.visual.image, pose)
    dep, mask = dep[::4, ::4].copy(), mask[::4, ::4].copy()  # 120 x 160 keeps the fixture small
    Kq = K.copy()
    Kq[:2] /= 4.0
    dep[rng.random(dep.shape) < 0.1] = 0
    masks = [mask.copy(), mask.copy(), np.zeros_like(mask), mask.copy()]
    vs, us = np.where(mask)
    masks[1][vs[0], us[0]] = False
    deps = [dep, dep, dep, np.zeros_like(dep)]
    fake_self = types.SimpleNamespace(debug=0, debug_dir="/tmp")
    gts = [np.asarray(ref["guess_translation"](fake_self, depth=dd, mask=mm, K=Kq), dtype=np.float64) for dd, mm in zip(deps, masks)]
    out.update(gt_depth=dep, gt_masks=np.stack(masks), gt_zero_depth_case=np.array([0, 0, 0, 1], dtype=np.int32), gt_K=Kq,
               gt_t=np.stack(gts))

    # ---- refiner xyz normalisation (transform_depth_to_xyzmap with ready xyz maps: the predictor's path)
    bs = 3
    xyzA = rng.normal(0, 0.3, size=(bs, 3, 20, 20)).astype(np.float32)
    xyzA[:, 2] += 0.6
    xyzA[:, :, :3] = 0  # invalid (z < 0.001)
    xyzB = rng.normal(0, 0.3, size=(bs, 3, 20, 20)).astype(np.float32)
    xyzB[:, 2] += 0.6
    xyzB[:, 2, 5:8] = 0.0005
    tA = np.array([[0.0, 0.0, 0.6], [0.05, -0.02, 0.7], [-0.1, 0.1, 0.5]], dtype=np.float32)
    poseA = np.tile(np.eye(4, dtype=np.float32), (bs, 1, 1))
    poseA[:, :3, 3] = tA
    batch = types.SimpleNamespace(rgbAs=torch.zeros(bs, 3, 20, 20), mesh_diameters=torch.full((bs,), 0.19), tf_to_crops=torch.eye(3)[None].repeat(bs, 1, 1),
                                  poseA=torch.from_numpy(poseA), Ks=torch.eye(3)[None].repeat(bs, 1, 1), xyz_mapAs=torch.from_numpy(xyzA.copy()),
                                  xyz_mapBs=torch.from_numpy(xyzB.copy()), depthAs=None, depthBs=None)
    fake_ds = types.SimpleNamespace(cfg={"normalize_xyz": True})
    nb = ref["transform_depth_to_xyzmap"](fake_ds, batch, 480, 640)
    out.update(nx_xyzA=xyzA, nx_xyzB=xyzB, nx_t=tA, nx_diameter=np.float32(0.19), nx_outA=nb.xyz_mapAs.numpy(), nx_outB=nb.xyz_mapBs.numpy())

    # ---- scorer xyz normalisation (TripletH5Dataset.transform_depth_to_xyzmap, tau = 0.1; ready xyz maps)
    xyzA2 = xyzA.copy()
    xyzA2[:, 2, 10:12] = 0.05  # 0.001 <= z < 0.1: invalid for the scorer only
    batch2 = types.SimpleNamespace(rgbAs=torch.zeros(bs, 3, 20, 20), mesh_diameters=torch.full((bs,), 0.19), tf_to_crops=torch.eye(3)[None].repeat(bs, 1, 1),
                                   poseA=torch.from_numpy(poseA), Ks=torch.eye(3)[None].repeat(bs, 1, 1), xyz_mapAs=torch.from_numpy(xyzA2.copy()),
                                   xyz_mapBs=torch.from_numpy(xyzB.copy()), depthAs=None, depthBs=None)
    nb2 = ref["transform_depth_to_xyzmap_scorer"](fake_ds, batch2, 480, 640)
    out.update(ns_xyzA=xyzA2, ns_outA=nb2.xyz_mapAs.numpy(), ns_outB=nb2.xyz_mapBs.numpy())

    # ---- depth filters: the reference's Warp kernel bodies executed thread by thread
    dimg = (0.6 + 0.05 * np.sin(np.arange(64)[None] / 7.0) + 0.03 * np.cos(np.arange(48)[:, None] / 5.0)).astype(np.float32)
    dimg += rng.normal(0, 0.0005, dimg.shape).astype(np.float32)
    dimg[rng.random(dimg.shape) < 0.06] = 0.0          # holes
    dimg[10:20, 30:40] += 0.05                          # a step edge
    dimg[rng.random(dimg.shape) < 0.02] += 0.2          # outliers
    er = np.zeros_like(dimg)
    _WarpStub.launch(ref["erode_depth_kernel"], 48, 64, dimg, er, 2, 0.001, 0.8, 100.0)
    bl = np.zeros_like(dimg)
    _WarpStub.launch(ref["bilateral_filter_depth_kernel"], 48, 64, er, bl, 2, 100.0, 2.0, 100000.0)
    out.update(df_depth=dimg, df_eroded=er, df_bilateral=bl)

    # ---- make_mesh_tensors (Utils.py:104-130) on a textured and on a vertex-coloured mesh
    from PIL import Image

    m2 = synth.make_mesh(1, tex_size=16)
    vis = _TextureVisuals()
    vis.uv = m2.visual.uv
    vis.material = types.SimpleNamespace(image=Image.fromarray(m2.visual.image))
    tmesh = types.SimpleNamespace(vertices=m2.vertices, faces=m2.faces, vertex_normals=m2.vertex_normals, visual=vis)
    mt = ref["make_mesh_tensors"](tmesh, device="cpu")
    vc = rng.integers(0, 256, size=(len(m2.vertices), 4)).astype(np.uint8)
    cmesh = types.SimpleNamespace(vertices=m2.vertices, faces=m2.faces, vertex_normals=m2.vertex_normals,
                                  visual=types.SimpleNamespace(vertex_colors=vc))
    mc = ref["make_mesh_tensors"](cmesh, device="cpu")
    out.update(mt_vertices=m2.vertices, mt_faces=m2.faces, mt_normals=m2.vertex_normals, mt_uv_in=m2.visual.uv, mt_tex_in=m2.visual.image,
               mt_pos=mt["pos"].numpy(), mt_faces_out=mt["faces"].numpy(), mt_vnormals=mt["vnormals"].numpy(), mt_uv=mt["uv"].numpy(),
               mt_tex=mt["tex"].numpy(), mt_vcolor_in=vc, mt_vertex_color=mc["vertex_color"].numpy())

    # ---- projection + bbox2d crop of nvdiffrast_render (Utils.py:159-181): where does a camera point land in the crop?
    H, W = 480, 640
    proj = ref["projection_matrix_from_intrinsics"](K, height=H, width=W, znear=0.001, zfar=100)
    pts_cam = np.stack([rng.uniform(-0.2, 0.2, 64), rng.uniform(-0.15, 0.15, 64), rng.uniform(0.3, 1.0, 64)], 1)
    glcam_in_cvcam = np.diag([1.0, -1.0, -1.0, 1.0])
    pts_gl = (glcam_in_cvcam @ np.concatenate([pts_cam, np.ones((64, 1))], 1).T).T
    clip = (proj @ pts_gl.T).T  # Utils.py:161-163: pos_clip = projection_mat @ pos_homo
    bbox = np.array([200.0, 150.0, 359.0, 309.0])  # umin, vmin, umax, vmax  (predict_pose_refine.py:44-45 style window)
    # Utils.py:166-181 (bbox2d branch), restated line by line on the reference's own clip coordinates:
    l, t_, r, b = bbox[0], H - bbox[1], bbox[2], H - bbox[3]
    tf_ndc = np.eye(4)
    tf_ndc[0, 0] = W / (r - l)
    tf_ndc[1, 1] = H / (t_ - b)
    tf_ndc[3, 0] = (W - r - l) / (r - l)
    tf_ndc[3, 1] = (H - t_ - b) / (t_ - b)
    clip_c = clip @ tf_ndc  # pos_clip @ tf (row-vector convention of the reference)
    ndc = clip_c[:, :3] / clip_c[:, 3:4]
    S = 160
    px = (ndc[:, 0] * 0.5 + 0.5) * S  # window x (pixel edge units), nvdiffrast R2: centres at +0.5
    py = (ndc[:, 1] * 0.5 + 0.5) * S  # window y, row 0 at the bottom; the reference flips vertically afterwards
    out.update(pj_K=K, pj_proj=proj, pj_pts_cam=pts_cam, pj_bbox=bbox, pj_px=px, pj_py_from_bottom=py)

    path = os.path.join(ROOT, "tests", "golden", "geometry_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()
