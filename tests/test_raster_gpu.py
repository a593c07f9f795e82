"""GPU parity of the tiled rasteriser's special paths (csrc/fp_crop.cu, csrc/fp_meshlet.cu) against oracle/raster.py:
open meshes (both sides rendered, no culling), triangles crossing the near plane (homogeneous path), vertex-coloured
meshes, back-face culling on/off equivalence for a closed mesh, and a mesh swap between graph replays."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(mesh, pose, poses, cull_env=None):
    from foundationpose_b200 import synth
    from foundationpose_b200.engine import Engine
    from oracle import pipeline

    rgb, depth, mask = synth.make_scene(synth.make_texture(0, 256), pose)
    d = synth.mesh_diameter(mesh.vertices)
    mt = pipeline.mesh_tensors(mesh)
    if cull_env is not None:
        os.environ["FPOSE_NO_CULL"] = cull_env
    try:
        e = Engine()
    finally:
        os.environ.pop("FPOSE_NO_CULL", None)
    e.set_mesh(mt["pos"], mt["normals"], mt["faces"], d, uv=mt.get("uv"), tex=mt.get("tex"), vertex_colors=mt.get("vcolor"))
    e.set_frame(rgb, depth, synth.DEFAULT_K, filter_depth=False)
    return e, mt, rgb, depth, d


def _check_A(e, mt, rgb, depth, d, poses, min_cov=0.999):
    from foundationpose_b200 import synth
    from oracle import geometry, pipeline

    _, dbg, _ = e.make_crops(poses, mode=0, want_dbg=True)
    xyz = geometry.depth2xyzmap(depth, synth.DEFAULT_K)
    A, B, _ = pipeline.make_crops(poses, mt, rgb, depth, xyz, synth.DEFAULT_K, d, 0)
    gA = dbg[:, 0].permute(0, 3, 1, 2).cpu()
    cov_g = gA[:, 3:].abs().sum(1) > 0
    cov_o = A[:, 3:].abs().sum(1) > 0
    agree = (cov_g == cov_o).float().mean().item()
    assert agree >= min_cov, f"raster coverage agreement {agree}"
    both = (cov_g & cov_o)[:, None].expand(-1, 3, -1, -1)
    assert both.any()
    assert ((gA[:, 3:] - A[:, 3:]).abs()[both] <= 2e-4).float().mean().item() >= 0.999
    assert ((gA[:, :3] - A[:, :3]).abs()[both] <= 2e-3).float().mean().item() >= 0.998
    return cov_o.float().mean().item()


def _base_pose():
    from foundationpose_b200 import synth

    pose = np.eye(4)
    pose[:3, :3] = synth.random_rotation(3)
    pose[:3, 3] = [0.01, 0.0, 0.6]
    return pose


def test_closed_mesh_is_detected_and_culled():
    from foundationpose_b200 import synth

    mesh = synth.make_mesh(3)
    pose = _base_pose()
    e, *_ = _setup(mesh, pose, None)
    info = e.mesh_info()
    assert info["closed"] and info["front_sign"] == -1 and info["F"] == 1280
    assert info["meshlets"] >= 1280 // 64


def test_open_mesh_renders_back_faces():
    """A bowl (sphere with one cap removed) seen through the opening: the visible surface is made of BACK faces, which
    nvdiffrast renders (no culling); the kernel must not cull them."""
    from foundationpose_b200 import synth

    mesh = synth.make_mesh(3)
    keep = mesh.vertices[mesh.faces].mean(1)[:, 2] < 0.04  # drop the +z cap
    mesh.faces = mesh.faces[keep]
    pose = np.eye(4)
    pose[:3, :3] = np.diag([1.0, -1.0, -1.0])  # object +z towards the camera: we look into the bowl
    pose[:3, 3] = [0.0, 0.0, 0.55]
    poses = np.stack([pose, pose]).astype(np.float32)
    poses[1, :3, :3] = poses[1, :3, :3] @ synth.random_rotation(5)[:3, :3]
    e, mt, rgb, depth, d = _setup(mesh, pose, poses)
    info = e.mesh_info()
    assert not info["closed"] and info["front_sign"] == 0
    _check_A(e, mt, rgb, depth, d, poses)


def test_near_plane_crossing_triangles():
    """Camera inside / touching the object: triangles with vertices on both sides of z = 1 mm are clipped per pixel."""
    from foundationpose_b200 import synth

    mesh = synth.make_mesh(2)
    pose = np.eye(4)
    pose[:3, :3] = synth.random_rotation(9)
    pose[:3, 3] = [0.004, -0.003, 0.03]  # the 0.05-0.095 m ellipsoid contains the camera centre
    poses = np.stack([pose, pose]).astype(np.float32)
    poses[1, :3, 3] = [0.0, 0.0, 0.052]
    e, mt, rgb, depth, d = _setup(mesh, pose, poses)
    st = e.crop_stats(poses)
    assert st["near_plane_triangles"] > 0, st
    cov = _check_A(e, mt, rgb, depth, d, poses, min_cov=0.995)
    assert cov > 0.5


def test_vertex_colour_mesh():
    from foundationpose_b200 import synth

    q, f = synth.icosphere(3)
    rng = np.random.default_rng(4)
    mesh = synth.SimpleMesh(q * synth.RADII, f, q, vertex_colors=rng.integers(0, 256, size=(len(q), 4)).astype(np.uint8))
    pose = _base_pose()
    poses = np.stack([pose]).astype(np.float32)
    e, mt, rgb, depth, d = _setup(mesh, pose, poses)
    _check_A(e, mt, rgb, depth, d, poses)


def test_culling_does_not_change_the_crops():
    """Closed mesh: crops with back-face + cone culling == crops with every triangle rasterised (FPOSE_NO_CULL=1)."""
    from foundationpose_b200 import synth

    mesh = synth.make_mesh(4)
    pose = _base_pose()
    poses = np.stack([pose] * 6).astype(np.float32)
    for i in range(1, 6):
        poses[i, :3, :3] = synth.random_rotation(40 + i)
    poses[5, :3, 3] = [0.2, 0.12, 0.45]
    e1, mt, rgb, depth, d = _setup(mesh, pose, poses)
    e2, *_ = _setup(mesh, pose, poses, cull_env="1")
    assert e1.mesh_info()["front_sign"] == -1 and e2.mesh_info()["front_sign"] == 0
    s1, s2 = e1.crop_stats(poses), e2.crop_stats(poses)
    assert s1["meshlet_visits"] < 0.75 * s2["meshlet_visits"], (s1, s2)
    c1, _, _ = e1.make_crops(poses, mode=0)
    c2, _, _ = e2.make_crops(poses, mode=0)
    diff = (c1.float() - c2.float()).abs().amax(dim=(1, 2, 3, 4))
    n_px = ((c1.float() - c2.float()).abs().amax(dim=-1) > 0).sum().item()
    # a back face can only win where snapping flips an edge-on sliver at the silhouette
    assert n_px <= 4 * len(poses), f"{n_px} pixels differ between culled and unculled rendering (max {diff.max():.3g})"


def test_tile_size_does_not_change_the_crops():
    """16 / 32 / 80-pixel tiles (track_one, sharded, full batch) produce bit-identical crop buffers."""
    from foundationpose_b200 import synth

    mesh = synth.make_mesh(4)
    pose = _base_pose()
    poses = np.stack([pose] * 5).astype(np.float32)
    for i in range(1, 5):
        poses[i, :3, :3] = synth.random_rotation(60 + i)
    poses[4, :3, 3] = [0.2, 0.12, 0.45]
    e, mt, rgb, depth, d = _setup(mesh, pose, poses)
    ref = None
    for tile in (16, 32, 80):
        e.set_crop_tile(tile)
        for mode in (0, 1):
            c, _, _ = e.make_crops(poses, mode=mode)
            if tile == 16:
                ref = ref or {}
                ref[mode] = c.clone()
            else:
                assert torch.equal(c, ref[mode]), f"tile {tile}, mode {mode}: crops differ from the 16-pixel tiling"
    e.set_crop_tile(0)


def test_mesh_swap_between_graph_replays():
    """a larger mesh after a (kind, N, iters) graph exists must not allocate during capture."""
    from foundationpose_b200 import synth
    from foundationpose_b200.weights import random_state_dict
    from oracle import pipeline

    small, big = synth.make_mesh(2), synth.make_mesh(4)
    pose = _base_pose()
    poses = np.stack([pose, pose]).astype(np.float32)
    poses[1, :3, 3] += [0.01, 0.0, 0.01]
    e, mt, rgb, depth, d = _setup(small, pose, poses)
    e.load_network("refine", random_state_dict("refine", 0))
    for _ in range(3):  # eager, capture, replay
        a, _, _ = e.refine(poses, 2)
    mtb = pipeline.mesh_tensors(big)
    e.set_mesh(mtb["pos"], mtb["normals"], mtb["faces"], d, uv=mtb["uv"], tex=mtb["tex"])
    for _ in range(3):  # re-capture with the larger mesh, replay
        b, _, _ = e.refine(poses, 2)
    torch.cuda.synchronize()
    assert torch.isfinite(b).all()
    e.set_mesh(mt["pos"], mt["normals"], mt["faces"], d, uv=mt["uv"], tex=mt["tex"])
    c, _, _ = e.refine(poses, 2)
    assert torch.equal(a, c), "same mesh, same poses: the result must be reproduced after the swap"
