"""CPU: self-consistency of the oracle's geometry restatement (it has no reference fixtures to pin
against — see oracle/__init__.py — so these are the properties the cited reference code implies)."""
import numpy as np
import torch

from foundationpose_b200 import synth
from oracle import geometry, pipeline, raster


def test_crop_window_is_square_integer_and_centred():
    poses = np.tile(np.eye(4, dtype=np.float32)[None], (3, 1, 1))
    poses[:, :3, 3] = [[0, 0, 0.6], [0.05, -0.02, 0.8], [-0.1, 0.1, 0.5]]
    win, tf = geometry.crop_window(poses, synth.DEFAULT_K, 0.19)
    for k in ("left", "right", "top", "bottom"):
        assert np.array_equal(win[k], np.round(win[k]))
    # radius = fx * r / z  (box_3d): width ~ 2 * 615 * 0.114 / z
    assert abs((win["right"][0] - win["left"][0]) - 2 * 615 * 0.19 * 1.2 / 2 / 0.6) <= 1
    assert np.allclose(tf[:, 0, 2], -win["left"] * win["sx"])


def test_kornia_warp_identity_quirk():
    """with align_corners=False kornia normalises with (size-1) but grid_sample
    un-normalises with size, so destination pixel j reads source x = j*W/(W-1) - 0.5.  Even the identity
    homography is therefore NOT the identity: the last row/column samples x = W - 0.5 -> rounds to W
    (out of bounds) -> 0.  The engine reproduces this, it does not 'fix' it."""
    g = torch.Generator().manual_seed(0)
    src = torch.rand(1, 3, 48, 64, generator=g)
    out = geometry.warp_perspective(src, torch.eye(3)[None], (48, 64), "nearest")
    assert torch.equal(out[..., :47, :63], src[..., :47, :63])
    assert out[..., 47, :].abs().max() == 0 and out[..., :, 63].abs().max() == 0
    # closed form of the coordinate chain
    j = torch.arange(64, dtype=torch.float32)
    assert torch.equal(torch.round(j * 64 / 63 - 0.5)[:63], j[:63])


def test_so3_exp_map_is_a_rotation_and_small_angle_safe():
    v = torch.tensor([[0.0, 0.0, 0.0], [0.1, -0.2, 0.3], [1e-5, 0, 0]])
    R = geometry.so3_exp_map(v)
    assert torch.allclose(R @ R.transpose(1, 2), torch.eye(3)[None].expand(3, -1, -1), atol=1e-5)
    assert torch.allclose(R[0], torch.eye(3), atol=1e-7)


def test_raster_is_watertight_and_consistent_with_analytic_depth():
    mesh = synth.make_mesh(3)
    mt = pipeline.mesh_tensors(mesh)
    pose = np.eye(4, dtype=np.float32)
    pose[:3, :3] = synth.random_rotation(1)
    pose[:3, 3] = [0, 0, 0.6]
    d = synth.mesh_diameter(mesh.vertices)
    win, _ = geometry.crop_window(pose[None], synth.DEFAULT_K, d)
    window = [w[0] for w in geometry.render_window(win)]
    rgb, xyz, tri = raster.render_crop(pose, mt, synth.DEFAULT_K, window)
    cov = tri >= 0
    assert 0.05 < cov.mean() < 0.6
    # watertight: the silhouette of a closed convex mesh has no holes in any row
    for r in range(160):
        idx = np.where(cov[r])[0]
        if len(idx):
            assert cov[r, idx[0]: idx[-1] + 1].all()
    # depth agrees with the analytic ellipsoid at the crop's sampling positions
    _, depth, hit = synth.make_scene(mesh.visual.image, pose.astype(np.float64), depth_noise=0.0)
    z = xyz[..., 2]
    us = window[0] + (np.arange(160) + 0.5) * (window[2] - window[0]) / 160
    vs = window[1] + (np.arange(160) + 0.5) * (window[3] - window[1]) / 160
    ui, vi = np.clip(np.round(us - 0.5).astype(int), 0, 639), np.clip(np.round(vs - 0.5).astype(int), 0, 479)
    ref = depth[vi][:, ui]
    inner = cov & hit[vi][:, ui]
    err = np.abs(z[inner] - ref[inner])  # grazing-angle pixels move by mm per half pixel; faceting adds ~0.5 mm
    assert np.median(err) < 1.5e-3 and np.percentile(err, 95) < 6e-3
    assert rgb.min() >= 0 and rgb.max() <= 1 and rgb[~cov].max() == 0


def test_depth_filters_match_naive_loops():
    rng = np.random.default_rng(0)
    d = (0.5 + 0.01 * rng.standard_normal((12, 14))).astype(np.float32)
    d[rng.random(d.shape) < 0.2] = 0
    d[3, 4] = 150.0

    def naive_erode(depth):
        H, W = depth.shape
        out = np.zeros_like(depth)
        for h in range(H):
            for w in range(W):
                d0 = depth[h, w]
                bad = tot = 0.0
                for u in range(w - 2, w + 3):
                    if u < 0 or u >= W:
                        continue
                    for v in range(h - 2, h + 3):
                        if v < 0 or v >= H:
                            continue
                        c = depth[v, u]
                        tot += 1
                        if c < 0.001 or c >= 100 or abs(c - d0) > np.float32(0.001):
                            bad += 1
                out[h, w] = 0 if bad / tot > 0.8 else d0
        return out

    assert np.array_equal(geometry.erode_depth(d), naive_erode(d))
    b = geometry.bilateral_filter_depth(d)
    assert b.shape == d.shape and np.isfinite(b).all()
    assert (b[d == 0] >= 0).all()
