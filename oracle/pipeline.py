"""CPU restatement of the predictors' data path and of the estimator's flow (TEST INFRASTRUCTURE, see
oracle/__init__.py): make_crop_data_batch for the refiner (predict_pose_refine.py:25-89) and the scorer
(predict_score.py:56-114), PoseRefinePredictor.predict's iteration (predict_pose_refine.py:182-239),
ScorePredictor.predict (predict_score.py:160-214), the ranking of estimater.py:226-235, register (estimater.py:159-240)
and track_one (:250-268).

Pinning: everything in this file EXCEPT the two third-party primitives it calls (raster.render_crop for nvdiffrast,
geometry.warp_perspective for kornia — PARITY UNPINNED, those packages are absent) is held to the reference's own
unmodified sources, executed by tools/make_golden_flow.py (tests/test_flow_golden_cpu.py).
"""
import numpy as np
import torch

from . import geometry, nets, raster

S = 160


def mesh_tensors(mesh):
    """Utils.py:104-130 make_mesh_tensors on a trimesh-like object (uv v-flip at :117)."""
    out = dict(pos=np.asarray(mesh.vertices, dtype=np.float32), faces=np.asarray(mesh.faces, dtype=np.int64),
               normals=np.asarray(mesh.vertex_normals, dtype=np.float32), tex=None)
    if getattr(mesh.visual, "uv", None) is not None and getattr(mesh.visual, "image", None) is not None:
        uv = np.asarray(mesh.visual.uv, dtype=np.float32).copy()
        uv[:, 1] = 1 - uv[:, 1]
        out["uv"] = uv
        out["tex"] = np.asarray(mesh.visual.image)[..., :3].astype(np.uint8)
    else:
        out["vcolor"] = np.asarray(mesh.visual.vertex_colors, dtype=np.float32)[..., :3] / 255.0
    return out


def make_crops(poses, mt, rgb, depth, xyz_map, K, mesh_diameter, mode, crop_ratio=1.2):
    """Returns A, B (N,6,160,160) float32 torch tensors (network inputs) and the crop windows.
    mode 0 = refiner (xyz_map warped nearest, tau 0.001); mode 1 = scorer (depth round trip, tau 0.1)."""
    poses = np.asarray(poses, dtype=np.float32)
    N = len(poses)
    H, W = depth.shape
    win, tf = geometry.crop_window(poses, K, mesh_diameter, crop_ratio, S)
    umin, vmin, umax, vmax = geometry.render_window(win, S)
    rgbA = np.zeros((N, S, S, 3), dtype=np.float32)
    xyzA = np.zeros((N, S, S, 3), dtype=np.float32)
    for n in range(N):
        rgbA[n], xyzA[n], _ = raster.render_crop(poses[n], mt, K, (umin[n], vmin[n], umax[n], vmax[n]))
    tf_t = torch.from_numpy(tf)
    rgb_t = torch.as_tensor(np.asarray(rgb), dtype=torch.float32).permute(2, 0, 1)[None].expand(N, -1, -1, -1)
    rgbB = geometry.warp_perspective(rgb_t, tf_t, (S, S), "bilinear")
    t = torch.from_numpy(poses[:, :3, 3].copy())
    if mode == 0:
        xyz_t = torch.as_tensor(xyz_map, dtype=torch.float32).permute(2, 0, 1)[None].expand(N, -1, -1, -1)
        xyzB = geometry.warp_perspective(xyz_t, tf_t, (S, S), "nearest")
        tau = 0.001
    else:
        # predict_score.py:90 + h5_dataset.py:158-161: depth crop -> full res -> xyz -> crop (all nearest)
        d_t = torch.as_tensor(depth, dtype=torch.float32)[None, None].expand(N, -1, -1, -1)
        depthB = geometry.warp_perspective(d_t, tf_t, (S, S), "nearest")
        # the inverse warp in closed form: its border pixels are rounding ties (see geometry.unwarp_nearest)
        depthB_ori = geometry.unwarp_nearest(depthB, win, (H, W))
        Kf = np.asarray(K, dtype=np.float32)
        xyz_ori = torch.stack([torch.from_numpy(geometry.depth2xyzmap(depthB_ori[n, 0].numpy(), Kf)) for n in range(N)]).permute(0, 3, 1, 2)
        xyzB = geometry.warp_perspective(xyz_ori, tf_t, (S, S), "nearest")
        tau = 0.1
    A_rgb = torch.from_numpy(rgbA).permute(0, 3, 1, 2) * 255 / 255.0  # predict_pose_refine.py:54, h5_dataset.py:215
    B_rgb = rgbB / 255.0
    A_xyz = geometry.normalise_xyz(torch.from_numpy(xyzA).permute(0, 3, 1, 2), t, mesh_diameter, tau)
    B_xyz = geometry.normalise_xyz(xyzB, t, mesh_diameter, tau)
    A = torch.cat([A_rgb, A_xyz], 1).float()
    B = torch.cat([B_rgb, B_xyz], 1).float()
    return A, B, win


def refine(sd, poses, mt, rgb, depth, K, mesh_diameter, iterations, rot_normalizer=0.3490658503988659, xyz_map=None):
    """PoseRefinePredictor.predict (predict_pose_refine.py:149-239). Returns poses, last trans/rot deltas."""
    if xyz_map is None:
        xyz_map = geometry.depth2xyzmap(depth, K)
    poses = torch.as_tensor(np.asarray(poses), dtype=torch.float32).clone()
    td = rd = None
    for _ in range(iterations):
        A, B, _ = make_crops(poses.numpy(), mt, rgb, depth, xyz_map, K, mesh_diameter, 0)
        out = nets.refine_forward(sd, A, B)
        poses, td, rd = geometry.pose_update(poses, out["trans"], out["rot"], mesh_diameter, rot_normalizer)
    return poses, td, rd


def score(sd, poses, mt, rgb, depth, K, mesh_diameter):
    """ScorePredictor.predict (predict_score.py:160-214): returns scores (+100) and the best index
    (estimater.py:226: argsort(descending)[0])."""
    A, B, _ = make_crops(np.asarray(poses), mt, rgb, depth, None, K, mesh_diameter, 1)
    logits = nets.score_forward(sd, A, B, L=len(A)).reshape(-1)
    scores = logits + 100
    return scores, int(scores.argmax())


def _to_centered(model_center):
    tf = torch.eye(4, dtype=torch.float32)
    tf[:3, 3] = -torch.as_tensor(np.asarray(model_center), dtype=torch.float32)
    return tf


def register(sd_r, sd_s, rot_grid, mt, rgb, depth, mask, K, mesh_diameter, model_center, iterations=5):
    """FoundationPose.register (estimater.py:159-240): depth filters, the < 4 valid pixels early-out, start poses =
    rotation grid + guessed translation, K refine iterations, scoring, argsort(descending), best pose in the frame of the
    ORIGINAL (un-centred) mesh.  Returns a dict: pose (4,4) float64 numpy; and unless `early`: ids, poses / scores in
    ranked order, pose_last, last_trans / last_rot.  Held to the reference's own method sources by
    tests/test_flow_golden_cpu.py (tools/make_golden_flow.py)."""
    depth = geometry.bilateral_filter_depth(geometry.erode_depth(np.asarray(depth)))
    center = geometry.guess_translation(depth, mask, K)
    if ((depth >= 0.001) & (np.asarray(mask) > 0)).sum() < 4:
        pose = np.eye(4)
        pose[:3, 3] = center
        return dict(pose=pose, early=True)
    poses = torch.as_tensor(np.asarray(rot_grid), dtype=torch.float32).clone()
    poses[:, :3, 3] = torch.as_tensor(center.reshape(1, 3), dtype=torch.float32)
    xyz_map = geometry.depth2xyzmap(depth, K)
    poses, lt, lr = refine(sd_r, poses, mt, rgb, depth, K, mesh_diameter, iterations, xyz_map=xyz_map)
    scores, _ = score(sd_s, poses.numpy(), mt, rgb, depth, K, mesh_diameter)
    ids = scores.argsort(descending=True)
    poses, scores = poses[ids], scores[ids]
    best = poses[0] @ _to_centered(model_center)
    return dict(pose=best.numpy(), early=False, ids=ids, poses=poses, scores=scores, pose_last=poses[0], last_trans=lt, last_rot=lr)


def track_one(sd_r, pose_last, mt, rgb, depth, K, mesh_diameter, model_center, iterations=2):
    """FoundationPose.track_one (estimater.py:250-268): depth filters, xyz map with zfar = inf, `iterations` refiner passes
    from the previous pose (centred-mesh frame).  Returns (pose (4,4) numpy in the original mesh frame, new pose_last,
    last_trans)."""
    depth = geometry.bilateral_filter_depth(geometry.erode_depth(np.asarray(depth, dtype=np.float32)))
    xyz_map = geometry.depth2xyzmap(depth, K, zfar=np.inf)
    pose, lt, _ = refine(sd_r, torch.as_tensor(np.asarray(pose_last), dtype=torch.float32).reshape(1, 4, 4), mt, rgb, depth, K, mesh_diameter,
                         iterations, xyz_map=xyz_map)
    return (pose @ _to_centered(model_center)).numpy().reshape(4, 4), pose, lt
