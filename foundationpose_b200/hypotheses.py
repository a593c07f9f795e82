"""Start-pose generation (init-time, host side): the 42 x 6 = 252 rotation grid and the translation guess.

  sample_views_icosphere   Utils.py:483-507
  make_rotation_grid       estimater.py:106-124
  cluster_poses            mycpp/src/app/pybind_api.cpp:24-68 + mycpp/src/Utils.cpp:21-26 (greedy de-duplication
                           by geodesic rotation distance under the symmetry group; init-only, so it stays
                           a host loop here)
  guess_translation        estimater.py:137-156
"""
import numpy as np

from .synth import icosphere


def sample_views_icosphere(n_views, subdivisions=None, radius=1.0):
    if subdivisions is None:
        subdivisions = 1
        while True:
            verts, _ = icosphere(subdivisions)
            if len(verts) >= n_views:
                break
            subdivisions += 1
    else:
        verts, _ = icosphere(subdivisions)
    verts = verts * radius
    cam_in_obs = np.tile(np.eye(4)[None], (len(verts), 1, 1))
    cam_in_obs[:, :3, 3] = verts
    up = np.array([0.0, 0.0, 1.0])
    z_axis = -cam_in_obs[:, :3, 3]
    z_axis /= np.linalg.norm(z_axis, axis=-1).reshape(-1, 1)
    x_axis = np.cross(up.reshape(1, 3), z_axis)
    invalid = (x_axis == 0).all(axis=-1)
    x_axis[invalid] = [1, 0, 0]
    x_axis /= np.linalg.norm(x_axis, axis=-1).reshape(-1, 1)
    y_axis = np.cross(z_axis, x_axis)
    y_axis /= np.linalg.norm(y_axis, axis=-1).reshape(-1, 1)
    cam_in_obs[:, :3, 0] = x_axis
    cam_in_obs[:, :3, 1] = y_axis
    cam_in_obs[:, :3, 2] = z_axis
    return cam_in_obs


def _rot_z(a):
    c, s = np.cos(a), np.sin(a)
    m = np.eye(4)
    m[0, 0], m[0, 1], m[1, 0], m[1, 1] = c, -s, s, c
    return m


def rotation_geodesic_distance(R1, R2):
    cos = np.float32((np.trace(R1 @ R2.T) - 1) / 2.0)
    return float(np.arccos(np.clip(cos, -1.0, 1.0)))


def cluster_poses(angle_diff_deg, dist_diff, poses_in, symmetry_tfs):
    poses_in = np.asarray(poses_in, dtype=np.float32)
    symmetry_tfs = np.asarray(symmetry_tfs, dtype=np.float32).reshape(-1, 4, 4)
    out = [poses_in[0]]
    thres = np.float32(angle_diff_deg / 180.0 * np.pi)
    for i in range(1, len(poses_in)):
        cur = poses_in[i]
        isnew = True
        for cluster in out:
            if np.linalg.norm(cluster[:3, 3] - cur[:3, 3]) >= dist_diff:
                continue
            for tf in symmetry_tfs:
                tmp = cur @ tf
                if rotation_geodesic_distance(tmp[:3, :3], cluster[:3, :3]) < thres:
                    isnew = False
                    break
            if not isnew:
                break
        if isnew:
            out.append(cur)
    return np.asarray(out)


def make_rotation_grid(min_n_views=40, inplane_step=60, symmetry_tfs=None):
    cam_in_obs = sample_views_icosphere(n_views=min_n_views)
    rot_grid = []
    for i in range(len(cam_in_obs)):
        for inplane_rot in np.deg2rad(np.arange(0, 360, inplane_step)):
            cam_in_ob = cam_in_obs[i] @ _rot_z(inplane_rot)
            rot_grid.append(np.linalg.inv(cam_in_ob))
    rot_grid = np.asarray(rot_grid)
    if symmetry_tfs is None:
        symmetry_tfs = np.eye(4)[None]
    return cluster_poses(30, 99999, rot_grid, symmetry_tfs).astype(np.float32)


def guess_translation(depth, mask, K):
    vs, us = np.where(mask > 0)
    if len(us) == 0:
        return np.zeros(3)
    uc = (us.min() + us.max()) / 2.0
    vc = (vs.min() + vs.max()) / 2.0
    valid = mask.astype(bool) & (depth >= 0.001)
    if not valid.any():
        return np.zeros(3)
    zc = np.median(depth[valid])
    center = (np.linalg.inv(K) @ np.asarray([uc, vc, 1]).reshape(3, 1)) * zc
    return center.reshape(3)
