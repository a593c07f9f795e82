"""CPU restatement of the geometry around the networks (TEST INFRASTRUCTURE; PINNED against the reference's own
function bodies — crop_window, depth2xyzmap, pose composition, normalise_xyz, erode / bilateral depth filters, see
oracle/__init__.py — except so3_exp_map (pytorch3d) and warp_perspective (kornia), which are PARITY UNPINNED).

Each function cites the reference lines it follows.  fp32 throughout, like the reference's CUDA
tensors; operation order is kept explicit where a rounding decides an integer (crop window edges).
"""
import numpy as np
import torch
import torch.nn.functional as F

f32 = np.float32


def crop_window(poses, K, mesh_diameter, crop_ratio=1.2, out_size=160):
    """Utils.py:577-621 compute_crop_window_tf_batch(method='box_3d') + :584-598.
    poses (N,4,4) -> dict(left, top, right, bottom, sx, sy) as float32 arrays, and tf (N,3,3)."""
    poses = np.asarray(poses, dtype=f32)
    K = np.asarray(K, dtype=f32)
    r = f32(float(mesh_diameter) * float(crop_ratio) / 2.0)
    offs = np.array([[0, 0, 0], [r, 0, 0], [-r, 0, 0], [0, r, 0], [0, -r, 0]], dtype=f32)
    pts = poses[:, None, :3, 3] + offs[None]  # (N,5,3)
    # projected = (K @ pts.T).T with K01 = K10 = K20 = K21 = 0, K22 = 1 (Utils.py:611)
    x = K[0, 0] * pts[..., 0] + K[0, 2] * pts[..., 2]
    y = K[1, 1] * pts[..., 1] + K[1, 2] * pts[..., 2]
    u = x / pts[..., 2]
    v = y / pts[..., 2]
    radius = np.maximum(np.abs(u - u[:, :1]).max(1), np.abs(v - v[:, :1]).max(1)).astype(f32)
    left = np.round(u[:, 0] - radius).astype(f32)  # torch.round: half to even, like np.round
    right = np.round(u[:, 0] + radius).astype(f32)
    top = np.round(v[:, 0] - radius).astype(f32)
    bottom = np.round(v[:, 0] + radius).astype(f32)
    # `out_size[0] / (right - left)` with a Python int on the left is Tensor.__rtruediv__ = reciprocal() * 160:
    # two fp32 roundings (pinned bit-exactly by tests/golden/geometry_golden.npz)
    sx = ((f32(1) / (right - left)).astype(f32) * f32(out_size)).astype(f32)
    sy = ((f32(1) / (bottom - top)).astype(f32) * f32(out_size)).astype(f32)
    tf = np.zeros((len(poses), 3, 3), dtype=f32)
    tf[:, 0, 0] = sx
    tf[:, 1, 1] = sy
    tf[:, 0, 2] = sx * (-left)
    tf[:, 1, 2] = sy * (-top)
    tf[:, 2, 2] = 1
    return dict(left=left, top=top, right=right, bottom=bottom, sx=sx, sy=sy), tf


def render_window(win, out_size=160):
    """predict_pose_refine.py:44-45: bbox2d_ori = tf_to_crop^-1 applied to (0,0), (S-1,S-1)."""
    s1 = f32(out_size - 1)
    umin, vmin = win["left"], win["top"]
    umax = (win["left"] + s1 / win["sx"]).astype(f32)
    vmax = (win["top"] + s1 / win["sy"]).astype(f32)
    return umin, vmin, umax, vmax


def depth2xyzmap(depth, K, zfar=np.inf):
    """Utils.py:399-417 / :420-438 (invalid: z < 0.001 or z > zfar -> 0).  depth (H,W) -> (H,W,3) float32."""
    depth = np.asarray(depth, dtype=f32)
    K = np.asarray(K, dtype=f32)
    H, W = depth.shape
    vs, us = np.meshgrid(np.arange(H, dtype=f32), np.arange(W, dtype=f32), indexing="ij")
    xs = (us - K[0, 2]) * depth / K[0, 0]
    ys = (vs - K[1, 2]) * depth / K[1, 1]
    xyz = np.stack([xs, ys, depth], -1).astype(f32)
    xyz[(depth < 0.001) | (depth > zfar)] = 0
    return xyz


def warp_perspective(src, M, dsize, mode):
    """kornia 0.7.2 warp_perspective(src, M, dsize, mode, align_corners=False), padding zeros, restated
    as its op sequence: normalise the homography with the (size-1) convention,
    invert, transform a linspace(-1,1) meshgrid, F.grid_sample(align_corners=False).
    src (B,C,H,W) torch float32, M (B,3,3)."""
    B, _, H, W = src.shape
    h_out, w_out = dsize

    def normal_transform_pixel(h, w):
        t = torch.tensor([[1.0, 0.0, -1.0], [0.0, 1.0, -1.0], [0.0, 0.0, 1.0]], dtype=src.dtype)
        t[0, 0] = t[0, 0] * 2.0 / max(w - 1.0, 1e-14)
        t[1, 1] = t[1, 1] * 2.0 / max(h - 1.0, 1e-14)
        return t[None]

    src_norm_trans_src_pix = normal_transform_pixel(H, W)
    src_pix_trans_src_norm = torch.inverse(src_norm_trans_src_pix)
    dst_norm_trans_dst_pix = normal_transform_pixel(h_out, w_out)
    dst_norm_trans_src_norm = dst_norm_trans_dst_pix @ (M @ src_pix_trans_src_norm)
    src_norm_trans_dst_norm = torch.inverse(dst_norm_trans_src_norm)
    xs = torch.linspace(-1, 1, w_out, dtype=src.dtype)
    ys = torch.linspace(-1, 1, h_out, dtype=src.dtype)
    gy, gx = torch.meshgrid(ys, xs, indexing="ij")
    grid = torch.stack([gx, gy], -1)[None].repeat(B, 1, 1, 1)  # (B,h,w,2)
    pts = torch.cat([grid, torch.ones_like(grid[..., :1])], -1)
    tp = (src_norm_trans_dst_norm[:, None, None] @ pts[..., None])[..., 0]
    grid = tp[..., :2] / tp[..., 2:3]
    return F.grid_sample(src, grid, mode=mode, padding_mode="zeros", align_corners=False)


def _kornia_src_coord(x, size):
    """Closed form of the coordinate chain above for one axis, fp32 in the kernel's operation
    order: (size-1)-normalisation, then grid_sample's align_corners=False un-normalisation."""
    x = np.asarray(x, dtype=f32)
    xn = ((f32(2) * x).astype(f32) / f32(size - 1)).astype(f32) - f32(1)
    return ((((xn + f32(1)).astype(f32) * f32(size)).astype(f32) - f32(1)).astype(f32) * f32(0.5)).astype(f32)


def unwarp_nearest(crop, win, full_hw):
    """warp_perspective(crop, tf_to_crop^-1, (H, W), 'nearest') of h5_dataset.py:158 for the axis-aligned crop
    transform, in CLOSED FORM: full-resolution pixel (v, u) reads crop pixel (rint(k(sy v - top sy)), rint(k(sx u - left sx))),
    zeros outside.  Why not the op sequence of `warp_perspective` here: the crop window has integer edges, so the first
    crop row / column maps back to EXACTLY -0.5, a rounding tie; through kornia's chain the outcome is decided by the
    last bit of a 3x3 LU inverse (this CPU: the column survives for 151 of 228 random windows and vanishes for 77;
    cuSOLVER rounds differently again), i.e. the reference's own result is implementation-defined there.  The closed
    form resolves the tie the way exact arithmetic does (round-half-even: -0.5 -> 0, valid) and is what csrc/fp_crop.cu
    evaluates.  crop (N,1,S,S) torch; win = dict(left, top, sx, sy) of float32 arrays."""
    N, _, S, _ = crop.shape
    H, W = full_hw
    out = torch.zeros(N, 1, H, W, dtype=crop.dtype)
    us = np.arange(W, dtype=f32)
    vs = np.arange(H, dtype=f32)
    for n in range(N):
        sx, sy, left, top = f32(win["sx"][n]), f32(win["sy"][n]), f32(win["left"][n]), f32(win["top"][n])
        jc = np.rint(_kornia_src_coord(((sx * us).astype(f32) + ((-left) * sx).astype(f32)).astype(f32), S)).astype(np.int64)
        ic = np.rint(_kornia_src_coord(((sy * vs).astype(f32) + ((-top) * sy).astype(f32)).astype(f32), S)).astype(np.int64)
        okj = (jc >= 0) & (jc < S)
        oki = (ic >= 0) & (ic < S)
        sub = crop[n, 0][torch.from_numpy(ic[oki])][:, torch.from_numpy(jc[okj])]
        o = out[n, 0]
        o[np.ix_(np.nonzero(oki)[0], np.nonzero(okj)[0])] = sub
    return out


def normalise_xyz(xyz, t, mesh_diameter, tau):
    """h5_dataset.py:93-99 (refiner, tau = 0.001) / :151-156 (scorer, tau = 0.1).
    xyz (B,3,H,W) torch, t (B,3)."""
    invalid = xyz[:, 2:3] < tau
    xyz = xyz - t.reshape(-1, 3, 1, 1)
    xyz = xyz * (1.0 / (torch.tensor(mesh_diameter, dtype=torch.float32) / 2))
    invalid = invalid.expand(-1, 3, -1, -1) | (xyz.abs() >= 2)
    xyz = xyz.clone()
    xyz[invalid] = 0
    return xyz


def so3_exp_map(v, eps=1e-4):
    """pytorch3d.transforms.so3_exp_map. v (N,3) -> (N,3,3)."""
    nrms = (v * v).sum(1)
    th = torch.clamp(nrms, eps).sqrt()
    ith = 1.0 / th
    fac1 = ith * th.sin()
    fac2 = ith * ith * (1.0 - th.cos())
    K = torch.zeros(len(v), 3, 3, dtype=v.dtype, device=v.device)
    K[:, 0, 1], K[:, 0, 2] = -v[:, 2], v[:, 1]
    K[:, 1, 0], K[:, 1, 2] = v[:, 2], -v[:, 0]
    K[:, 2, 0], K[:, 2, 1] = -v[:, 1], v[:, 0]
    return fac1[:, None, None] * K + fac2[:, None, None] * (K @ K) + torch.eye(3, dtype=v.dtype, device=v.device)[None]


def pose_update(poses, trans, rot, mesh_diameter, rot_normalizer):
    """predict_pose_refine.py:195-231 (tracknet + normalize_xyz, axis_angle) + Utils.py:848-855.
    Returns new poses (N,4,4), trans_delta (N,3), rot_mat_delta (N,3,3)."""
    trans_delta = trans * (mesh_diameter / 2)
    rot_mat_delta = so3_exp_map(torch.tanh(rot) * rot_normalizer).permute(0, 2, 1)
    out = torch.eye(4, dtype=torch.float32, device=poses.device)[None].repeat(len(poses), 1, 1)
    out[:, :3, 3] = poses[:, :3, 3] + trans_delta
    out[:, :3, :3] = rot_mat_delta @ poses[:, :3, :3]
    return out, trans_delta, rot_mat_delta


def erode_depth(depth, radius=2, depth_diff_thres=0.001, ratio_thres=0.8, zfar=100):
    """Utils.py:359-384, including the fall-through for invalid centres."""
    d = np.asarray(depth, dtype=f32)
    H, W = d.shape
    bad = np.zeros((H, W), dtype=f32)
    total = np.zeros((H, W), dtype=f32)
    pad = np.full((H + 2 * radius, W + 2 * radius), np.nan, dtype=f32)
    pad[radius:-radius, radius:-radius] = d
    for du in range(-radius, radius + 1):
        for dv in range(-radius, radius + 1):
            cur = pad[radius + dv: radius + dv + H, radius + du: radius + du + W]
            inb = ~np.isnan(cur)
            total += inb
            with np.errstate(invalid="ignore"):
                isbad = (cur < 0.001) | (cur >= zfar) | (np.abs(cur - d) > f32(depth_diff_thres))
            bad += inb & isbad
    out = np.where(bad / total > f32(ratio_thres), f32(0), d)
    return out.astype(f32)


def bilateral_filter_depth(depth, radius=2, zfar=100, sigmaD=2, sigmaR=100000):
    """Utils.py:304-343."""
    d = np.asarray(depth, dtype=f32)
    H, W = d.shape
    pad = np.full((H + 2 * radius, W + 2 * radius), np.nan, dtype=f32)
    pad[radius:-radius, radius:-radius] = d
    shifts = [(du, dv) for du in range(-radius, radius + 1) for dv in range(-radius, radius + 1)]
    mean = np.zeros((H, W), dtype=f32)
    nvalid = np.zeros((H, W), dtype=np.int32)
    for du, dv in shifts:
        cur = pad[radius + dv: radius + dv + H, radius + du: radius + du + W]
        with np.errstate(invalid="ignore"):
            ok = (cur >= 0.001) & (cur < zfar)
        nvalid += ok
        mean = (mean + np.where(ok, cur, f32(0))).astype(f32)
    with np.errstate(invalid="ignore", divide="ignore"):
        mean = (mean / nvalid.astype(f32)).astype(f32)
    sw = np.zeros((H, W), dtype=f32)
    s = np.zeros((H, W), dtype=f32)
    for du, dv in shifts:
        cur = pad[radius + dv: radius + dv + H, radius + du: radius + du + W]
        with np.errstate(invalid="ignore"):
            ok = (cur >= 0.001) & (cur < zfar) & (np.abs(cur - mean) < f32(0.01))
            w = np.exp(-f32(du * du + dv * dv) / f32(2.0 * sigmaD * sigmaD) - (d - cur) * (d - cur) / f32(2.0 * sigmaR * sigmaR)).astype(f32)
        w = np.where(ok, w, f32(0))
        sw = (sw + w).astype(f32)
        s = (s + w * np.where(ok, cur, f32(0))).astype(f32)
    with np.errstate(invalid="ignore", divide="ignore"):
        out = np.where((sw > 0) & (nvalid > 0), s / sw, f32(0))
    return out.astype(f32)


def guess_translation(depth, mask, K):
    """estimater.py:137-156: centre of the mask's bounding box back-projected at the median valid masked depth; zeros for
    an empty mask or a mask without valid depth.  (The product's host copy, hypotheses.guess_translation, and the device
    kernel are held to the same reference body by tests/test_geometry_golden_cpu.py.)"""
    vs, us = np.where(np.asarray(mask) > 0)
    if len(us) == 0:
        return np.zeros(3)
    uc, vc = (us.min() + us.max()) / 2.0, (vs.min() + vs.max()) / 2.0
    valid = np.asarray(mask).astype(bool) & (np.asarray(depth) >= 0.001)
    if not valid.any():
        return np.zeros(3)
    zc = np.median(np.asarray(depth)[valid])
    return ((np.linalg.inv(K) @ np.asarray([uc, vc, 1.0]).reshape(3, 1)) * zc).reshape(3)
