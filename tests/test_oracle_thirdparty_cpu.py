"""CPU: the oracle pieces that restate ABSENT third-party packages (pytorch3d, kornia, nvdiffrast — SURVEY.md §8c),
held to INDEPENDENT implementations of the same published algorithms that do exist in this image (scipy).  Not a pin
against the reference's own dependency — that stays "parity unpinned" in DESIGN.md §4 — but a known-answer check that
the restated maths is the maths: Rodrigues' formula, bilinear resampling under kornia's coordinate chain, bilinear
wrap-around texture filtering with texel centres at half-integers."""
import numpy as np
import torch
from scipy import ndimage
from scipy.spatial.transform import Rotation

from oracle import geometry, raster


def test_so3_exp_map_equals_scipy_rodrigues():
    """pytorch3d.transforms.so3_exp_map (P1) vs scipy's rotation-vector exponential, angles 0.02 .. 3 rad (the
    refiner's deltas are <= 0.35 rad); below pytorch3d's eps clamp (|v| < 0.01) both are I + [v]x to 1e-7."""
    rng = np.random.default_rng(0)
    axis = rng.standard_normal((200, 3))
    axis /= np.linalg.norm(axis, axis=1, keepdims=True)
    ang = rng.uniform(0.02, 3.0, (200, 1))
    v = (axis * ang).astype(np.float32)
    R = geometry.so3_exp_map(torch.from_numpy(v)).numpy()
    ref = Rotation.from_rotvec(v.astype(np.float64)).as_matrix()
    assert np.abs(R - ref).max() < 2e-6
    tiny = (axis[:20] * 1e-3).astype(np.float32)
    R = geometry.so3_exp_map(torch.from_numpy(tiny)).numpy()
    ref = Rotation.from_rotvec(tiny.astype(np.float64)).as_matrix()
    assert np.abs(R - ref).max() < 1e-6


def test_bilinear_warp_equals_scipy_map_coordinates_under_kornia_coordinates():
    """kornia.warp_perspective(bilinear, zeros padding, align_corners=False) for the axis-aligned crop transform (K1):
    destination pixel (i, j) samples source (y, x) = affine^-1 of the (size-1)-normalised, size-un-normalised chain.
    The oracle's torch op sequence must equal scipy's order-1 map_coordinates at the closed-form coordinates."""
    g = torch.Generator().manual_seed(1)
    H, W, S = 60, 80, 32
    src = torch.rand(1, 2, H, W, generator=g)
    left, top, size = 13.0, 7.0, 40.0  # integer-edged window like compute_crop_window_tf_batch
    s = S / size
    M = torch.tensor([[[s, 0, -left * s], [0, s, -top * s], [0, 0, 1]]], dtype=torch.float32)
    out = geometry.warp_perspective(src, M, (S, S), "bilinear")[0].numpy()
    # closed form: dst pixel j -> normalised xn = 2 j / (S-1) - 1 -> src normalised (affine inverse in normalised
    # space) -> grid_sample(align_corners=False) pixel ((xn' + 1) W - 1) / 2
    j = np.arange(S, dtype=np.float64)
    dst_pix = j  # (S-1)-normalisation followed by its own inverse is the identity on the destination side
    x_src_pix = dst_pix / s + left           # M^-1 in pixel units (kornia's src (size-1)-normalisation convention)
    y_src_pix = dst_pix / s + top
    xn = 2 * x_src_pix / (W - 1) - 1
    yn = 2 * y_src_pix / (H - 1) - 1
    xs = ((xn + 1) * W - 1) / 2
    ys = ((yn + 1) * H - 1) / 2
    yy, xx = np.meshgrid(ys, xs, indexing="ij")
    for c in range(2):
        ref = ndimage.map_coordinates(src[0, c].numpy().astype(np.float64), [yy, xx], order=1, mode="constant", cval=0.0)
        assert np.abs(out[c] - ref).max() < 2e-5


def test_texture_filter_equals_scipy_grid_wrap():
    """nvdiffrast dr.texture(filter_mode='linear', boundary_mode='wrap') (R3): texel centres at (i + 0.5) / size,
    wrap-around at the borders = scipy's order-1 'grid-wrap' interpolation at x = u W - 0.5."""
    rng = np.random.default_rng(2)
    tex = rng.integers(0, 256, (16, 24, 3), dtype=np.uint8)
    uv = rng.uniform(-0.3, 1.3, (500, 2)).astype(np.float32)  # includes coordinates outside [0, 1): they wrap
    got = raster._texture_linear_wrap(tex, uv)
    x = uv[:, 0].astype(np.float64) * 24 - 0.5
    y = uv[:, 1].astype(np.float64) * 16 - 0.5
    for c in range(3):
        ref = ndimage.map_coordinates(tex[..., c].astype(np.float64), [y, x], order=1, mode="grid-wrap") / 255.0
        assert np.abs(got[:, c] - ref).max() < 3e-5


def test_nearest_unwarp_equals_scipy_order0_away_from_ties():
    """h5_dataset.py:158 round trip (nearest): the closed-form `unwarp_nearest` equals scipy's order-0 sampling at
    kornia's coordinates wherever the coordinate is not within 1e-3 of a rounding tie (ties are
    implementation-defined in the reference, see the function's docstring)."""
    g = torch.Generator().manual_seed(3)
    S, H, W = 160, 120, 160
    crop = torch.rand(1, 1, S, S, generator=g)
    left, top, size = 31.0, 17.0, 73.0
    s = np.float32(S / size)
    win = dict(left=np.array([left], np.float32), top=np.array([top], np.float32), sx=np.array([s]), sy=np.array([s]))
    out = geometry.unwarp_nearest(crop, win, (H, W))[0, 0].numpy()
    u = np.arange(W, dtype=np.float64)
    v = np.arange(H, dtype=np.float64)
    jc = ((2 * (s * u - left * s) / (S - 1) - 1 + 1) * S - 1) / 2
    ic = ((2 * (s * v - top * s) / (S - 1) - 1 + 1) * S - 1) / 2
    safe_j = np.abs(jc - np.floor(jc) - 0.5) > 1e-3
    safe_i = np.abs(ic - np.floor(ic) - 0.5) > 1e-3
    yy, xx = np.meshgrid(ic, jc, indexing="ij")
    ref = ndimage.map_coordinates(crop[0, 0].numpy().astype(np.float64), [yy, xx], order=0, mode="constant", cval=0.0)
    inside = (yy > -0.5 + 1e-3) & (yy < S - 0.5 - 1e-3) & (xx > -0.5 + 1e-3) & (xx < S - 0.5 - 1e-3)
    ref = np.where(inside, ref, 0.0)
    m = np.outer(safe_i, safe_j) & (inside | (yy < -0.5 - 1e-3) | (yy > S - 0.5 + 1e-3) | (xx < -0.5 - 1e-3) | (xx > S - 0.5 + 1e-3))
    assert m.mean() > 0.9
    assert np.abs(out[m] - ref[m]).max() == 0
