"""An independent software implementation of the three nvdiffrast operations the reference calls — `rasterize`,
`interpolate`, `texture` (Utils.py:182-207) — written from nvdiffrast's PUBLISHED semantics, for
tools/make_golden_render.py only.  nvdiffrast itself is not in the image, so this is NOT the package's code; it exists so
that the reference's own, unmodified `nvdiffrast_render` (everything around those three calls: the OpenGL projection, the
clip-space bbox2d crop, which attribute is interpolated how, the lighting, clipping, masking and the vertical flips) can
be executed and compared with oracle/raster.py.

Deliberately different from oracle/raster.py in every free choice: clip-space input (what the reference hands over),
float64 homogeneous edge functions evaluated exactly at the pixel centres — no vertex snapping, no fixed point — and a
scipy-based texture filter.  Where the two disagree is therefore the sub-pixel-snapping / tie-rule sensitivity that
tests/test_raster_rules_cpu.py quantifies (silhouette pixels), not a difference of conventions.

R1  rast[..., 0:2] = perspective-correct barycentrics (u, v) of triangle vertices 0 and 1 (w2 = 1 - u - v), rast[..., 2] =
    z/w, rast[..., 3] = triangle id + 1 (0 = empty); nearest z/w wins, -1 <= z/w <= 1; pixel centres at +0.5; image row 0
    is the BOTTOM row (OpenGL window coordinates).
R2  interpolate(attr, rast, tri) = u a0 + v a1 + (1 - u - v) a2, zero on empty pixels; attr (V, C), (1, V, C) or (N, V, C).
R3  texture(tex, uv, 'linear'): bilinear, texel centres at +0.5, wrap addressing.
R4  barycentrics are computed in clip space (perspective-correct), both faces of a triangle are rendered.
"""
import numpy as np
import torch


class RasterizeCudaContext:
    def __init__(self, *a, **k):
        pass


RasterizeGLContext = RasterizeCudaContext


def rasterize(glctx, pos, tri, resolution, ranges=None, grad_db=True):
    pos = pos.detach().cpu().numpy().astype(np.float64)  # (N, V, 4) clip coordinates
    tri = tri.detach().cpu().numpy().astype(np.int64)
    H, W = int(resolution[0]), int(resolution[1])
    N = pos.shape[0]
    out = np.zeros((N, H, W, 4), dtype=np.float32)
    px = (np.arange(W) + 0.5) / W * 2 - 1  # NDC of the pixel centres
    py = (np.arange(H) + 0.5) / H * 2 - 1  # row 0 = bottom
    for n in range(N):
        best = np.full((H, W), np.inf)
        for f, (i0, i1, i2) in enumerate(tri):
            P = pos[n, [i0, i1, i2]]  # (3, 4): x, y, z, w
            M = np.stack([P[:, 0], P[:, 1], P[:, 3]])  # rows x, y, w; columns = vertices
            det = np.linalg.det(M)
            if abs(det) < 1e-30:
                continue
            Minv = np.linalg.inv(M)
            # pixel range: the NDC bounding box when the whole triangle is in front of the eye, else the whole image
            if (P[:, 3] > 1e-9).all():
                ndc = P[:, :2] / P[:, 3:4]
                j0 = max(int(np.floor((ndc[:, 0].min() + 1) / 2 * W - 0.5)) - 1, 0)
                j1 = min(int(np.ceil((ndc[:, 0].max() + 1) / 2 * W - 0.5)) + 1, W - 1)
                r0 = max(int(np.floor((ndc[:, 1].min() + 1) / 2 * H - 0.5)) - 1, 0)
                r1 = min(int(np.ceil((ndc[:, 1].max() + 1) / 2 * H - 0.5)) + 1, H - 1)
                if j0 > j1 or r0 > r1:
                    continue
            else:
                j0, j1, r0, r1 = 0, W - 1, 0, H - 1
            X, Y = np.meshgrid(px[j0:j1 + 1], py[r0:r1 + 1])
            lam = Minv[:, 0, None, None] * X + Minv[:, 1, None, None] * Y + Minv[:, 2, None, None]  # (3, h, w)
            s = lam.sum(0)
            inside = (lam >= 0).all(0) & (s > 0)
            if not inside.any():
                continue
            with np.errstate(divide="ignore", invalid="ignore"):
                b = lam / s  # perspective-correct barycentrics
                zw = (b[0] * P[0, 2] + b[1] * P[1, 2] + b[2] * P[2, 2]) / (b[0] * P[0, 3] + b[1] * P[1, 3] + b[2] * P[2, 3])
            ok = inside & (zw >= -1) & (zw <= 1) & (zw < best[r0:r1 + 1, j0:j1 + 1])
            if not ok.any():
                continue
            sub = out[n, r0:r1 + 1, j0:j1 + 1]
            sub[ok, 0] = b[0][ok]
            sub[ok, 1] = b[1][ok]
            sub[ok, 2] = zw[ok]
            sub[ok, 3] = f + 1
            best[r0:r1 + 1, j0:j1 + 1][ok] = zw[ok]
    return torch.from_numpy(out), None


def interpolate(attr, rast, tri, rast_db=None, diff_attrs=None):
    a = attr.detach().cpu().float()
    if a.dim() == 2:
        a = a[None]
    r = rast.detach().cpu()
    t = tri.detach().cpu().long()
    N, H, W, _ = r.shape
    fid = r[..., 3].long() - 1
    cov = fid >= 0
    out = torch.zeros(N, H, W, a.shape[-1], dtype=torch.float32)
    for n in range(N):
        an = a[n if a.shape[0] > 1 else 0]
        v = t[fid[n][cov[n]]]  # (P, 3)
        u_, v_ = r[n][cov[n]][:, 0:1], r[n][cov[n]][:, 1:2]
        out[n][cov[n]] = u_ * an[v[:, 0]] + v_ * an[v[:, 1]] + (1 - u_ - v_) * an[v[:, 2]]
    return out, None


def texture(tex, uv, filter_mode="linear", boundary_mode="wrap"):
    from scipy import ndimage

    assert filter_mode == "linear" and boundary_mode == "wrap"
    t = tex.detach().cpu().numpy().astype(np.float64)
    u = uv.detach().cpu().numpy().astype(np.float64)
    assert t.shape[0] == 1
    Ht, Wt, C = t.shape[1:]
    coords = np.stack([(u[..., 1] * Ht - 0.5).reshape(-1), (u[..., 0] * Wt - 0.5).reshape(-1)])
    out = np.stack([ndimage.map_coordinates(t[0, :, :, c], coords, order=1, mode="grid-wrap") for c in range(C)], -1)
    return torch.from_numpy(out.reshape(*u.shape[:-1], C).astype(np.float32))
