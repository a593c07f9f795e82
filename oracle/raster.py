"""CPU software rasteriser = the oracle of csrc/fp_crop.cu's render pass (TEST INFRASTRUCTURE;
PARITY UNPINNED w.r.t. nvdiffrast, whose source is not in the reference tree — SURVEY.md §8c R1-R4).

Follows Utils.py:133-219 (nvdiffrast_render with bbox2d, use_light=True): project with K, restrict to the
crop window, rasterise nearest-depth triangles at pixel centres (+0.5), perspective-correct
interpolation of camera-space xyz / uv / per-vertex diffuse term, bilinear wrap texture fetch,
`color*0.8 + diffuse*color*0.5`, clip, mask, top-down rows (the reference's y-flips are folded in).

Coverage rule (shared with the CUDA kernel so that coverage and triangle ids are bit-identical):
vertices snapped to 1/2^SUBPIX_BITS px (8 by default), exact integer edge functions, tie rule dy>0 or
(dy==0 and dx>0), depth test on interpolated 1/Z (largest wins), ties -> lowest triangle id, depth range
znear < Z < zfar (Utils.py:161).  Both sides of every triangle are rendered (nvdiffrast does not cull).
Triangles that cross the near plane are not dropped: their coverage and perspective-correct barycentrics
come from the homogeneous (clip-space) form  [P0 P1 P2] w = ray  — nvdiffrast computes its barycentrics in
clip space too — with the depth range test per pixel.

`subpix_bits=4` gives the nvdiffrast-STYLE variant used by tests/test_raster_rules_cpu.py to state how many
silhouette pixels move when the sub-pixel grid changes (cudaraster snaps to 1/16 px as far as is known; its
source is not in the reference tree, so this is a sensitivity statement, not a pinned parity).
"""
import numpy as np

f32 = np.float32
S = 160


def _project(pose, verts, K, umin, vmin, rsx, rsy, subpix_bits=8):
    """fp32, same operation order as xform_vertex() in fp_crop.cu (no fused multiply-add)."""
    P = pose.astype(f32)
    x, y, z = verts[:, 0].astype(f32), verts[:, 1].astype(f32), verts[:, 2].astype(f32)
    X = ((P[0, 0] * x + P[0, 1] * y) + P[0, 2] * z) + P[0, 3]
    Y = ((P[1, 0] * x + P[1, 1] * y) + P[1, 2] * z) + P[1, 3]
    Z = ((P[2, 0] * x + P[2, 1] * y) + P[2, 2] * z) + P[2, 3]
    with np.errstate(divide="ignore"):
        iz = (f32(1) / Z).astype(f32)
    u = ((f32(K[0, 0]) * X) * iz + f32(K[0, 2])).astype(f32)
    v = ((f32(K[1, 1]) * Y) * iz + f32(K[1, 2])).astype(f32)
    px = np.clip((u - umin) * rsx, -30000, 30000).astype(f32)
    py = np.clip((v - vmin) * rsy, -30000, 30000).astype(f32)
    xi = np.rint(px * f32(1 << subpix_bits)).astype(np.int64)
    yi = np.rint(py * f32(1 << subpix_bits)).astype(np.int64)
    return X.astype(f32), Y.astype(f32), Z.astype(f32), iz, xi, yi


def _edge_ok(e, dx, dy):
    return (e > 0) | ((e == 0) & ((dy > 0) | ((dy == 0) & (dx > 0))))


def _hom_cover(P0, P1, P2, dx, dy, znear, zfar):
    """Homogeneous coverage of one triangle (camera-space vertices) for pixel rays (dx, dy, 1): solves
    [P0 P1 P2] w = d.  Returns inside mask, perspective-correct weights (3, ...) and 1/Z.  fp32, same
    operation order as hom_setup / hom_cover in csrc/fp_crop.cu."""
    def cross(a, b):
        return (f32(a[1] * b[2]) - f32(a[2] * b[1]), f32(a[2] * b[0]) - f32(a[0] * b[2]), f32(a[0] * b[1]) - f32(a[1] * b[0]))

    n0, n1, n2 = cross(P1, P2), cross(P2, P0), cross(P0, P1)
    det = f32(f32(f32(P0[0] * n0[0]) + f32(P0[1] * n0[1])) + f32(P0[2] * n0[2]))
    if det == 0:
        z = np.zeros_like(dx, dtype=bool)
        return z, None, None
    ws = []
    for n in (n0, n1, n2):
        ws.append((((n[0] * dx).astype(f32) + (n[1] * dy).astype(f32)).astype(f32) + n[2]).astype(f32) / det)
    w0, w1, w2 = [w.astype(f32) for w in ws]
    iz = ((w0 + w1).astype(f32) + w2).astype(f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        z = (f32(1) / iz).astype(f32)
        inside = (w0 >= 0) & (w1 >= 0) & (w2 >= 0) & (iz > 0) & (z > f32(znear)) & (z < f32(zfar))
    lam = np.stack([(w0 * z).astype(f32), (w1 * z).astype(f32), (w2 * z).astype(f32)], -1)
    return inside, lam, iz


def rasterize(pose, verts, faces, K, window, znear=0.001, zfar=100.0, subpix_bits=8):
    """Returns tri_id (S,S) int64 (-1 = empty), bary (S,S,3) float32 screen-space weights of the
    face's vertices in their original order (for pixels won by a near-plane-crossing triangle: the
    perspective-correct weights, flagged in `persp` (S,S) bool), and the per-vertex camera data."""
    umin, vmin, umax, vmax = [f32(w) for w in window]
    rsx = f32(S) / (umax - umin)
    rsy = f32(S) / (vmax - vmin)
    X, Y, Z, iz, xi, yi = _project(pose, verts, K, umin, vmin, rsx, rsy, subpix_bits)
    one = 1 << subpix_bits
    half = one >> 1
    best_key = np.zeros((S, S), dtype=np.float32)  # 1/Z of the winner (0 = empty)
    tri_id = np.full((S, S), -1, dtype=np.int64)
    bary = np.zeros((S, S, 3), dtype=f32)
    persp = np.zeros((S, S), dtype=bool)
    iz_far = f32(1.0) / f32(zfar)
    # pixel rays of the crop (only used by near-plane-crossing triangles)
    jj = (np.arange(S, dtype=f32) + f32(0.5))
    ray_x = (((umin + (jj / rsx).astype(f32)).astype(f32) - f32(K[0, 2])) / f32(K[0, 0])).astype(f32)
    ray_y = (((vmin + (jj / rsy).astype(f32)).astype(f32) - f32(K[1, 2])) / f32(K[1, 1])).astype(f32)
    for f, (i0, i1, i2) in enumerate(faces):
        nfront = int(Z[i0] > znear) + int(Z[i1] > znear) + int(Z[i2] > znear)
        if nfront == 0:
            continue
        if nfront < 3:
            dy, dx = np.meshgrid(ray_y, ray_x, indexing="ij")
            inside, lam, izp = _hom_cover((X[i0], Y[i0], Z[i0]), (X[i1], Y[i1], Z[i1]), (X[i2], Y[i2], Z[i2]), dx, dy, znear, zfar)
            if not inside.any():
                continue
            win = inside & (izp.view(np.uint32) > best_key.view(np.uint32))
            best_key[win] = izp[win]
            tri_id[win] = f
            bary[win] = lam[win]
            persp[win] = True
            continue
        x0, y0, x1, y1, x2, y2 = int(xi[i0]), int(yi[i0]), int(xi[i1]), int(yi[i1]), int(xi[i2]), int(yi[i2])
        area2 = (x1 - x0) * (y2 - y0) - (y1 - y0) * (x2 - x0)
        if area2 == 0:
            continue
        swapped = area2 < 0
        if swapped:
            x1, y1, x2, y2 = x2, y2, x1, y1
            area2 = -area2
        j0 = max((min(x0, x1, x2) + half - 1) >> subpix_bits, 0)
        j1 = min((max(x0, x1, x2) - half) >> subpix_bits, S - 1)
        r0 = max((min(y0, y1, y2) + half - 1) >> subpix_bits, 0)
        r1 = min((max(y0, y1, y2) - half) >> subpix_bits, S - 1)
        if j0 > j1 or r0 > r1:
            continue
        py, px = np.meshgrid(np.arange(r0, r1 + 1) * one + half, np.arange(j0, j1 + 1) * one + half, indexing="ij")
        e0 = (x2 - x1) * (py - y1) - (y2 - y1) * (px - x1)
        e1 = (x0 - x2) * (py - y2) - (y0 - y2) * (px - x2)
        e2 = area2 - e0 - e1
        inside = _edge_ok(e0, x2 - x1, y2 - y1) & _edge_ok(e1, x0 - x2, y0 - y2) & _edge_ok(e2, x1 - x0, y1 - y0)
        if not inside.any():
            continue
        fa = f32(area2)
        b0 = (e0.astype(f32) / fa).astype(f32)
        w1 = (e1.astype(f32) / fa).astype(f32)
        w2 = (e2.astype(f32) / fa).astype(f32)
        b1, b2 = (w2, w1) if swapped else (w1, w2)
        izp = ((b0 * iz[i0] + b1 * iz[i1]) + b2 * iz[i2]).astype(f32)
        sub_key = best_key[r0:r1 + 1, j0:j1 + 1]
        # strictly nearer wins; ties keep the lower triangle id (faces are visited in increasing id)
        win = inside & (izp > iz_far) & (izp.view(np.uint32) > sub_key.view(np.uint32))
        if not win.any():
            continue
        sub_key[win] = izp[win]
        tri_id[r0:r1 + 1, j0:j1 + 1][win] = f
        sb = bary[r0:r1 + 1, j0:j1 + 1]
        sb[win] = np.stack([b0, b1, b2], -1)[win]
        persp[r0:r1 + 1, j0:j1 + 1][win] = False
    return tri_id, bary, (X, Y, Z, iz), persp


def _texture_linear_wrap(tex_u8, uv):
    """dr.texture(tex, uv, filter_mode='linear') with the default wrap boundary mode (R3)."""
    Ht, Wt = tex_u8.shape[:2]
    x = uv[:, 0].astype(f32) * f32(Wt) - f32(0.5)
    y = uv[:, 1].astype(f32) * f32(Ht) - f32(0.5)
    xf, yf = np.floor(x), np.floor(y)
    ax, ay = (x - xf).astype(f32), (y - yf).astype(f32)
    x0 = np.mod(xf.astype(np.int64), Wt)
    y0 = np.mod(yf.astype(np.int64), Ht)
    x1 = np.mod(x0 + 1, Wt)
    y1 = np.mod(y0 + 1, Ht)
    t = tex_u8.astype(f32)
    w00, w01, w10, w11 = (1 - ax) * (1 - ay), ax * (1 - ay), (1 - ax) * ay, ax * ay
    c = w00[:, None] * t[y0, x0] + w01[:, None] * t[y0, x1] + w10[:, None] * t[y1, x0] + w11[:, None] * t[y1, x1]
    return (c * f32(1.0 / 255.0)).astype(f32)


def render_crop(pose, mesh, K, window, w_ambient=0.8, w_diffuse=0.5, subpix_bits=8):
    """One hypothesis: returns rgb (S,S,3) in 0..1, xyz (S,S,3) camera-space metres (0 on background),
    tri_id (S,S).  mesh: dict(pos, normals, faces, uv [v-flipped] + tex uint8 | vcolor)."""
    pose = np.asarray(pose, dtype=f32)
    tri_id, bary, (X, Y, Z, iz), persp = rasterize(pose, mesh["pos"], mesh["faces"], K, window, subpix_bits=subpix_bits)
    cov = tri_id >= 0
    rgb = np.zeros((S, S, 3), dtype=f32)
    xyz = np.zeros((S, S, 3), dtype=f32)
    if not cov.any():
        return rgb, xyz, tri_id
    fid = tri_id[cov]
    vi = mesh["faces"][fid]  # (P,3)
    b = bary[cov]
    izv = iz[vi]  # (P,3)
    izp = ((b[:, 0] * izv[:, 0] + b[:, 1] * izv[:, 1]) + b[:, 2] * izv[:, 2]).astype(f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = (b * izv / izp[:, None]).astype(f32)  # perspective-correct weights (R4)
    w[persp[cov]] = b[persp[cov]]  # near-plane path: the weights are perspective-correct already
    cam = np.stack([X, Y, Z], -1)
    xyz[cov] = (w[..., None] * cam[vi]).sum(1)
    # per-vertex diffuse term: clip(normalize(R n) . (0,0,-1), 0, 1)   (Utils.py:203-207)
    n_cam = mesh["normals"].astype(f32) @ pose[:3, :3].T
    n_cam = n_cam / np.maximum(np.linalg.norm(n_cam, axis=1, keepdims=True), 1e-12)
    dif_v = np.clip(-n_cam[:, 2], 0, 1).astype(f32)
    diffuse = (w * dif_v[vi]).sum(1)
    if mesh.get("tex") is not None:
        uv = (w[..., None] * mesh["uv"].astype(f32)[vi]).sum(1)
        col = _texture_linear_wrap(mesh["tex"], uv)
    else:
        col = (w[..., None] * mesh["vcolor"].astype(f32)[vi]).sum(1)
    shaded = np.clip(col * f32(w_ambient) + diffuse[:, None] * col * f32(w_diffuse), 0, 1)
    rgb[cov] = shaded
    return rgb, xyz, tri_id
