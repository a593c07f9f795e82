"""CPU software rasteriser = the oracle of csrc/fp_crop.cu's render pass (TEST INFRASTRUCTURE;
PARITY UNPINNED w.r.t. nvdiffrast, whose source is not in the reference tree — SURVEY.md §8c R1-R4).

Follows Utils.py:133-219 (nvdiffrast_render with bbox2d, use_light=True): project with K, restrict to the
crop window, rasterise nearest-depth triangles at pixel centres (+0.5), perspective-correct
interpolation of camera-space xyz / uv / per-vertex diffuse term, bilinear wrap texture fetch,
`color*0.8 + diffuse*color*0.5`, clip, mask, top-down rows (the reference's y-flips are folded in).

Coverage rule (shared with the CUDA kernel so that coverage and triangle ids are bit-identical):
vertices snapped to 1/256 px, exact integer edge functions, tie rule dy>0 or (dy==0 and dx>0),
depth test on interpolated 1/Z (largest wins), ties -> lowest triangle id.
"""
import numpy as np

f32 = np.float32
S = 160


def _project(pose, verts, K, umin, vmin, rsx, rsy):
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
    xi = np.rint(px * f32(256)).astype(np.int64)
    yi = np.rint(py * f32(256)).astype(np.int64)
    return X.astype(f32), Y.astype(f32), Z.astype(f32), iz, xi, yi


def _edge_ok(e, dx, dy):
    return (e > 0) | ((e == 0) & ((dy > 0) | ((dy == 0) & (dx > 0))))


def rasterize(pose, verts, faces, K, window, znear=0.001):
    """Returns tri_id (S,S) int64 (-1 = empty), bary (S,S,3) float32 screen-space weights of the
    face's vertices in their original order, and the per-vertex camera data."""
    umin, vmin, umax, vmax = [f32(w) for w in window]
    rsx = f32(S) / (umax - umin)
    rsy = f32(S) / (vmax - vmin)
    X, Y, Z, iz, xi, yi = _project(pose, verts, K, umin, vmin, rsx, rsy)
    best_key = np.zeros((S, S), dtype=np.float32)  # 1/Z of the winner (0 = empty)
    tri_id = np.full((S, S), -1, dtype=np.int64)
    bary = np.zeros((S, S, 3), dtype=f32)
    for f, (i0, i1, i2) in enumerate(faces):
        if not (Z[i0] > znear and Z[i1] > znear and Z[i2] > znear):
            continue
        x0, y0, x1, y1, x2, y2 = int(xi[i0]), int(yi[i0]), int(xi[i1]), int(yi[i1]), int(xi[i2]), int(yi[i2])
        area2 = (x1 - x0) * (y2 - y0) - (y1 - y0) * (x2 - x0)
        if area2 == 0:
            continue
        swapped = area2 < 0
        if swapped:
            x1, y1, x2, y2 = x2, y2, x1, y1
            area2 = -area2
        j0 = max((min(x0, x1, x2) + 127) >> 8, 0)
        j1 = min((max(x0, x1, x2) - 128) >> 8, S - 1)
        r0 = max((min(y0, y1, y2) + 127) >> 8, 0)
        r1 = min((max(y0, y1, y2) - 128) >> 8, S - 1)
        if j0 > j1 or r0 > r1:
            continue
        py, px = np.meshgrid(np.arange(r0, r1 + 1) * 256 + 128, np.arange(j0, j1 + 1) * 256 + 128, indexing="ij")
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
        win = inside & (izp.view(np.uint32) > sub_key.view(np.uint32))
        if not win.any():
            continue
        sub_key[win] = izp[win]
        tri_id[r0:r1 + 1, j0:j1 + 1][win] = f
        sb = bary[r0:r1 + 1, j0:j1 + 1]
        sb[win] = np.stack([b0, b1, b2], -1)[win]
    return tri_id, bary, (X, Y, Z, iz)


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


def render_crop(pose, mesh, K, window, w_ambient=0.8, w_diffuse=0.5):
    """One hypothesis: returns rgb (S,S,3) in 0..1, xyz (S,S,3) camera-space metres (0 on background),
    tri_id (S,S).  mesh: dict(pos, normals, faces, uv [v-flipped] + tex uint8 | vcolor)."""
    pose = np.asarray(pose, dtype=f32)
    tri_id, bary, (X, Y, Z, iz) = rasterize(pose, mesh["pos"], mesh["faces"], K, window)
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
    w = (b * izv / izp[:, None]).astype(f32)  # perspective-correct weights (R4)
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
