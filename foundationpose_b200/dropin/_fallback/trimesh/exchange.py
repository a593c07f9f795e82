"""Wavefront OBJ / MTL reader and writer."""
import os

import numpy as np

from . import visual as _visual
from .base import Trimesh


def _parse_mtl(path):
    """-> {material name: texture file path} from `map_Kd` statements."""
    out, cur = {}, None
    if not os.path.exists(path):
        return out
    with open(path) as fh:
        for line in fh:
            t = line.split()
            if not t:
                continue
            if t[0] == "newmtl":
                cur = " ".join(t[1:])
            elif t[0] == "map_Kd" and cur is not None:
                out[cur] = os.path.join(os.path.dirname(path), t[-1])
    return out


def load_obj(path):
    v, vt, vn, vc = [], [], [], []
    corners = []  # per face corner: (v, vt, vn) indices, -1 = absent
    mtllib, usemtl = None, None
    with open(path) as fh:
        for line in fh:
            t = line.split()
            if not t or t[0].startswith("#"):
                continue
            if t[0] == "v":
                v.append([float(x) for x in t[1:4]])
                if len(t) >= 7:
                    vc.append([float(x) for x in t[4:7]])
            elif t[0] == "vt":
                vt.append([float(x) for x in t[1:3]])
            elif t[0] == "vn":
                vn.append([float(x) for x in t[1:4]])
            elif t[0] == "f":
                poly = []
                for c in t[1:]:
                    idx = (c.split("/") + ["", ""])[:3]
                    ids = []
                    for s, pool in zip(idx, (v, vt, vn)):
                        if s == "":
                            ids.append(-1)
                        else:
                            i = int(s)
                            ids.append(i - 1 if i > 0 else len(pool) + i)
                    poly.append(tuple(ids))
                for k in range(1, len(poly) - 1):  # fan triangulation
                    corners.append((poly[0], poly[k], poly[k + 1]))
            elif t[0] == "mtllib":
                mtllib = " ".join(t[1:])
            elif t[0] == "usemtl" and usemtl is None:
                usemtl = " ".join(t[1:])
    v = np.asarray(v, dtype=np.float64).reshape(-1, 3)
    has_vt = len(vt) > 0 and all(c[1] >= 0 for tri in corners for c in tri)
    has_vn = len(vn) > 0 and all(c[2] >= 0 for tri in corners for c in tri)
    # one output vertex per distinct (position, uv) pair, in first-use order (texture seams duplicate positions)
    remap, verts, uvs, nrms, faces = {}, [], [], [], []
    for tri in corners:
        f = []
        for (iv, it, inn) in tri:
            key = (iv, it if has_vt else -1)
            j = remap.get(key)
            if j is None:
                j = remap[key] = len(verts)
                verts.append(iv)
                uvs.append(it)
                nrms.append(inn)
            f.append(j)
        faces.append(f)
    verts = np.asarray(verts, dtype=np.int64)
    vertices = v[verts]
    normals = np.asarray(vn, dtype=np.float64)[np.asarray(nrms)] if has_vn else None
    vis = None
    if has_vt:
        tex_path = None
        if mtllib is not None:
            maps = _parse_mtl(os.path.join(os.path.dirname(path), mtllib))
            tex_path = maps.get(usemtl) if usemtl in maps else (next(iter(maps.values())) if maps else None)
        if tex_path is not None and os.path.exists(tex_path):
            from PIL import Image

            vis = _visual.TextureVisuals(uv=np.asarray(vt, dtype=np.float64)[np.asarray(uvs)], image=Image.open(tex_path).convert("RGB"))
    if vis is None and len(vc) == len(v) and len(v) > 0:
        vis = _visual.ColorVisuals(np.asarray(vc, dtype=np.float64)[verts])
    return Trimesh(vertices, np.asarray(faces, dtype=np.int64), normals, vis)


def load(path, *args, **kwargs):
    ext = os.path.splitext(str(path))[1].lower()
    if ext != ".obj":
        raise NotImplementedError(f"trimesh stand-in: only Wavefront OBJ is supported, got '{ext}' (install trimesh for other formats)")
    return load_obj(str(path))


def export_obj(mesh, path):
    """Writes <path> (+ .mtl and texture .png when the mesh is textured)."""
    path = str(path)
    stem = os.path.splitext(os.path.basename(path))[0]
    uv = getattr(mesh.visual, "uv", None)
    img = getattr(getattr(mesh.visual, "material", None), "image", None)
    vn = mesh.vertex_normals
    with open(path, "w") as fh:
        if uv is not None and img is not None:
            fh.write(f"mtllib {stem}.mtl\nusemtl material_0\n")
        for p in mesh.vertices:
            fh.write(f"v {p[0]:.9g} {p[1]:.9g} {p[2]:.9g}\n")
        for n in vn:
            fh.write(f"vn {n[0]:.9g} {n[1]:.9g} {n[2]:.9g}\n")
        if uv is not None:
            for t in uv:
                fh.write(f"vt {t[0]:.9g} {t[1]:.9g}\n")
        for f in mesh.faces + 1:
            if uv is not None:
                fh.write("f " + " ".join(f"{i}/{i}/{i}" for i in f) + "\n")
            else:
                fh.write("f " + " ".join(f"{i}//{i}" for i in f) + "\n")
    if uv is not None and img is not None:
        with open(os.path.join(os.path.dirname(path), stem + ".mtl"), "w") as fh:
            fh.write(f"newmtl material_0\nKa 1 1 1\nKd 1 1 1\nKs 0 0 0\nmap_Kd {stem}.png\n")
        img.save(os.path.join(os.path.dirname(path), stem + ".png"))
    return path
