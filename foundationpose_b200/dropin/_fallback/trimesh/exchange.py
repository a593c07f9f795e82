"""Wavefront OBJ / MTL and Stanford PLY readers and writers."""
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


_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2",
              "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4", "double": "f8", "float64": "f8"}


def load_ply(path):
    """Stanford PLY, ascii or binary (either endianness): vertex x y z [nx ny nz] [red green blue [alpha]]
    [texture_u texture_v | s t], face `vertex_indices` lists (fan-triangulated).  The BOP model files
    (datareader.py:291-296, :489-505) are of this kind."""
    with open(path, "rb") as fh:
        if fh.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, elements = None, []  # elements: [name, count, [(prop, type) | (prop, ('list', count_type, item_type))]]
        while True:
            line = fh.readline()
            if not line:
                raise ValueError(f"{path}: truncated PLY header")
            t = line.decode("ascii", "replace").split()
            if not t or t[0] == "comment" or t[0] == "obj_info":
                continue
            if t[0] == "format":
                fmt = t[1]
            elif t[0] == "element":
                elements.append([t[1], int(t[2]), []])
            elif t[0] == "property":
                if t[1] == "list":
                    elements[-1][2].append((t[4], ("list", _PLY_TYPES[t[2]], _PLY_TYPES[t[3]])))
                else:
                    elements[-1][2].append((t[2], _PLY_TYPES[t[1]]))
            elif t[0] == "end_header":
                break
        data = {}
        if fmt == "ascii":
            tokens = fh.read().split()
            pos = 0
            for name, count, props in elements:
                cols = {pn: [] for pn, _ in props}
                for _ in range(count):
                    for pn, pt in props:
                        if isinstance(pt, tuple):
                            n = int(tokens[pos]); pos += 1
                            cols[pn].append([int(float(x)) for x in tokens[pos:pos + n]]); pos += n
                        else:
                            cols[pn].append(float(tokens[pos])); pos += 1
                data[name] = cols
        else:
            end = "<" if fmt == "binary_little_endian" else ">"
            for name, count, props in elements:
                if all(not isinstance(pt, tuple) for _, pt in props):
                    dt = np.dtype([(pn, end + pt) for pn, pt in props])
                    arr = np.frombuffer(fh.read(dt.itemsize * count), dtype=dt, count=count)
                    data[name] = {pn: arr[pn] for pn, _ in props}
                else:
                    cols = {pn: [] for pn, _ in props}
                    for _ in range(count):
                        for pn, pt in props:
                            if isinstance(pt, tuple):
                                n = int(np.frombuffer(fh.read(np.dtype(pt[1]).itemsize), dtype=end + pt[1])[0])
                                cols[pn].append(np.frombuffer(fh.read(np.dtype(pt[2]).itemsize * n), dtype=end + pt[2]).astype(np.int64).tolist())
                            else:
                                cols[pn].append(np.frombuffer(fh.read(np.dtype(pt).itemsize), dtype=end + pt)[0])
                    data[name] = cols
    V = data.get("vertex", {})
    vertices = np.stack([np.asarray(V[k], dtype=np.float64) for k in ("x", "y", "z")], 1)
    normals = np.stack([np.asarray(V[k], dtype=np.float64) for k in ("nx", "ny", "nz")], 1) if all(k in V for k in ("nx", "ny", "nz")) else None
    faces = []
    F = data.get("face", {})
    for key in ("vertex_indices", "vertex_index"):
        if key in F:
            for poly in F[key]:
                for k in range(1, len(poly) - 1):
                    faces.append([poly[0], poly[k], poly[k + 1]])
            break
    vis = None
    uvk = ("texture_u", "texture_v") if "texture_u" in V else (("s", "t") if "s" in V else None)
    if all(k in V for k in ("red", "green", "blue")):
        ch = [np.asarray(V[k]) for k in ("red", "green", "blue")] + ([np.asarray(V["alpha"])] if "alpha" in V else [])
        vis = _visual.ColorVisuals(np.stack(ch, 1).astype(np.uint8))
    elif uvk is not None:
        vis = _visual.TextureVisuals(uv=np.stack([np.asarray(V[uvk[0]], dtype=np.float64), np.asarray(V[uvk[1]], dtype=np.float64)], 1))
    m = Trimesh(vertices, np.asarray(faces, dtype=np.int64).reshape(-1, 3), normals, vis)
    if vis is not None and uvk is not None and vis.uv is None:
        vis.uv = np.stack([np.asarray(V[uvk[0]], dtype=np.float64), np.asarray(V[uvk[1]], dtype=np.float64)], 1)
    return m


def export_ply(mesh, path):
    """binary_little_endian PLY with normals and (when present) per-vertex colours."""
    vn = mesh.vertex_normals
    vc = getattr(mesh.visual, "vertex_colors", None)
    fields = [("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("nx", "<f4"), ("ny", "<f4"), ("nz", "<f4")]
    if vc is not None:
        fields += [("red", "u1"), ("green", "u1"), ("blue", "u1"), ("alpha", "u1")]
    arr = np.zeros(len(mesh.vertices), dtype=np.dtype(fields))
    for k, col in zip(("x", "y", "z"), mesh.vertices.T):
        arr[k] = col
    for k, col in zip(("nx", "ny", "nz"), vn.T):
        arr[k] = col
    if vc is not None:
        for k, col in zip(("red", "green", "blue", "alpha"), np.asarray(vc, dtype=np.uint8).T):
            arr[k] = col
    fa = np.zeros(len(mesh.faces), dtype=np.dtype([("n", "u1"), ("i", "<i4", (3,))]))
    fa["n"] = 3
    fa["i"] = mesh.faces
    names = {"<f4": "float", "u1": "uchar"}
    with open(str(path), "wb") as fh:
        hdr = ["ply", "format binary_little_endian 1.0", f"element vertex {len(arr)}"]
        hdr += [f"property {names[t]} {n}" for n, t in fields]
        hdr += [f"element face {len(fa)}", "property list uchar int vertex_indices", "end_header"]
        fh.write(("\n".join(hdr) + "\n").encode("ascii"))
        fh.write(arr.tobytes())
        fh.write(fa.tobytes())
    return str(path)


def load(path, *args, **kwargs):
    ext = os.path.splitext(str(path))[1].lower()
    if ext == ".obj":
        return load_obj(str(path))
    if ext == ".ply":
        return load_ply(str(path))
    raise NotImplementedError(f"trimesh stand-in: Wavefront OBJ and Stanford PLY are supported, got '{ext}' (install trimesh for other formats)")


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
