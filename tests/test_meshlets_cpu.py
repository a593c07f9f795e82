"""CPU: the host-side mesh preparation of fp_set_mesh (csrc/fp_meshlet.cu) through the C ABI (no GPU needed):
meshlets partition the faces, respect the 64 / 64 limits, closed / open / inside-out meshes are told apart."""
import ctypes as C

import numpy as np
import pytest


def _build(verts, faces):
    from foundationpose_b200 import _lib

    lib = _lib.lib
    lib.fp_op_build_meshlets.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.c_void_p, C.c_void_p]
    lib.fp_op_build_meshlets.restype = C.c_int
    pos = np.ascontiguousarray(verts, dtype=np.float32)
    fc = np.ascontiguousarray(faces, dtype=np.int32)
    info = (C.c_int * 6)()
    face_of = np.full(len(fc), -1, dtype=np.int32)
    rc = lib.fp_op_build_meshlets(len(pos), len(fc), pos.ctypes.data, fc.ctypes.data, info, face_of.ctypes.data, None)
    assert rc == 0, _lib.lib.fp_last_error()
    return dict(meshlets=info[0], closed=info[1], front_sign=info[2], max_tris=info[3], max_verts=info[4], total=info[5]), face_of


@pytest.mark.parametrize("sub", [0, 2, 4])
def test_icosphere_partition(sub):
    from foundationpose_b200 import synth

    v, f = synth.icosphere(sub)
    info, face_of = _build(v * synth.RADII, f)
    assert info["total"] == len(f) and sorted(face_of.tolist()) == list(range(len(f))), "every face exactly once"
    assert info["max_tris"] <= 64 and info["max_verts"] <= 64
    assert info["closed"] == 1 and info["front_sign"] == -1  # outward-oriented closed surface
    assert info["meshlets"] >= -(-len(f) // 64)
    if sub == 4:  # 5120 faces: the octant split + vertex limit should cost well under 2x the ideal count
        assert info["meshlets"] <= 2 * (len(f) // 64)


def test_inside_out_and_open_meshes():
    from foundationpose_b200 import synth

    v, f = synth.icosphere(2)
    info, _ = _build(v, f[:, ::-1])
    assert info["closed"] == 1 and info["front_sign"] == 1
    info, _ = _build(v, f[:-1])  # one face missing: open
    assert info["closed"] == 0 and info["front_sign"] == 0
    flipped = f.copy()
    flipped[0] = flipped[0, ::-1]  # one inconsistently wound face
    info, _ = _build(v, flipped)
    assert info["closed"] == 0 and info["front_sign"] == 0


def test_texture_seam_duplicates_are_welded():
    """OBJ loaders duplicate the vertices along UV seams; closedness is decided on positions, not indices."""
    from foundationpose_b200 import synth

    v, f = synth.icosphere(1)
    v2 = np.concatenate([v, v[f[0]]])  # duplicate the three vertices of face 0 ...
    f2 = f.copy()
    f2[0] = [len(v), len(v) + 1, len(v) + 2]  # ... and let face 0 use the copies
    info, face_of = _build(v2, f2)
    assert info["closed"] == 1 and info["front_sign"] == -1
    assert sorted(face_of.tolist()) == list(range(len(f2)))


def test_degenerate_and_large_mesh():
    from foundationpose_b200 import synth

    v, f = synth.icosphere(5)
    rng = np.random.default_rng(0)
    f = f[rng.permutation(len(f))]  # arbitrary face order
    f = np.concatenate([f, [[0, 0, 1]]])  # a degenerate face must not break anything
    info, face_of = _build(v, f)
    assert info["total"] == len(f) and sorted(face_of.tolist()) == list(range(len(f)))
    assert info["max_tris"] <= 64 and info["max_verts"] <= 64
