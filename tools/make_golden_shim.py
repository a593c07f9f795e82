"""Golden vectors for the helper functions of the drop-in `Utils` module, produced by the REFERENCE's own function
bodies.

/root/reference/Utils.py cannot be imported (pytorch3d, nvdiffrast, open3d, trimesh, transformations, ... are absent), so
— as tools/make_golden_geometry.py does for the hot-path geometry — the unmodified source of each helper is extracted
with `ast` and executed in a namespace holding the modules it really uses (numpy, cv2, torch, collections).  The one
third-party function among them, `transformations.euler_matrix` (static xyz), is supplied by scipy
(`Rotation.from_euler('xyz', ...)`, extrinsic x-y-z), an independent implementation of the same convention.

    python tools/make_golden_shim.py      # needs /root/reference; writes tests/golden/shim_golden.npz

tests/test_dropin_golden_cpu.py holds foundationpose_b200/dropin/Utils.py to these vectors.
"""
import logging
import os
import sys
from collections import OrderedDict, defaultdict

import cv2
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
REF = os.environ.get("FPOSE_REFERENCE", "/root/reference")

NAMES = ("NestDict", "set_seed", "to_homo", "transform_pts", "project_3d_to_2d", "draw_xyz_axis", "draw_posed_3d_box",
         "symmetry_tfs_from_info", "make_yaml_dumpable", "depth2xyzmap", "depth2xyzmap_batch", "compute_mesh_diameter")


def euler_matrix(ai, aj, ak, axes="sxyz"):
    from scipy.spatial.transform import Rotation

    assert axes == "sxyz"
    M = np.eye(4)
    M[:3, :3] = Rotation.from_euler("xyz", [ai, aj, ak]).as_matrix()
    return M


def reference_namespace():
    from make_golden_geometry import extract

    ns = {"np": np, "cv2": cv2, "torch": torch, "logging": logging, "OrderedDict": OrderedDict, "defaultdict": defaultdict,
          "euler_matrix": euler_matrix}
    for name in NAMES:
        exec(extract(os.path.join(REF, "Utils.py"), name), ns)
    return ns


def main():
    import shim_cases

    real_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self  # depth2xyzmap_batch hard-codes .cuda(); run it where we are
    if not torch.cuda.is_available():
        torch.cuda.manual_seed_all = lambda *a, **k: None  # set_seed: a no-op without a device anyway
    want = shim_cases.collect(reference_namespace())
    torch.Tensor.cuda = real_cuda
    dst = os.path.join(ROOT, "tests", "golden", "shim_golden.npz")
    np.savez_compressed(dst, **want)
    print(f"wrote {dst}: {len(want)} entries, {os.path.getsize(dst) / 1024:.0f} KiB")
    sys.path.insert(0, os.path.join(ROOT, "foundationpose_b200", "dropin"))
    import Utils as shim

    got = shim_cases.collect(shim)
    for k in want:
        a, b = np.asarray(got[k]), np.asarray(want[k])
        if a.shape != b.shape:
            print(f"  {k}: shape {a.shape} vs {b.shape}")
        elif a.dtype.kind in "US":
            if str(a) != str(b):
                print(f"  {k}: text differs")
        elif not np.allclose(a.astype(np.float64), b.astype(np.float64), rtol=1e-12, atol=1e-12):
            d = np.abs(a.astype(np.float64) - b.astype(np.float64))
            print(f"  {k}: max |diff| {d.max():.3g} at {int((d > 0).sum())} of {d.size} entries")
    print("compared")


if __name__ == "__main__":
    main()
