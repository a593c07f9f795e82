"""How many pixels does the oracle rasteriser's fixed-point rule (vertices snapped to 1/256 px, integer edge functions,
tie bias) move relative to an UN-SNAPPED rasteriser with nvdiffrast's published semantics, over the 252 start poses of the
headline golden scene?  Both go through the reference's own `nvdiffrast_render` conventions: the left side is
oracle.raster.render_crop, the right side the reference's unmodified function on tools/nvdiffrast_semantics.py
(tools/make_golden_render.py machinery).  Needs /root/reference.

    python tools/silhouette_stats.py [n_poses]      # prints the statistics recorded in DESIGN.md §4
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    import logging

    import make_golden_register as mgr
    import make_golden_render as mg
    import nvdiffrast_semantics as dr
    from make_golden_flow import _TorchProxy
    from make_golden_geometry import extract, load_reference_functions
    from oracle import geometry, raster

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 252
    torch.Tensor.cuda = lambda self, *a, **k: self
    ref = load_reference_functions()
    ns = {"np": np, "torch": _TorchProxy("torch"), "logging": logging, "F": torch.nn.functional, "dr": dr,
          "make_mesh_tensors": ref["make_mesh_tensors"], "projection_matrix_from_intrinsics": ref["projection_matrix_from_intrinsics"]}
    exec(mg.extract_assign(os.path.join(mg.REF, "Utils.py"), "glcam_in_cvcam"), ns)
    for name in ("nvdiffrast_render", "transform_pts", "transform_dirs", "to_homo_torch"):
        exec(extract(os.path.join(mg.REF, "Utils.py"), name), ns)
    mesh, mt, gt, rgb, depth, mask, K, d = mgr.scene()
    gold = dict(np.load(os.path.join(ROOT, "tests", "golden", "register_252x5.npz")))
    poses = gold["start"][:: max(252 // n, 1)][:n].astype(np.float32)
    win, _ = geometry.crop_window(poses, K, d)
    umin, vmin, umax, vmax = geometry.render_window(win)
    boxes = np.stack([umin, vmin, umax, vmax], 1).astype(np.float32)
    mt_ref = mg.reference_mesh_tensors(ref, mesh, None)
    tot_cov = mism = 0
    dx_all, dc_all = [], []
    for i, (pose, box) in enumerate(zip(poses, boxes)):
        o_rgb, o_xyz, _ = raster.render_crop(pose, mt, K, tuple(box))
        extra = {}
        color, _, _ = ns["nvdiffrast_render"](K=K, H=480, W=640, ob_in_cams=torch.from_numpy(pose)[None], context="cuda", get_normal=False, glctx="ctx",
                                              mesh_tensors=mt_ref, output_size=(160, 160), bbox2d=torch.from_numpy(box)[None], use_light=True, extra=extra)
        r_xyz, r_rgb = extra["xyz_map"][0].numpy(), color[0].numpy()
        co, cr = o_xyz[..., 2] > 0, r_xyz[..., 2] > 0
        mism += int((co != cr).sum())
        tot_cov += int((co | cr).sum())
        both = co & cr
        dx_all.append(np.abs(o_xyz - r_xyz)[both].max(-1))
        dc_all.append(np.abs(o_rgb - r_rgb)[both].max(-1))
        if (i + 1) % 50 == 0:
            print(f"{i + 1} poses ...", flush=True)
    dx, dc = np.concatenate(dx_all), np.concatenate(dc_all)
    print(f"{len(poses)} poses, {tot_cov} covered pixels: coverage differs at {mism} pixels ({100.0 * mism / tot_cov:.4f} %)")
    print(f"xyz |diff| at commonly covered pixels: median {np.median(dx):.2e}, 99 % {np.quantile(dx, 0.99):.2e}, 99.9 % {np.quantile(dx, 0.999):.2e}, max {dx.max():.2e}")
    print(f"rgb |diff|: median {np.median(dc):.2e}, 99 % {np.quantile(dc, 0.99):.2e}, 99.9 % {np.quantile(dc, 0.999):.2e}, max {dc.max():.2e}")


if __name__ == "__main__":
    main()
