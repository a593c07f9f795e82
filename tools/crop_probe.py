"""On-GPU probe of the crop producer (csrc/fp_crop.cu): device time per pass at the C2 batch (252 hypotheses) and
at N = 1 / 32, work counters, and achieved fraction of the HBM roofline (algorithmic bytes = the two fp16 6-channel
crops per hypothesis)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from foundationpose_b200 import hypotheses, synth  # noqa: E402
from foundationpose_b200.engine import Engine  # noqa: E402
from foundationpose_b200.estimater import make_mesh_tensors  # noqa: E402


def main():
    sub = int(os.environ.get("PROBE_SUBDIV", "5"))
    mesh, gt, K, rgb, depth, mask = synth.default_scene(subdivisions=sub, seed=0)
    mt = make_mesh_tensors(mesh)
    d = synth.mesh_diameter(mesh.vertices)
    e = Engine()
    e.set_mesh(mt["pos"], mt["normals"], mt["faces"], d, uv=mt["uv"], tex=mt["tex"])
    e.set_frame(rgb, depth, K, filter_depth=True)
    print("mesh", e.mesh_info())
    grid = hypotheses.make_rotation_grid()
    center = hypotheses.guess_translation(e.get_depth()[0].cpu().numpy(), mask, K)
    poses = grid.copy().astype(np.float32)
    poses[:, :3, 3] = center
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    tiles = [int(t) for t in os.environ.get("PROBE_TILES", "0").split(",")]
    for n, tile in [(n, t) for n in (252, 32, 1) for t in tiles]:
        p = torch.from_numpy(poses[:n]).cuda()
        e.set_crop_tile(tile)
        print(f"--- N={n} tile={tile or 'auto'}")
        for mode in (0, 1):
            for _ in range(5):
                e.make_crops(p, mode=mode, want_crops=False)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 50
            e0.record()
            for _ in range(reps):
                e.make_crops(p, mode=mode, want_crops=False)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / reps * 1e3
            gbs = n * 2 * 6 * 160 * 160 * 2 / (us * 1e-6) / 1e9
            print(f"[crop] N={n:3d} mode={mode}: {us:8.1f} us/pass  {gbs:7.1f} GB/s algorithmic = {gbs / peak:.3f} of the measured HBM peak ({peak:.0f} GB/s)")
        st = e.crop_stats(p, 0)
        print(f"[crop] N={n:3d} stats per hypothesis: " + ", ".join(f"{k} {v / n:.0f}" for k, v in st.items()))


if __name__ == "__main__":
    main()
