"""Where does the host-side time of FoundationPose.register() go?  Phase timings (wall clock, a device synchronise after
every phase: diagnostic only) and the un-instrumented end-to-end time, for the bench scene."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from foundationpose_b200 import synth  # noqa: E402
from foundationpose_b200.engine import Engine  # noqa: E402
from foundationpose_b200.estimater import FoundationPose, PoseRefinePredictor, ScorePredictor  # noqa: E402
from foundationpose_b200.weights import random_state_dict  # noqa: E402


def main():
    mesh, gt, K, rgb, depth, mask = synth.default_scene(5, 0)
    e = Engine()
    est = FoundationPose(model_pts=mesh.vertices, model_normals=mesh.vertex_normals, mesh=mesh,
                         scorer=ScorePredictor(engine=e, state_dict=random_state_dict("score", 0)),
                         refiner=PoseRefinePredictor(engine=e, state_dict=random_state_dict("refine", 0)))
    for _ in range(4):
        est.register(K=K, rgb=rgb, depth=depth, ob_mask=mask, iteration=5)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        est.register(K=K, rgb=rgb, depth=depth, ob_mask=mask, iteration=5)
    print(f"register() end to end: {(time.perf_counter() - t0) / 10 * 1e3:.3f} ms")
    ph = {}

    def tick(name, t):
        torch.cuda.synchronize()
        ph[name] = ph.get(name, 0.0) + (time.perf_counter() - t)

    for _ in range(10):
        t = time.perf_counter(); e.set_frame(rgb, depth, K, filter_depth=True); t1 = time.perf_counter(); ph["set_frame call"] = ph.get("set_frame call", 0) + t1 - t; tick("set_frame", t)
        t = time.perf_counter(); poses, info = e.start_poses(mask, est.rot_grid); tick("start_poses", t)
        t = time.perf_counter(); p, lt, lr = e.refine(poses, 5); t1 = time.perf_counter(); ph["refine call"] = ph.get("refine call", 0) + t1 - t; tick("refine", t)
        t = time.perf_counter(); s, b = e.score(p); tick("score", t)
        t = time.perf_counter(); ids = s.argsort(descending=True); out = torch.cat([info, (p[ids][0] @ est.get_tf_to_centered_mesh()).reshape(-1)]).cpu().numpy(); tick("rank + read-back", t)
    for k, v in ph.items():
        print(f"  {k:18s} {v / 10 * 1e3:8.3f} ms")


if __name__ == "__main__":
    main()
