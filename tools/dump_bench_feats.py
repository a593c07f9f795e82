"""Dump the scorer features of the bench scene's 252 refined hypotheses (native path) so that the seeded scorer tail can
be chosen with a robust top-2 margin on the bench scene as well as on the golden scene (tools/pick_tail_seed.py)."""
import os
import sys

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
    e.set_frame(rgb, depth, K, filter_depth=True)
    poses, info = e.start_poses(mask, est.rot_grid)
    p, _, _ = e.refine(poses, 5)
    f = e.score_features(p)
    s, b = e.score(p)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez(os.path.join(ROOT, "gpurun_out", "bench_feats.npz"), feats=f.cpu().numpy(), poses=p.cpu().numpy(), scores=s.cpu().numpy())
    ss = np.sort(s.cpu().numpy())
    print("best", int(b.item()), "margin", ss[-1] - ss[-2], "spread", ss.std())


if __name__ == "__main__":
    main()
