"""Device time of the per-frame fixed costs of register(): set_frame (H2D + depth filters + xyz), start_poses,
scorer tail on 252 feature rows.  CUDA events, 5 warm-up + 20 timed each."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from foundationpose_b200 import synth  # noqa: E402
from foundationpose_b200.estimater import FoundationPose, PoseRefinePredictor, ScorePredictor  # noqa: E402
from foundationpose_b200.weights import random_state_dict  # noqa: E402


def timed(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    mesh = synth.make_mesh(5)
    pose = np.eye(4)
    pose[:3, :3] = synth.random_rotation(0)
    pose[:3, 3] = [0.0, 0.0, 0.6]
    rgb, depth, mask = synth.make_scene(mesh.visual.image, pose, seed=1)
    K = synth.DEFAULT_K
    refiner = PoseRefinePredictor(state_dict=random_state_dict("refine", 0))
    scorer = ScorePredictor(engine=refiner.engine, state_dict=random_state_dict("score", 0))
    est = FoundationPose(model_pts=mesh.vertices, model_normals=mesh.vertex_normals, mesh=mesh, scorer=scorer, refiner=refiner)
    e = est.engine
    print(f"set_frame(filter_depth=True)  {timed(lambda: e.set_frame(rgb, depth, K, filter_depth=True)):.3f} ms")
    print(f"start_poses                   {timed(lambda: e.start_poses(mask, est.rot_grid)):.3f} ms")
    poses, _ = e.start_poses(mask, est.rot_grid)
    feats = e.score_features(poses)
    print(f"score_tail (252 rows)         {timed(lambda: e.score_tail(feats)):.3f} ms")
    print(f"score_features (252)          {timed(lambda: e.score_features(poses)):.3f} ms")
    p32 = poses[:32].contiguous()
    print(f"score_features (32)           {timed(lambda: e.score_features(p32)):.3f} ms")
    print(f"refine x5 (32)                {timed(lambda: e.refine(p32, 5)):.3f} ms")


if __name__ == "__main__":
    main()
