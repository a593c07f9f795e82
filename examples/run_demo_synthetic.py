"""The flow of the reference's run_demo.py (run_demo.py:38-63) on this package, with a synthetic sequence instead of
`demo_data/mustard0` (no dataset offline): register on the first frame, track on the following ones, write one
ob_in_cam/<frame>.txt per frame.  Only the import line and the data source differ from the reference driver:

    from foundationpose_b200.estimater import FoundationPose, PoseRefinePredictor, ScorePredictor

    python examples/run_demo_synthetic.py --frames 30 --out_dir /tmp/fpose_demo
"""
import argparse
import logging
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from foundationpose_b200 import synth  # noqa: E402
from foundationpose_b200.estimater import FoundationPose, PoseRefinePredictor, ScorePredictor  # noqa: E402


def rot_err_deg(Ra, Rb):
    c = (np.trace(Ra.T @ Rb) - 1) / 2
    return float(np.degrees(np.arccos(np.clip(c, -1, 1))))


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("--frames", type=int, default=30)
    parser.add_argument("--est_refine_iter", type=int, default=5)
    parser.add_argument("--track_refine_iter", type=int, default=2)
    parser.add_argument("--out_dir", type=str, default="/tmp/fpose_b200_demo")
    args = parser.parse_args()
    logging.basicConfig(level=logging.INFO, format="%(message)s")
    os.makedirs(f"{args.out_dir}/ob_in_cam", exist_ok=True)

    mesh = synth.make_mesh(5)
    K = synth.DEFAULT_K
    scorer = ScorePredictor()      # no checkpoints offline: seeded synthetic weights (see weights.py)
    refiner = PoseRefinePredictor()
    est = FoundationPose(model_pts=mesh.vertices, model_normals=mesh.vertex_normals, mesh=mesh, scorer=scorer, refiner=refiner)
    logging.info("estimator initialization done")

    rng = np.random.default_rng(0)
    gt = np.eye(4)
    gt[:3, :3] = synth.random_rotation(0)
    gt[:3, 3] = [0.0, 0.0, 0.6]
    for i in range(args.frames):
        color, depth, mask = synth.make_scene(mesh.visual.image, gt, seed=1 + i)
        if i == 0:
            pose = est.register(K=K, rgb=color, depth=depth, ob_mask=mask.astype(bool), iteration=args.est_refine_iter)
        else:
            pose = est.track_one(rgb=color, depth=depth, K=K, iteration=args.track_refine_iter)
        np.savetxt(f"{args.out_dir}/ob_in_cam/{i:06d}.txt", pose.reshape(4, 4))
        logging.info(f"i:{i}  |t - t_gt| = {np.linalg.norm(pose[:3, 3] - gt[:3, 3]) * 1e3:7.2f} mm   "
                     f"rot err = {rot_err_deg(pose[:3, :3], gt[:3, :3]):6.2f} deg   (random-init networks: not a pose-accuracy claim)")
        # the object drifts a little between frames
        ang = np.deg2rad(1.5) * rng.uniform(-1, 1)
        c, s = np.cos(ang), np.sin(ang)
        gt = gt.copy()
        gt[:3, :3] = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]]) @ gt[:3, :3]
        gt[:3, 3] += rng.uniform(-0.003, 0.003, 3)
    logging.info(f"poses written to {args.out_dir}/ob_in_cam")
