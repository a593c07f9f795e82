"""Per-rank cost of the sharded register at the shard sizes of N = 1, 2, 4, 8 GPUs, measured on ONE GPU:
refine (5 iterations) + score features on ceil(252 / N) hypotheses through the graphed engine calls, CUDA
events, 3 warm-up + 10 timed.  Shows how much of the strong-scaling loss is per-rank (tile quantisation,
launch ramps, fixed-cost kernels) rather than communication.

    python tools/shard_probe.py            # prints one line per shard size
    FP_PROF=1 python tools/shard_probe.py  # adds the share of the GEMM kernels (per-launch events)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from foundationpose_b200 import _lib, hypotheses, synth  # noqa: E402
from foundationpose_b200.estimater import FoundationPose, PoseRefinePredictor, ScorePredictor  # noqa: E402
from foundationpose_b200.weights import random_state_dict  # noqa: E402


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
    e.set_frame(rgb, depth, K, filter_depth=True)
    poses_all, _ = e.start_poses(mask, est.rot_grid)
    base = None
    for world in [int(w) for w in os.environ.get("FP_WORLDS", "1,2,4,8").split(",")]:
        n = (252 + world - 1) // world
        poses = poses_all[:n].contiguous()
        for _ in range(3):
            out, _, _ = e.refine(poses, 5)
            e.score_features(out)
        if os.environ.get("FP_PROF"):
            _lib.prof_enable(True)
            out, _, _ = e.refine(poses, 5)
            e.score_features(out)
            g_ms, g_work, g_n = _lib.prof_collect(0)
            _lib.prof_enable(False)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            out, _, _ = e.refine(poses, 5)
            e.score_features(out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        if base is None:
            base = ms / n
        line = f"[shard] world {world}: {n:3d} hyp/rank  {ms:7.3f} ms/rank-step  {ms / n * 1e3:7.1f} us/hyp  per-rank efficiency {base / (ms / n) * 100:5.1f} %"
        if os.environ.get("FP_PROF"):
            line += f"  | gemm kernels (ungraphed pass) {g_ms:6.3f} ms in {g_n} launches, {g_work / g_ms / 1e9:6.0f} TFLOP/s"
        print(line, flush=True)


if __name__ == "__main__":
    main()
