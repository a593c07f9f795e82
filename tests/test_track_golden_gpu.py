"""track_one (estimater.py:250-268) against the CPU oracle on the 50-frame synthetic sequence of
tools/make_golden_track.py (tests/golden/track_seq.npz): every frame through fp_track — ONE graph launch: upload,
erode + bilateral, xyz map, two refiner passes, read-back — SE(3) delta within 1e-3 per frame."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "track_seq.npz")


@pytest.fixture(scope="module")
def rig():
    from foundationpose_b200 import synth
    from foundationpose_b200.engine import Engine
    from foundationpose_b200.estimater import make_mesh_tensors
    from foundationpose_b200.weights import random_state_dict

    g = dict(np.load(GOLD))
    mesh = synth.make_mesh(3)
    mt = make_mesh_tensors(mesh)
    e = Engine()
    e.load_network("refine", random_state_dict("refine", 0))
    e.set_mesh(mt["pos"], mt["normals"], mt["faces"], synth.mesh_diameter(mesh.vertices), uv=mt["uv"], tex=mt["tex"])
    return e, g, mesh


def _frame(mesh, g, i):
    from foundationpose_b200 import synth

    rgb, depth, _ = synth.make_scene(mesh.visual.image, g["gt"][i], seed=1 + i)
    return rgb, depth


def test_every_frame_within_1e3(rig):
    from foundationpose_b200 import synth

    e, g, mesh = rig
    worst_t = worst_r = 0.0
    for k in range(len(g["pose_in"])):
        rgb, depth = _frame(mesh, g, k + 1)
        pin = torch.from_numpy(g["pose_in"][k]).cuda()
        pose_dev, pose_host = e.track(rgb, depth, synth.DEFAULT_K, pin, 2)
        assert torch.equal(pose_dev.cpu(), torch.from_numpy(pose_host)), "device and host copies of the pose differ"
        err = np.abs(pose_host - g["pose_out"][k])
        worst_t, worst_r = max(worst_t, err[:3, 3].max()), max(worst_r, err[:3, :3].max())
        assert err.max() <= 1e-3, f"frame {k + 1}: pose off by {err.max():.2e}"
    print(f"track_one over {len(g['pose_in'])} frames: worst translation error {worst_t:.2e} m, rotation {worst_r:.2e}")


def test_pose_fed_back_over_frames(rig):
    """fp_track with pose_in = NULL continues from the pose the context produced last (the tracker's feedback)."""
    from foundationpose_b200 import synth

    e, g, mesh = rig
    chain = g["chain"]
    pose = torch.from_numpy(chain[0]).cuda()
    for i in range(1, len(chain)):
        rgb, depth = _frame(mesh, g, i)
        _, host = e.track(rgb, depth, synth.DEFAULT_K, pose if i == 1 else None, 2)
        err = np.abs(host - chain[i]).max()
        assert err <= 2e-3 * i, f"chained frame {i}: {err:.2e}"


def test_track_one_api_uses_the_graph_path(rig):
    """FoundationPose.track_one with host frames == the staged path (set_frame + predict) on the same inputs."""
    from foundationpose_b200 import synth
    from foundationpose_b200.estimater import FoundationPose, PoseRefinePredictor, ScorePredictor
    from foundationpose_b200.weights import random_state_dict

    e, g, mesh = rig
    refiner = PoseRefinePredictor(engine=e, state_dict=random_state_dict("refine", 0))
    scorer = ScorePredictor(engine=e, state_dict=random_state_dict("score", 0))
    est = FoundationPose(model_pts=mesh.vertices, model_normals=mesh.vertex_normals, mesh=mesh, scorer=scorer, refiner=refiner)
    rgb, depth = _frame(mesh, g, 1)
    est.pose_last = torch.from_numpy(g["pose_in"][0]).cuda().reshape(1, 4, 4)
    a = est.track_one(rgb=rgb, depth=depth, K=synth.DEFAULT_K, iteration=2)
    est.pose_last = torch.from_numpy(g["pose_in"][0]).cuda().reshape(1, 4, 4)
    b = est.track_one(rgb=torch.from_numpy(rgb).cuda(), depth=torch.from_numpy(depth).cuda(), K=synth.DEFAULT_K, iteration=2)
    np.testing.assert_allclose(a, b, atol=1e-6, rtol=0)
    np.testing.assert_allclose(a, g["pose_out"][0] @ np.linalg.inv(np.eye(4)), atol=1e-3, rtol=0)  # model_center = 0
