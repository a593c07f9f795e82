"""CPU: the oracle's register / track_one flow (oracle/pipeline.py — what the GPU parity tests compare the CUDA path with)
against golden vectors produced by the REFERENCE's own unmodified method sources (tools/make_golden_flow.py):
FoundationPose.register / track_one / generate_random_pose_hypo / guess_translation / get_tf_to_centered_mesh
(estimater.py:77-268), PoseRefinePredictor.predict (predict_pose_refine.py:148-239), ScorePredictor.predict
(predict_score.py:160-214) driving the reference's own network classes — with only the absent third-party pieces
substituted: nvdiffrast_render -> oracle rasteriser, kornia warp_perspective -> oracle warp, so3_exp_map, the Warp depth
filters (see the generator's header).  Both `make_crop_data_batch` functions and the dataset transforms are the
reference's code, so this also pins the glue of oracle.pipeline.make_crops.

Bars.  Poses: 2e-5 (measured 6e-6).  What separates the two sides is implementation-defined in the reference itself: its
render window goes through a fp32 LU `tf.inverse()` (predict_pose_refine.py:45) where the oracle — and the CUDA kernel —
use the closed form, which moves ~100 of 76 800 rendered values per pose by up to 1e-4 m / 0.02 grey levels; through the
stand-in scorer's x60 read-out that is a common score offset of 0.07 (bar 0.15) and 0.004 between hypotheses (bar 0.02,
against a top-2 margin of 0.94).  The scorer's crop -> full-resolution warp has a genuine rounding tie on the first crop
column (oracle.geometry.unwarp_nearest): the fixture resolves it exactly, and also records the scores when kornia's op
sequence decides it on the generating machine — same ranking, offset 2.4."""
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def setup():
    import sys

    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_golden_flow as gen  # scene(), start_grid() and the iteration counts only: nothing of the reference is read

    from foundationpose_b200.weights import random_state_dict

    torch.set_num_threads(os.cpu_count())
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "flow_golden.npz")))
    return gen, g, gen.scene(), random_state_dict("refine", 0), random_state_dict("score", 0)


def test_register_flow_matches_the_reference_methods(setup):
    from oracle import pipeline

    gen, g, (mesh, mt, gt, rgb, depth, mask, K, d, center), sd_r, sd_s = setup
    r = pipeline.register(sd_r, sd_s, gen.start_grid(), mt, rgb, depth, mask, K, d, center, iterations=gen.REGISTER_ITERS)
    assert not r["early"]
    assert int(r["ids"][0]) == int(g["reg_best_id"])
    diff = r["scores"].numpy() - g["reg_scores_sorted"]
    assert np.abs(diff).max() < 0.15 and np.abs(diff - diff.mean()).max() < 0.02, diff
    margin = g["reg_scores_sorted"][0] - g["reg_scores_sorted"][1]
    assert margin > 0.5, "the golden scene must have a clear winner for the ranking check to mean something"
    assert np.abs(r["poses"].numpy() - g["reg_poses_sorted"]).max() < 2e-5       # same ORDER, same refined poses
    assert np.abs(r["pose_last"].numpy() - g["reg_pose_last"]).max() < 2e-5
    assert np.abs(r["pose"] - g["reg_best_pose"]).max() < 2e-5                  # incl. the @ T(-model_center)
    assert np.abs(r["last_trans"].numpy() - g["reg_last_trans"]).max() < 1e-5   # deltas of the LAST iteration, input order
    assert np.abs(r["last_rot"].numpy() - g["reg_last_rot"]).max() < 2e-5
    # sensitivity to the implementation-defined tie of the scorer's inverse warp: ranking unchanged
    assert int(g["opseq_best_id"]) == int(g["reg_best_id"])
    d2 = g["opseq_scores_sorted"] - g["reg_scores_sorted"]
    assert np.abs(d2).max() < 3.0 and np.abs(d2 - d2.mean()).max() < 0.05, d2


def test_early_out_matches_the_reference(setup):
    from oracle import pipeline

    gen, g, (mesh, mt, gt, rgb, depth, mask, K, d, center), sd_r, sd_s = setup
    tiny = np.zeros_like(mask)
    ys, xs = np.nonzero(mask)
    tiny[ys[:3], xs[:3]] = True
    for m, key in ((tiny, "early_out_3px"), (np.zeros_like(mask), "early_out_empty")):
        r = pipeline.register(sd_r, sd_s, gen.start_grid(), mt, rgb, depth, m, K, d, center, iterations=gen.REGISTER_ITERS)
        assert r["early"] and np.abs(r["pose"] - g[key]).max() < 1e-9
    assert np.array_equal(g["early_out_empty"], np.eye(4))
    assert bool(g["track_before_register_raises"])


def test_track_one_flow_matches_the_reference_methods(setup):
    from foundationpose_b200 import synth
    from oracle import pipeline

    gen, g, (mesh, mt, gt, rgb, depth, mask, K, d, center), sd_r, sd_s = setup
    seq = synth.track_sequence(3, gt)
    assert np.allclose(np.stack(seq[1:]), g["track_gt"])
    pose_last = g["reg_pose_last"]
    for i, p in enumerate(seq[1:]):
        rgb_i, depth_i, _ = synth.make_scene(mesh.visual.image, p, seed=11 + i)
        out, pose_last, lt = pipeline.track_one(sd_r, pose_last, mt, rgb_i, depth_i, K, d, center, iterations=gen.TRACK_ITERS)
        assert np.abs(out - g[f"track_pose{i}"]).max() < 2e-5
        assert np.abs(pose_last.numpy().reshape(4, 4) - g[f"track_pose_last{i}"].reshape(4, 4)).max() < 2e-5
        assert np.abs(lt.numpy() - g[f"track_last_trans{i}"]).max() < 1e-5
        pose_last = pose_last.numpy()


def test_crop_glue_matches_the_reference_make_crop_data_batch(setup):
    """oracle.pipeline.make_crops vs the reference's two `make_crop_data_batch` + `transform_batch` (run on the same two
    substituted primitives).  Observed crops B — colour and xyz, refiner and scorer incl. the depth round trip — are
    BIT-IDENTICAL (sha1).  Rendered crops A: four of the five poses bit-identical too; the fifth (the grid's first,
    axis-aligned pose, where the symmetric mesh puts many vertices on sub-pixel snapping ties) differs at 766 of its
    6 624 covered pixels by <= 1.3e-4 m / 0.03 grey levels, because the reference's render window comes out of a fp32 LU
    `tf.inverse()` and differs from the closed form in the last bit (see the module docstring).  Bar: < 1 % of the values,
    none by more than 0.05."""
    import hashlib

    from oracle import geometry, pipeline

    gen, g, (mesh, mt, gt, rgb, depth, mask, K, d, center), sd_r, sd_s = setup
    sha = lambda t: np.frombuffer(hashlib.sha1(np.ascontiguousarray(t.numpy()).tobytes()).digest(), dtype=np.uint8)
    depth_f = geometry.bilateral_filter_depth(geometry.erode_depth(depth))
    poses = g["crop_poses"]
    assert np.allclose(poses[:, :3, 3], geometry.guess_translation(depth_f, mask, K))
    xyz_map = geometry.depth2xyzmap(depth_f, K)
    A0, B0, _ = pipeline.make_crops(poses, mt, rgb, depth_f, xyz_map, K, d, 0)
    A1, B1, _ = pipeline.make_crops(poses, mt, rgb, depth_f, None, K, d, 1)
    assert np.array_equal(sha(B0[:, :3].contiguous()), g["crop_refine_rgbB_sha1"])
    assert np.array_equal(sha(B0[:, 3:].contiguous()), g["crop_refine_xyzB_sha1"])
    assert np.array_equal(sha(B1[:, :3].contiguous()), g["crop_score_rgbB_sha1"])
    assert np.array_equal(sha(B1[:, 3:].contiguous()), g["crop_score_xyzB_sha1"])
    for mine, want, what in ((A0.numpy(), g["crop_refine_A"], "refiner A"), (A1[:, 3:].numpy(), g["crop_score_xyzA"], "scorer xyz A")):
        diff = np.abs(mine - want)
        frac = float((diff > 1e-6).mean())
        print(f"{what}: {int((diff > 1e-6).sum())} of {diff.size} values differ ({100 * frac:.3f} %), max {diff.max():.3g}")
        assert frac < 1e-2 and diff.max() < 0.05, what


def test_headline_golden_equals_the_reference_register():
    """The committed 252 x 5 oracle golden that the GPU test holds the CUDA path to (register_252x5.npz) vs the reference's
    own `register` executed over the same configuration (`tools/make_golden_flow.py --headline`, 283 s on 8 cores): same
    selected hypothesis, same top ten, no pair of hypotheses more than 0.04 apart ranked differently, refined poses equal
    to 2.4e-5 for 99 % of the hypotheses (one at 1.8e-4 after five free-running iterations: the LU-inverse render window,
    see the module docstring)."""
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "register_252x5.npz")))
    r = dict(np.load(os.path.join(ROOT, "tests", "golden", "register_252x5_reference_flow.npz")))
    assert int(r["best_id"]) == int(g["best"][0]) == 13
    dp = np.abs(r["poses"] - g["poses"][5]).reshape(252, -1).max(1)
    assert np.quantile(dp, 0.99) < 5e-5 and dp.max() < 5e-4, (np.quantile(dp, 0.99), dp.max())
    assert np.abs(r["last_trans"] - g["last_trans"][4]).max() < 5e-5 and np.abs(r["last_rot"] - g["last_rot"][4]).max() < 1e-4
    ds = r["scores"] - g["scores"]
    assert abs(ds.mean()) < 0.02 and np.abs(ds - ds.mean()).max() < 0.03, (ds.mean(), np.abs(ds - ds.mean()).max())
    ref_ids = np.argsort(-r["scores"], kind="stable")
    assert np.array_equal(ref_ids[:10], g["ids"][:10])
    gap = g["scores"][:, None] - g["scores"][None]
    flipped = (np.abs(gap) > 0.04) & (np.sign(gap) != np.sign(r["scores"][:, None] - r["scores"][None]))
    assert not flipped.any()
    margin = np.sort(r["scores"])[-1] - np.sort(r["scores"])[-2]
    assert abs(margin - float(g["top2_margin"][0])) < 5e-3 and margin > 0.09
    # the returned pose = best refined pose @ T(-model_center) with model_center = 0 here
    assert np.abs(r["best_pose"] - r["poses"][13]).max() < 1e-6


def test_track_golden_equals_the_reference_track_one():
    """The committed 49-frame oracle golden of the GPU track test (track_seq.npz) vs the reference's own `track_one`
    executed over the same frames and over the 5-frame fed-back chain (`tools/make_golden_flow.py --track`): 2e-7."""
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "track_seq.npz")))
    r = dict(np.load(os.path.join(ROOT, "tests", "golden", "track_seq_reference_flow.npz")))
    assert r["pose_out"].shape == g["pose_out"].shape == (49, 4, 4)
    assert np.abs(r["pose_out"] - g["pose_out"]).max() < 5e-6
    assert np.abs(r["last_trans"] - g["last_trans"]).max() < 2e-6
    assert np.abs(r["chain"] - g["chain"]).max() < 5e-6
