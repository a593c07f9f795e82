"""GPU parity of the end-to-end hot loop (fp_refine / fp_score / FoundationPose.register) against the
CPU oracle on the same seeded scene, mesh, weights and start poses.

Bars (BASELINE.json north_star): predicted SE(3) delta within 1e-3 (translation in metres, rotation
matrix entries), selected hypothesis index identical.  At the full 252-hypothesis size the oracle
is too slow for CI, so size-independent properties are checked instead (determinism, sharding
invariance, arg-max consistency).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    from foundationpose_b200 import synth
    from foundationpose_b200.engine import Engine
    from foundationpose_b200.weights import random_state_dict
    from oracle import pipeline

    mesh = synth.make_mesh(3)
    pose = np.eye(4)
    pose[:3, :3] = synth.random_rotation(0)
    pose[:3, 3] = [0.02, -0.01, 0.6]
    rgb, depth, mask = synth.make_scene(mesh.visual.image, pose)
    d = synth.mesh_diameter(mesh.vertices)
    mt = pipeline.mesh_tensors(mesh)
    sd_r, sd_s = random_state_dict("refine", 0), random_state_dict("score", 0)
    e = Engine()
    e.load_network("refine", sd_r)
    e.load_network("score", sd_s)
    e.set_mesh(mt["pos"], mt["normals"], mt["faces"], d, uv=mt["uv"], tex=mt["tex"])
    e.set_frame(rgb, depth, synth.DEFAULT_K, filter_depth=False)
    rng = np.random.default_rng(5)
    poses = np.tile(pose[None], (6, 1, 1)).astype(np.float32)
    for i in range(1, 6):
        poses[i, :3, :3] = synth.random_rotation(10 + i) if i % 2 else poses[i, :3, :3]
        poses[i, :3, 3] += rng.normal(0, 0.01, 3)
    return dict(e=e, mesh=mesh, mt=mt, rgb=rgb, depth=depth, mask=mask, K=synth.DEFAULT_K, d=d, poses=poses, sd_r=sd_r, sd_s=sd_s, gt=pose)


def test_refine_one_iteration_matches_oracle(setup):
    from oracle import pipeline

    s = setup
    poses = s["poses"][:4]
    out, lt, lr = s["e"].refine(poses, 1)
    ref, rt, rr = pipeline.refine(s["sd_r"], poses, s["mt"], s["rgb"], s["depth"], s["K"], s["d"], 1)
    np.testing.assert_allclose(lt.cpu().numpy(), rt.numpy(), atol=1e-3, rtol=0)  # metres
    np.testing.assert_allclose(lr.cpu().numpy(), rr.numpy(), atol=1e-3, rtol=0)
    np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), atol=1e-3, rtol=0)


def test_refine_two_iterations_matches_oracle(setup):
    from oracle import pipeline

    s = setup
    poses = s["poses"][:3]
    out, lt, lr = s["e"].refine(poses, 2)
    ref, rt, rr = pipeline.refine(s["sd_r"], poses, s["mt"], s["rgb"], s["depth"], s["K"], s["d"], 2)
    np.testing.assert_allclose(lt.cpu().numpy(), rt.numpy(), atol=1e-3, rtol=0)
    np.testing.assert_allclose(lr.cpu().numpy(), rr.numpy(), atol=1e-3, rtol=0)
    np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), atol=2e-3, rtol=0)


def test_score_matches_oracle(setup):
    from oracle import pipeline

    s = setup
    scores, best = s["e"].score(s["poses"])
    ref_scores, ref_best = pipeline.score(s["sd_s"], s["poses"], s["mt"], s["rgb"], s["depth"], s["K"], s["d"])
    got = scores.cpu().numpy()
    ref = ref_scores.numpy()
    print("scores", got, ref)
    spread = float(ref.std())
    err = got - ref
    rank_err = np.abs(err - err.mean()).max()  # a common offset cannot change the ranking
    top2 = np.sort(ref)[-2:]
    print(f"score spread {spread:.3f}, max err {np.abs(err).max():.2e}, rank-relevant err {rank_err:.2e}, top-2 margin {top2[1] - top2[0]:.4f}")
    assert rank_err <= 0.25 * spread, f"score error {rank_err:.3g} vs spread {spread:.3g}"
    # six arbitrary poses: the winner is only defined when the oracle separates the two leaders by more than the
    # error; the unconditional index test is the 252-hypothesis golden (tests/test_register_golden_gpu.py)
    if top2[1] - top2[0] > 4 * rank_err:
        assert int(best.item()) == ref_best
    assert int(best.item()) == int(np.argmax(got))


def test_register_properties_full_size(setup):
    """252 hypotheses x 5 iterations: determinism, sharding invariance, arg-max consistency."""
    from foundationpose_b200 import hypotheses

    s, e = setup, setup["e"]
    grid = hypotheses.make_rotation_grid()
    assert grid.shape == (252, 4, 4)
    depth = s["depth"]
    center = hypotheses.guess_translation(depth, s["mask"], s["K"])
    poses = grid.copy()
    poses[:, :3, 3] = center
    p1, _, _ = e.refine(poses, 5)
    p2, _, _ = e.refine(poses, 5)
    assert torch.equal(p1, p2), "refine is not deterministic"
    sc1, b1 = e.score(p1)
    sc2, b2 = e.score(p1)
    assert torch.equal(sc1, sc2) and int(b1) == int(b2)
    assert int(b1) == int(sc1.argmax())
    # sharded: two halves refined / featurised independently, then one tail on the gathered features
    h = 126
    pa, _, _ = e.refine(poses[:h], 5)
    pb, _, _ = e.refine(poses[h:], 5)
    assert torch.equal(torch.cat([pa, pb]), p1), "refinement must be per-hypothesis independent"
    fa = e.score_features(pa).clone()
    fb = e.score_features(pb).clone()
    sc3, b3 = e.score_tail(torch.cat([fa, fb]))
    assert torch.equal(sc3, sc1) and int(b3) == int(b1)
    # host-buffer entry point agrees with the device one
    ph, sh, bh = e.register_host(poses, 5)
    assert torch.equal(ph.cuda(), p1) and bh == int(b1)
    assert torch.isfinite(sc1).all() and torch.isfinite(p1).all()


def test_foundationpose_api(setup):
    """estimater.FoundationPose.register()/track_one() drop-in surface."""
    from foundationpose_b200 import synth
    from foundationpose_b200.estimater import FoundationPose, PoseRefinePredictor, ScorePredictor

    s = setup
    mesh = s["mesh"]
    scorer = ScorePredictor(engine=s["e"], state_dict=s["sd_s"])
    refiner = PoseRefinePredictor(engine=s["e"], state_dict=s["sd_r"])
    est = FoundationPose(model_pts=mesh.vertices, model_normals=mesh.vertex_normals, mesh=mesh, scorer=scorer, refiner=refiner)
    with pytest.raises(RuntimeError):
        est.track_one(rgb=s["rgb"], depth=s["depth"], K=s["K"], iteration=2)
    pose = est.register(K=s["K"], rgb=s["rgb"], depth=s["depth"], ob_mask=s["mask"], iteration=2)
    assert pose.shape == (4, 4) and np.isfinite(pose).all()
    assert est.poses.shape == (252, 4, 4) and est.scores.shape == (252,)
    assert float(est.scores[0]) == float(est.scores.max())
    p2 = est.track_one(rgb=s["rgb"], depth=s["depth"], K=s["K"], iteration=2)
    assert p2.shape == (4, 4) and np.isfinite(p2).all()
    # empty mask -> identity rotation + zero translation fallback (estimater.py:184-189, :139-141)
    p3 = est.register(K=s["K"], rgb=s["rgb"], depth=s["depth"], ob_mask=np.zeros_like(s["mask"]), iteration=1)
    np.testing.assert_array_equal(p3, np.eye(4))
    # restore the fixture's frame/mesh state for other tests
    s["e"].set_mesh(s["mt"]["pos"], s["mt"]["normals"], s["mt"]["faces"], s["d"], uv=s["mt"]["uv"], tex=s["mt"]["tex"])
    s["e"].set_frame(s["rgb"], s["depth"], s["K"], filter_depth=False)


def test_predict_honours_per_call_arguments(setup):
    """PoseRefinePredictor.predict / ScorePredictor.predict with the reference's per-call arguments
    (predict_pose_refine.py:150-177, predict_score.py:160-180): another mesh is uploaded, a caller-supplied xyz map is
    used, each predictor keeps its own crop_ratio."""
    from foundationpose_b200 import synth
    from foundationpose_b200.engine import Engine
    from foundationpose_b200.estimater import PoseRefinePredictor, ScorePredictor, make_mesh_tensors
    from oracle import geometry

    s = setup
    e = Engine()
    refiner = PoseRefinePredictor(engine=e, state_dict=s["sd_r"])
    scorer = ScorePredictor(engine=e, state_dict=s["sd_s"], cfg={"crop_ratio": 1.1})
    assert scorer.cfg["crop_ratio"] == 1.1 and refiner.cfg["crop_ratio"] == 1.2
    poses = s["poses"][:3]
    mt = make_mesh_tensors(s["mesh"])
    xyz = geometry.depth2xyzmap(s["depth"], s["K"])
    # no mesh in the context and none passed: an error, not garbage
    with pytest.raises(ValueError):
        refiner.predict(rgb=s["rgb"], depth=s["depth"], K=s["K"], ob_in_cams=poses, xyz_map=xyz, iteration=1)
    a, _ = refiner.predict(rgb=s["rgb"], depth=s["depth"], K=s["K"], ob_in_cams=poses, xyz_map=xyz, mesh_tensors=mt, mesh_diameter=s["d"], iteration=1)
    ref, _, _ = s["e"].refine(poses, 1)
    assert torch.equal(a, ref), "same mesh / frame through the per-call arguments must give the same poses"
    # a caller-supplied xyz map is what the observed crop is built from
    b, _ = refiner.predict(rgb=s["rgb"], depth=s["depth"], K=s["K"], ob_in_cams=poses, xyz_map=np.zeros_like(xyz), mesh_tensors=mt,
                           mesh_diameter=s["d"], iteration=1)
    assert not torch.equal(a, b)
    # another mesh passed per call replaces the one in the context
    small = make_mesh_tensors(synth.make_mesh(2))
    c, _ = refiner.predict(rgb=s["rgb"], depth=s["depth"], K=s["K"], ob_in_cams=poses, xyz_map=xyz, mesh_tensors=small, mesh_diameter=s["d"], iteration=1)
    assert e.mesh_info()["F"] == 320 and not torch.equal(a, c)
    # the scorer's own crop_ratio: scores differ from a scorer configured with the refiner's 1.2
    sc_a, _ = scorer.predict(rgb=s["rgb"], depth=s["depth"], K=s["K"], ob_in_cams=poses, mesh_tensors=mt, mesh_diameter=s["d"])
    e2 = Engine()
    scorer2 = ScorePredictor(engine=e2, state_dict=s["sd_s"])
    sc_b, _ = scorer2.predict(rgb=s["rgb"], depth=s["depth"], K=s["K"], ob_in_cams=poses, mesh_tensors=mt, mesh_diameter=s["d"])
    assert not torch.equal(sc_a, sc_b)


def test_decoder_heads_on_two_streams_is_bitwise_serial(setup, monkeypatch):
    """The two refiner decoder heads run on two streams at small batches (fp_api.cu run_refine_heads, fork / join
    captured into the graph).  Same kernels on the same data: the poses must be bit-identical to the serial order,
    eagerly (first call), while capturing (second) and on graph replay (third)."""
    from foundationpose_b200.engine import Engine

    s = setup
    outs = {}
    for fork in ("0", "1000"):
        monkeypatch.setenv("FPOSE_FORK_MAX_N", fork)  # read by fp_create
        e = Engine()
        e.load_network("refine", s["sd_r"])
        e.set_mesh(s["mt"]["pos"], s["mt"]["normals"], s["mt"]["faces"], s["d"], uv=s["mt"]["uv"], tex=s["mt"]["tex"])
        e.set_frame(s["rgb"], s["depth"], s["K"], filter_depth=False)
        outs[fork] = [tuple(t.cpu().clone() for t in e.refine(s["poses"], 3)) for _ in range(3)]
    for call in range(3):
        for a, b in zip(outs["0"][call], outs["1000"][call]):
            assert torch.equal(a, b)
        for a, b in zip(outs["0"][0], outs["0"][call]):
            assert torch.equal(a, b)
