"""2-GPU: the sharded register (one process per GPU, NCCL all-gather of per-hypothesis features) returns
bit-identical poses, scores and best index to the single-GPU run.  Skipped on boxes with < 2 GPUs."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from foundationpose_b200 import hypotheses, synth
    from foundationpose_b200.engine import Engine
    from foundationpose_b200.estimater import make_mesh_tensors
    from foundationpose_b200.parallel import ShardedRegister
    from foundationpose_b200.weights import random_state_dict

    mesh, gt, K, rgb, depth, mask = synth.default_scene(3, 0)
    e = Engine()
    e.load_network("refine", random_state_dict("refine", 0))
    e.load_network("score", random_state_dict("score", 0))
    mt = make_mesh_tensors(mesh)
    e.set_mesh(mt["pos"], mt["normals"], mt["faces"], synth.mesh_diameter(mesh.vertices), uv=mt["uv"], tex=mt["tex"])
    e.set_frame(rgb, depth, K, filter_depth=True)
    d, _ = e.get_depth()
    poses = hypotheses.make_rotation_grid()
    poses[:, :3, 3] = hypotheses.guess_translation(d.cpu().numpy(), mask, K)
    poses = torch.from_numpy(poses)
    sh = ShardedRegister(e)
    p, s, b = sh.run(poses, 2)
    out = dict(rank=rank, poses=p.cpu().numpy(), scores=s.cpu().numpy(), best=int(b.item()))
    if rank == 0:
        # single-GPU reference on the same device
        p1, _, _ = e.refine(poses, 2)
        s1, b1 = e.score(p1)
        out.update(poses1=p1.cpu().numpy(), scores1=s1.cpu().numpy(), best1=int(b1.item()))
    q.put(out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_register_equals_single_gpu():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29611, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in procs], key=lambda d: d["rank"])
    for p in procs:
        p.join(timeout=120)
    r0, r1 = res
    assert np.array_equal(r0["poses"], r1["poses"]) and np.array_equal(r0["scores"], r1["scores"]) and r0["best"] == r1["best"]
    assert np.array_equal(r0["poses"], r0["poses1"]), "sharded refinement differs from single-GPU"
    assert np.array_equal(r0["scores"], r0["scores1"]) and r0["best"] == r0["best1"]


def _group_case(device_ids):
    """fp_group (one process, one host thread): sharded register == single-context register, bit for bit."""
    import numpy as np
    import torch

    from foundationpose_b200 import hypotheses, synth
    from foundationpose_b200.engine import Engine
    from foundationpose_b200.estimater import make_mesh_tensors
    from foundationpose_b200.group import EngineGroup
    from foundationpose_b200.weights import random_state_dict

    mesh = synth.make_mesh(3)
    pose = np.eye(4)
    pose[:3, :3] = synth.random_rotation(0)
    pose[:3, 3] = [0.02, -0.01, 0.6]
    rgb, depth, mask = synth.make_scene(mesh.visual.image, pose)
    d = synth.mesh_diameter(mesh.vertices)
    mt = make_mesh_tensors(mesh)
    sd_r, sd_s = random_state_dict("refine", 0), random_state_dict("score", 0)
    grid = hypotheses.make_rotation_grid()[:60]
    # single context
    e = Engine()
    e.load_network("refine", sd_r)
    e.load_network("score", sd_s)
    e.set_mesh(mt["pos"], mt["normals"], mt["faces"], d, uv=mt["uv"], tex=mt["tex"])
    e.set_frame(rgb, depth, synth.DEFAULT_K, filter_depth=True)
    start, info1 = e.start_poses(mask, torch.from_numpy(grid).cuda())
    p1, _, _ = e.refine(start, 2)
    s1, b1 = e.score(p1)
    # group
    g = EngineGroup(device_ids)
    g.load_network("refine", sd_r)
    g.load_network("score", sd_s)
    g.set_mesh(mt["pos"], mt["normals"], mt["faces"], d, uv=mt["uv"], tex=mt["tex"])
    for _ in range(3):  # eager, capture, replay
        p2, s2, b2, info2 = g.register(rgb, depth, synth.DEFAULT_K, mask, grid, iterations=2)
    assert np.array_equal(p2, p1.cpu().numpy()), "sharded refinement differs"
    assert np.array_equal(s2, s1.cpu().numpy()) and b2 == int(b1.item())
    assert np.array_equal(info2, info1.cpu().numpy())
    g.close()


@pytest.mark.gpu
def test_group_two_contexts_one_device():
    """The sharding / gather logic of fp_group with both contexts on device 0 (runs on a 1-GPU box)."""
    _group_case([0, 0])


@pytest.mark.gpu
def test_group_two_devices():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _group_case([0, 1])


@pytest.mark.gpu
def test_replica_pool_matches_single_estimator():
    """ReplicaPool (N3: independent frames over estimator replicas) returns, frame by frame, what one estimator returns."""
    import numpy as np
    import torch

    from foundationpose_b200 import synth
    from foundationpose_b200.engine import Engine
    from foundationpose_b200.estimater import FoundationPose, PoseRefinePredictor, ScorePredictor
    from foundationpose_b200.replicas import ReplicaPool
    from foundationpose_b200.weights import random_state_dict

    sds = {"refine": random_state_dict("refine", 0), "score": random_state_dict("score", 0)}
    mesh = synth.make_mesh(3)
    pose0 = np.eye(4)
    pose0[:3, :3] = synth.random_rotation(0)
    pose0[:3, 3] = [0.02, -0.01, 0.6]
    seq = synth.track_sequence(5, pose0)
    frames = [(synth.DEFAULT_K, *synth.make_scene(mesh.visual.image, p, seed=1 + i)) for i, p in enumerate(seq)]
    e = Engine()
    est = FoundationPose(model_pts=mesh.vertices, model_normals=mesh.vertex_normals, mesh=mesh,
                         scorer=ScorePredictor(engine=e, state_dict=sds["score"]), refiner=PoseRefinePredictor(engine=e, state_dict=sds["refine"]))
    ref = [est.register(K=K, rgb=rgb, depth=depth, ob_mask=mask, iteration=2) for (K, rgb, depth, mask) in frames]
    ids = [0, 1] if torch.cuda.device_count() >= 2 else [0, 0]
    pool = ReplicaPool(ids, state_dicts=sds)
    pool.reset_object(mesh.vertices, mesh.vertex_normals, mesh=mesh)
    got = pool.register_many(frames, iteration=2)
    pool.close()
    for a, b in zip(ref, got):
        np.testing.assert_array_equal(a, b)
