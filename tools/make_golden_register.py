"""Generate tests/golden/register_252x5.npz: the CPU oracle run over the WHOLE headline configuration —
252 start poses x 5 refine iterations + scoring + ranking (estimater.py:159-240) — on the icosphere-3
scene of tests/test_pipeline_gpu.py (1 280 faces keeps the Python rasteriser to minutes).

    python tools/make_golden_register.py            # encoder passes (slow, ~10 min on 8 cores); cached in /tmp
    python tools/make_golden_register.py --tail     # only re-run the cross-hypothesis tail on the cached features

Stored (all fp32 unless noted): start poses, per-iteration refined poses [6][252][4][4] (index 0 = start),
last_trans / last_rot deltas of every iteration, scorer features [252][512], scores [252], argsort (int64),
best index, top-1/top-2 margin.  The GPU test (tests/test_register_golden_gpu.py) holds the CUDA path to these:
SE(3) delta of every hypothesis at every iteration within 1e-3, scores within tolerance, index equal.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CACHE = "/tmp/register_252x5_cache.npz"
N_ITER = 5


def scene():
    """Same scene as tests/test_pipeline_gpu.py::setup."""
    from foundationpose_b200 import synth
    from oracle import pipeline

    mesh = synth.make_mesh(3)
    pose = np.eye(4)
    pose[:3, :3] = synth.random_rotation(0)
    pose[:3, 3] = [0.02, -0.01, 0.6]
    rgb, depth, mask = synth.make_scene(mesh.visual.image, pose)
    d = synth.mesh_diameter(mesh.vertices)
    mt = pipeline.mesh_tensors(mesh)
    return mesh, mt, pose, rgb, depth, mask, synth.DEFAULT_K.copy(), d


def encoder_passes():
    from foundationpose_b200 import hypotheses
    from foundationpose_b200.weights import random_state_dict
    from oracle import geometry, nets, pipeline

    torch.set_num_threads(os.cpu_count())
    mesh, mt, gt, rgb, depth, mask, K, d = scene()
    # register() front end (estimater.py:173-174, :203-209): filtered depth, translation guess, 252 start poses
    depth_f = geometry.bilateral_filter_depth(geometry.erode_depth(depth))
    center = hypotheses.guess_translation(depth_f, mask, K)
    grid = hypotheses.make_rotation_grid()
    assert grid.shape == (252, 4, 4)
    start = grid.copy().astype(np.float32)
    start[:, :3, 3] = center
    xyz_map = geometry.depth2xyzmap(depth_f, K)
    sd_r, sd_s = random_state_dict("refine", 0), random_state_dict("score", 0)

    N = len(start)
    poses = torch.from_numpy(start.copy())
    all_poses = [poses.numpy().copy()]
    all_lt, all_lr, all_raw_t, all_raw_r = [], [], [], []
    CH = 42
    for it in range(N_ITER):
        t0 = time.time()
        trans = torch.empty(N, 3)
        rot = torch.empty(N, 3)
        for lo in range(0, N, CH):
            A, B, _ = pipeline.make_crops(poses[lo:lo + CH].numpy(), mt, rgb, depth_f, xyz_map, K, d, 0)
            o = nets.refine_forward(sd_r, A, B)
            trans[lo:lo + CH] = o["trans"]
            rot[lo:lo + CH] = o["rot"]
        poses, td, rd = geometry.pose_update(poses, trans, rot, d, 0.3490658503988659)
        all_poses.append(poses.numpy().copy())
        all_lt.append(td.numpy().copy())
        all_lr.append(rd.numpy().copy())
        all_raw_t.append(trans.numpy().copy())
        all_raw_r.append(rot.numpy().copy())
        print(f"iteration {it}: {time.time() - t0:.1f} s, |dt| max {float(td.abs().max()):.4f}", flush=True)
    feats = torch.empty(N, 512)
    t0 = time.time()
    for lo in range(0, N, CH):
        A, B, _ = pipeline.make_crops(poses[lo:lo + CH].numpy(), mt, rgb, depth_f, None, K, d, 1)
        feats[lo:lo + CH] = nets.score_features(sd_s, A, B)
    print(f"score features: {time.time() - t0:.1f} s", flush=True)
    np.savez(CACHE, start=start, poses=np.stack(all_poses), last_trans=np.stack(all_lt), last_rot=np.stack(all_lr),
             raw_trans=np.stack(all_raw_t), raw_rot=np.stack(all_raw_r), feats=feats.numpy(), center=np.asarray(center, dtype=np.float64),
             gt_pose=gt)
    print("cached", CACHE)


def score_feats_only():
    from foundationpose_b200.weights import random_state_dict
    from oracle import geometry, nets, pipeline

    torch.set_num_threads(os.cpu_count())
    mesh, mt, gt, rgb, depth, mask, K, d = scene()
    depth_f = geometry.bilateral_filter_depth(geometry.erode_depth(depth))
    sd_s = random_state_dict("score", 0)
    c = dict(np.load(CACHE))
    poses = c["poses"][N_ITER]
    feats = torch.empty(len(poses), 512)
    for lo in range(0, len(poses), 42):
        A, B, _ = pipeline.make_crops(poses[lo:lo + 42], mt, rgb, depth_f, None, K, d, 1)
        feats[lo:lo + 42] = nets.score_features(sd_s, A, B)
    c["feats"] = feats.numpy()
    np.savez(CACHE, **c)
    print("recomputed scorer features")


def tail():
    from foundationpose_b200.weights import random_state_dict
    from oracle import nets

    c = dict(np.load(CACHE))
    sd_s = random_state_dict("score", 0)
    feats = torch.from_numpy(c["feats"])
    logits = nets.score_tail(sd_s, feats, len(feats)).reshape(-1)
    scores = (logits + 100).numpy().astype(np.float32)  # predict_score.py:206
    ids = np.argsort(-scores, kind="stable")  # estimater.py:226 argsort(descending)
    c["scores"] = scores
    c["ids"] = ids.astype(np.int64)
    c["best"] = np.array([ids[0]], dtype=np.int64)
    c["top2_margin"] = np.array([scores[ids[0]] - scores[ids[1]]], dtype=np.float32)
    c["score_wsum"] = np.array([float(sum(v.double().abs().sum() for k, v in sd_s.items() if v.dtype.is_floating_point))])
    path = os.path.join(ROOT, "tests", "golden", "register_252x5.npz")
    np.savez_compressed(path, **c)
    print(f"scores: min {scores.min():.5f} max {scores.max():.5f} std {scores.std():.5f}; best {ids[0]} second {ids[1]} "
          f"top-2 margin {float(c['top2_margin'][0]):.5f}")
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--tail", action="store_true")
    ap.add_argument("--feats", action="store_true", help="recompute only the scorer features on the cached final poses")
    a = ap.parse_args()
    if a.feats and os.path.exists(CACHE):
        score_feats_only()
    elif not a.tail or not os.path.exists(CACHE):
        encoder_passes()
    tail()
