"""Frames/s of multi-frame registration with one estimator replica per GPU (the reference's
run_ycb_video.py / run_linemod.py loops, which are sequential on one GPU).

    python tools/replica_bench.py --gpus 8 --frames 64
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--frames", type=int, default=32)
    ap.add_argument("--objects", type=int, default=2)
    a = ap.parse_args()
    from foundationpose_b200 import synth
    from foundationpose_b200.replicas import ReplicaPool
    from foundationpose_b200.weights import random_state_dict

    sds = {"refine": random_state_dict("refine", 0), "score": random_state_dict("score", 0)}
    pool = ReplicaPool(range(a.gpus), state_dicts=sds)
    res = []
    for ob in range(a.objects):
        mesh = synth.make_mesh(5 if ob == 0 else 4, tex_seed=ob)
        pose0 = np.eye(4)
        pose0[:3, :3] = synth.random_rotation(ob)
        pose0[:3, 3] = [0.02, -0.01, 0.6]
        seq = synth.track_sequence(8, pose0, seed=3 + ob)
        distinct = [synth.make_scene(mesh.visual.image, p, seed=1 + i) for i, p in enumerate(seq)]
        frames = [(synth.DEFAULT_K, *distinct[i % len(distinct)]) for i in range(a.frames)]
        t0 = time.perf_counter()
        pool.reset_object(mesh.vertices, mesh.vertex_normals, mesh=mesh)
        t_reset = time.perf_counter() - t0
        pool.register_many(frames[: 2 * a.gpus])  # graphs captured on every replica
        t0 = time.perf_counter()
        poses = pool.register_many(frames)
        dt = time.perf_counter() - t0
        assert all(np.isfinite(p).all() for p in poses)
        res.append({"object": ob, "faces": int(len(mesh.faces)), "reset_object_s": t_reset, "frames": a.frames, "seconds": dt,
                    "frames_per_s": a.frames / dt, "hyp_per_s": 252 * a.frames / dt})
    pool.close()
    print(json.dumps({"what": "register(252 hyp x 5 iters + score) over independent frames, one estimator replica per GPU, one process", "gpus": a.gpus,
                      "per_object": res, "frames_per_s": float(np.mean([r["frames_per_s"] for r in res]))}))


if __name__ == "__main__":
    main()
