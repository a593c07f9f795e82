"""Generate tests/golden/track_seq.npz: the CPU oracle over the track_one path (estimater.py:250-268: erode + bilateral
depth, depth2xyzmap_batch(zfar=inf), ONE pose through `iteration` = 2 refiner passes) on a 50-frame synthetic sequence
(foundationpose_b200.synth.write_demo_scene / track_sequence: <= 5 mm, <= 2 deg of object motion per frame).

Two records:
  * per-frame: frame i is tracked from the previous frame's ground-truth pose plus a small seeded perturbation (what a
    converged tracker hands over).  The stand-in weights are random-init — they do not converge — so feeding the pose
    back for 100 passes would walk the object out of every crop after a few frames and the test would compare empty
    renders; anchoring each frame keeps all 50 comparisons on real render-and-compare inputs.  Inside a frame the
    second pass IS fed from the first.
  * chain: frames 1..5 with the pose fed back from frame to frame (10 chained passes), as track_one does.

    python tools/make_golden_track.py
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

N_FRAMES = 50
K_ITERS = 2


def main():
    from foundationpose_b200 import synth
    from foundationpose_b200.weights import random_state_dict
    from oracle import geometry, pipeline

    torch.set_num_threads(os.cpu_count())
    mesh = synth.make_mesh(3)
    pose0 = np.eye(4)
    pose0[:3, :3] = synth.random_rotation(0)
    pose0[:3, 3] = [0.02, -0.01, 0.6]
    gt = synth.track_sequence(N_FRAMES, pose0)
    d = synth.mesh_diameter(mesh.vertices)
    mt = pipeline.mesh_tensors(mesh)
    K = synth.DEFAULT_K
    sd_r = random_state_dict("refine", 0)
    rng = np.random.default_rng(11)

    def frame(i):
        rgb, depth, _ = synth.make_scene(mesh.visual.image, gt[i], seed=1 + i)
        depth_f = geometry.bilateral_filter_depth(geometry.erode_depth(depth))
        return rgb, depth_f

    pose_in, pose_out, lt, lr = [], [], [], []
    t0 = time.time()
    for i in range(1, N_FRAMES):
        rgb, depth_f = frame(i)
        p = gt[i - 1].copy()
        ax = rng.normal(size=3)
        ax /= np.linalg.norm(ax)
        ang = np.deg2rad(1.0)
        Kx = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
        p[:3, :3] = (np.eye(3) + np.sin(ang) * Kx + (1 - np.cos(ang)) * (Kx @ Kx)) @ p[:3, :3]
        p[:3, 3] += rng.normal(0, 0.003, 3)
        p = p.astype(np.float32)
        out, td, rd = pipeline.refine(sd_r, p[None], mt, rgb, depth_f, K, d, K_ITERS)
        pose_in.append(p)
        pose_out.append(out[0].numpy())
        lt.append(td[0].numpy())
        lr.append(rd[0].numpy())
        if i % 10 == 0:
            print(f"frame {i}: {time.time() - t0:.0f} s", flush=True)
    chain = [pose_in[0].copy()]
    for i in range(1, 6):
        rgb, depth_f = frame(i)
        out, _, _ = pipeline.refine(sd_r, chain[-1][None], mt, rgb, depth_f, K, d, K_ITERS)
        chain.append(out[0].numpy())
    path = os.path.join(ROOT, "tests", "golden", "track_seq.npz")
    np.savez_compressed(path, gt=gt, pose_in=np.stack(pose_in), pose_out=np.stack(pose_out), last_trans=np.stack(lt),
                        last_rot=np.stack(lr), chain=np.stack(chain))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
