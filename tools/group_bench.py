"""Sharded register through fp_group (ONE process, one host thread, G GPUs): ms per register from host buffers.

    python tools/group_bench.py --gpus 8
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
    ap.add_argument("--gpus", type=int, default=torch.cuda.device_count())
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--hyp", type=int, default=252, help="use only the first HYP rotations of the grid (shard-size experiments)")
    ap.add_argument("--devices", type=str, default="", help="explicit device list, e.g. 0,0: TWO contexts (two streams) on one GPU")
    a = ap.parse_args()
    from foundationpose_b200 import hypotheses, synth
    from foundationpose_b200.estimater import make_mesh_tensors
    from foundationpose_b200.group import EngineGroup
    from foundationpose_b200.weights import random_state_dict

    mesh, gt, K, rgb, depth, mask = synth.default_scene(5, 0)
    mt = make_mesh_tensors(mesh)
    d = synth.mesh_diameter(mesh.vertices)
    grid = hypotheses.make_rotation_grid()[: a.hyp]
    devs = [int(x) for x in a.devices.split(",")] if a.devices else list(range(a.gpus))
    g = EngineGroup(devs)
    g.load_network("refine", random_state_dict("refine", 0))
    g.load_network("score", random_state_dict("score", 0))
    g.set_mesh(mt["pos"], mt["normals"], mt["faces"], d, uv=mt["uv"], tex=mt["tex"])
    for _ in range(4):
        poses, scores, best, info = g.register(rgb, depth, K, mask, grid, iterations=5)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        poses, scores, best, info = g.register(rgb, depth, K, mask, grid, iterations=5)
    ms = (time.perf_counter() - t0) / a.steps * 1e3
    print(json.dumps({"what": "fp_group_register: one process, one host thread; host frame / mask / grid in, host poses / scores out; "
                              "features gathered into device 0 over NVLink peer memory", "devices": devs, "ms_per_register": ms,
                      "hypotheses": len(grid), "hyp_per_s": len(grid) / (ms * 1e-3), "best_index": best}))
    g.close()


if __name__ == "__main__":
    main()
