"""Where does the host-side time of FoundationPose.register() go?  Phase timings (wall clock, a device synchronise after
every phase: diagnostic only) and the un-instrumented end-to-end time, for the bench scene."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from foundationpose_b200 import synth  # noqa: E402
from foundationpose_b200.engine import Engine  # noqa: E402
from foundationpose_b200.estimater import FoundationPose, PoseRefinePredictor, ScorePredictor  # noqa: E402
from foundationpose_b200.weights import random_state_dict  # noqa: E402


def main():
    mesh, gt, K, rgb, depth, mask = synth.default_scene(5, 0)
    e = Engine()
    est = FoundationPose(model_pts=mesh.vertices, model_normals=mesh.vertex_normals, mesh=mesh,
                         scorer=ScorePredictor(engine=e, state_dict=random_state_dict("score", 0)),
                         refiner=PoseRefinePredictor(engine=e, state_dict=random_state_dict("refine", 0)))
    for _ in range(4):
        est.register(K=K, rgb=rgb, depth=depth, ob_mask=mask, iteration=5)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        est.register(K=K, rgb=rgb, depth=depth, ob_mask=mask, iteration=5)
    print(f"register() end to end: {(time.perf_counter() - t0) / 10 * 1e3:.3f} ms")
    ph = {}

    def tick(name, t):
        torch.cuda.synchronize()
        ph[name] = ph.get(name, 0.0) + (time.perf_counter() - t)

    for _ in range(10):
        t = time.perf_counter(); e.set_frame(rgb, depth, K, filter_depth=True); t1 = time.perf_counter(); ph["set_frame call"] = ph.get("set_frame call", 0) + t1 - t; tick("set_frame", t)
        t = time.perf_counter(); poses, info = e.start_poses(mask, est.rot_grid); tick("start_poses", t)
        t = time.perf_counter(); p, lt, lr = e.refine(poses, 5); t1 = time.perf_counter(); ph["refine call"] = ph.get("refine call", 0) + t1 - t; tick("refine", t)
        t = time.perf_counter(); s, b = e.score(p); tick("score", t)
        t = time.perf_counter(); ids = s.argsort(descending=True); out = torch.cat([info, (p[ids][0] @ est.get_tf_to_centered_mesh()).reshape(-1)]).cpu().numpy(); tick("rank + read-back", t)
    for k, v in ph.items():
        print(f"  {k:18s} {v / 10 * 1e3:8.3f} ms")
    # raw transfer costs of this box (the e2e - value gap moves between 0.1 and 3 ms from box to box)
    pin = torch.empty(2457600, dtype=torch.uint8).pin_memory()
    dev = torch.empty(2457600, dtype=torch.uint8, device="cuda")
    page = np.zeros(2457600, dtype=np.uint8)
    small = torch.zeros(20, device="cuda")
    for name, fn in (("H2D 2.4 MB pinned", lambda: dev.copy_(pin, non_blocking=True)),
                     ("H2D 2.4 MB pageable", lambda: dev.copy_(torch.from_numpy(page))),
                     ("host memcpy 2.4 MB -> pinned", lambda: pin.copy_(torch.from_numpy(page))),
                     ("D2H 80 B + sync", lambda: small.cpu())):
        ts = []
        for _ in range(30):
            torch.cuda.synchronize()
            t = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t) * 1e3)
        print(f"  {name:30s} first {ts[0]:7.3f} ms   median {sorted(ts)[15]:7.3f} ms   max {max(ts):7.3f} ms")
    # the same call after 50 ms of idle GPU (does the link / clock need to wake up?)
    for idle in (0.0, 0.05, 0.5):
        ts = []
        for _ in range(5):
            time.sleep(idle)
            t = time.perf_counter()
            est.register(K=K, rgb=rgb, depth=depth, ob_mask=mask, iteration=5)
            ts.append((time.perf_counter() - t) * 1e3)
        print(f"  register() after {idle * 1e3:5.0f} ms idle: median {sorted(ts)[2]:7.3f} ms  (min {min(ts):.3f}, max {max(ts):.3f})")
    os.system("nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current,clocks.sm,power.draw --format=csv,noheader")
    os.system("lscpu | grep -E 'Model name|^CPU\\(s\\)|MHz' | head -4")


if __name__ == "__main__":
    main()
