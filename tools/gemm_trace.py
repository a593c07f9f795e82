"""Phase-by-phase cost of every gemm_tile_kernel launch of a refine pass (needs the -DFP_GEMM_TRACE build):

    tools/build_variant.sh trace -DFP_GEMM_TRACE
    FPOSE_LIB_PATH=$PWD/foundationpose_b200/lib/variants/libfpose_trace.so python tools/gemm_trace.py [N] [iters]

CTA 0 of each launch stamps: entry, prologue done, dependency wait done, first operand stage landed, last MMA
issued, first accumulator ready (epilogue start), last store issued, stores complete; plus the global timer at
entry / exit (gaps between launches).  The graph is replayed several times; the LAST replay's stamps are read.
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from foundationpose_b200 import _lib, synth  # noqa: E402
from foundationpose_b200.engine import Engine  # noqa: E402
from foundationpose_b200.weights import random_state_dict  # noqa: E402
from oracle import pipeline  # noqa: E402  (mesh tensors only: this is a tool, not the product)


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    lib = _lib.lib
    mesh = synth.make_mesh(5)
    pose = np.eye(4)
    pose[:3, :3] = synth.random_rotation(0)
    pose[:3, 3] = [0.0, 0.0, 0.6]
    rgb, depth, mask = synth.make_scene(mesh.visual.image, pose, seed=1)
    mt = pipeline.mesh_tensors(mesh)
    e = Engine()
    e.load_network("refine", random_state_dict("refine", 0))
    e.set_mesh(mt["pos"], mt["normals"], mt["faces"], synth.mesh_diameter(mesh.vertices), uv=mt["uv"], tex=mt["tex"])
    e.set_frame(rgb, depth, synth.DEFAULT_K, filter_depth=True)
    poses = np.tile(pose[None], (N, 1, 1)).astype(np.float32)
    e.refine(poses, iters)          # eager
    lib.fp_op_gemm_trace_reset()
    e.refine(poses, iters)          # captured: trace slots are assigned here
    for _ in range(5):
        e.refine(poses, iters)      # replays overwrite the same slots
    torch.cuda.synchronize()
    buf = np.zeros((512, 10), dtype=np.uint64)
    info = np.zeros((512, 10), dtype=np.int32)
    lib.fp_op_gemm_trace_read.restype = C.c_int
    n = lib.fp_op_gemm_trace_read(C.c_void_p(buf.ctypes.data), C.c_void_p(info.ctypes.data), 512)
    mhz = float(os.environ.get("SM_MHZ", "1900"))
    print(f"# {n} gemm_tile_kernel launches, N = {N}, {iters} iterations; cycles -> us at {mhz:.0f} MHz (SM_MHZ)")
    print("# idx  BN CG P grid  vt  kb res |  prolog  depwait  1st-data   mma-loop  acc-ready(from data)  epilogue  store-drain | total(us)  gap-to-prev(us, global timer)")
    per = n // iters if iters else n
    prev_exit = None
    for i in range(n):
        t = buf[i].astype(np.int64)
        bn, cg, slabs, patch, grid, vt, kb, mt_, cout, res = info[i]
        us = lambda a, b: (t[b] - t[a]) / mhz  # noqa: E731
        gap = (t[8] - prev_exit) / 1e3 if prev_exit is not None else float("nan")
        prev_exit = t[9]
        print(f"{i:4d} {bn:4d} {cg:2d} {patch:1d} {grid:4d} {vt:4d} {kb:3d} {res:2d} | {us(0, 1):7.2f} {us(1, 2):8.2f} {us(2, 3):9.2f} {us(3, 4):10.2f} "
              f"{us(3, 5):10.2f} {us(5, 6):9.2f} {us(6, 7):11.2f} | {(t[9] - t[8]) / 1e3:8.2f} {gap:10.2f}")
        if per and (i + 1) % per == 0:
            print("# ---- iteration boundary")


if __name__ == "__main__":
    main()
