"""the LINEMOD evaluation loop of the reference (run_linemod.py:88-128 — for every object:
`reset_object`, then `register` on every frame of the object's scene, results to linemod_res.yml) with the frames of
an object spread over ALL GPUs of the box by `foundationpose_b200.replicas.ReplicaPool` instead of running one after
the other on one GPU.  Same readers, same result file, same poses (tests/test_dropin_gpu.py compares them).

    python examples/run_linemod_replicas.py --linemod_dir <root> [--gpus 8] [--debug_dir out]
    python examples/run_linemod_replicas.py --synthetic 8          # writes a synthetic dataset with 8 frames per object first
"""
import argparse
import os
import sys
import tempfile
import time

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "foundationpose_b200", "dropin"), ROOT]

from datareader import LinemodReader  # noqa: E402  (drop-in module tree)
from Utils import NestDict, make_yaml_dumpable, set_seed  # noqa: E402

from foundationpose_b200.replicas import ReplicaPool  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--linemod_dir", type=str, default=None)
    ap.add_argument("--synthetic", type=int, default=0, help="write a synthetic LINEMOD-layout dataset with this many frames per object")
    ap.add_argument("--gpus", type=int, default=torch.cuda.device_count())
    ap.add_argument("--debug_dir", type=str, default=os.path.join(ROOT, "debug"))
    opt = ap.parse_args()
    set_seed(0)
    if opt.synthetic:
        from foundationpose_b200 import synth

        opt.linemod_dir = opt.linemod_dir or os.path.join(tempfile.mkdtemp(prefix="fpose_lm_"), "LINEMOD")
        synth.write_bop_dataset(opt.linemod_dir, "lm", n_frames=opt.synthetic)
    assert opt.linemod_dir, "--linemod_dir or --synthetic"
    os.makedirs(opt.debug_dir, exist_ok=True)
    reader_tmp = LinemodReader(f"{opt.linemod_dir}/lm_test_all/test/000002", split=None)
    pool = ReplicaPool(range(opt.gpus))
    res = NestDict()
    n_frames, t_reg = 0, 0.0
    for ob_id in reader_tmp.ob_ids:
        ob_id = int(ob_id)
        mesh = reader_tmp.get_gt_mesh(ob_id)
        reader = LinemodReader(f"{opt.linemod_dir}/lm_test_all/test/{ob_id:06d}", split=None)
        pool.reset_object(model_pts=mesh.vertices.copy(), model_normals=mesh.vertex_normals.copy(),
                          symmetry_tfs=reader_tmp.symmetry_tfs[ob_id], mesh=mesh)
        frames, keys = [], []
        for i in range(len(reader.color_files)):
            mask = reader.get_mask(i, ob_id)
            if mask is None:  # run_linemod.py:66-69
                res[reader.get_video_id()][reader.id_strs[i]][ob_id] = np.eye(4)
                continue
            frames.append((reader.K, reader.get_color(i), reader.get_depth(i), mask > 0))
            keys.append((reader.get_video_id(), reader.id_strs[i]))
        t0 = time.perf_counter()
        poses = pool.register_many(frames, iteration=5)
        t_reg += time.perf_counter() - t0
        n_frames += len(frames)
        for (vid, id_str), pose in zip(keys, poses):
            res[vid][id_str][ob_id] = pose
    pool.close()
    with open(f"{opt.debug_dir}/linemod_res.yml", "w") as fh:
        yaml.safe_dump(make_yaml_dumpable(res), fh)
    print(f"{n_frames} frames, {len(reader_tmp.ob_ids)} objects on {opt.gpus} GPU(s): {n_frames / max(t_reg, 1e-9):.1f} registers/s "
          f"(252 hypotheses x 5 iterations each); results in {opt.debug_dir}/linemod_res.yml")


if __name__ == "__main__":
    main()
