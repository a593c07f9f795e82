"""Differential cases for the dataset readers (not a test module; imported by tools/make_golden_readers.py and
tests/test_readers_golden_cpu.py).

`build(root)` writes small synthetic datasets in every directory layout the reference's `datareader.py` understands
(:57-152 demo / YCBInEOAT scenes, :155-613 the BOP readers) and returns the environment the reader module needs;
`collect(module, root)` drives a reader module — the reference's own file or this repository's drop-in — through its
whole public surface on those trees and returns {name: array}.  The golden fixture holds what the REFERENCE's
unmodified classes return (run on the drop-in `Utils`, which supplies cv2 / imageio / trimesh / depth2xyzmap /
symmetry_tfs_from_info); the CPU test holds the drop-in readers to it.

Big arrays are recorded as (shape, dtype, sha1 of the bytes) when they must be identical (images, masks, meshes) and as a
strided sample + sum when ulp-level differences are acceptable (metric depth, xyz maps).
"""
import hashlib
import json
import os
import shutil

import numpy as np


def _sha(a):
    a = np.ascontiguousarray(a)
    return np.frombuffer(hashlib.sha1(a.tobytes()).digest(), dtype=np.uint8).copy()


def build(root):
    """Writes the trees under `root` (deterministic) and returns {env var: value} to set BEFORE importing a reader module
    (`BOP_DIR` is read at import, datareader.py:14)."""
    import cv2

    from foundationpose_b200 import synth

    # ---- demo / YCBInEOAT scenes (datareader.py:57-152)
    demo = os.path.join(root, "demo", "mustard0")
    mesh, _ = synth.write_demo_scene(demo, n_frames=3, subdivisions=2)
    os.makedirs(os.path.join(demo, "masks_hand"))
    hand = np.zeros((480, 640), np.uint8)
    hand[50:120, 300:420] = 255
    cv2.imwrite(os.path.join(demo, "masks_hand", "000000.png"), hand)
    colour_mask = os.path.join(root, "demo", "bleach0")  # same scene, 3-channel mask with one filled channel, no poses
    shutil.copytree(demo, colour_mask)
    shutil.rmtree(os.path.join(colour_mask, "annotated_poses"))
    m = cv2.imread(os.path.join(colour_mask, "masks", "000000.png"), -1)
    cv2.imwrite(os.path.join(colour_mask, "masks", "000000.png"), np.stack([np.zeros_like(m), m, np.zeros_like(m)], -1))

    # ---- LINEMOD and YCB-Video in the drivers' layouts (run_linemod.py:90-112, run_ycb_video.py:85-118)
    lm = os.path.join(root, "LINEMOD")
    synth.write_bop_dataset(lm, "lm", n_frames=2)
    scene = os.path.join(lm, "lm_test_all", "test", "000001")
    gt = json.load(open(os.path.join(scene, "scene_gt.json")))
    a = dict(gt["0"][0])
    b = dict(a, cam_t_m2c=[100.0, 0.0, 700.0])                        # second instance of object 1
    c = dict(a, obj_id=5, cam_t_m2c=[-100.0, 50.0, 650.0])            # another object in the same frame
    gt["0"] = [a, c, b]
    json.dump(gt, open(os.path.join(scene, "scene_gt.json"), "w"))
    m0 = cv2.imread(os.path.join(scene, "mask_visib", "000000_000000.png"), -1)
    m1 = np.zeros_like(m0)
    m1[100:200, 400:500] = 255
    m2 = np.zeros_like(m0)
    m2[300:400, 100:200] = 255
    cv2.imwrite(os.path.join(scene, "mask_visib", "000000_000001.png"), m1)
    cv2.imwrite(os.path.join(scene, "mask_visib", "000000_000002.png"), m2)
    ycb = os.path.join(root, "YCB_Video")
    synth.write_bop_dataset(ycb, "ycbv", n_frames=2)
    synth.write_obj(mesh, os.path.join(ycb, "models", "006_mustard_bottle", "textured_simple.obj"))
    os.rmdir(os.path.join(ycb, "models", "006_synthetic_object"))    # keep 21 model directories (names sort into ids)

    # ---- the BOP'19 trees under $BOP_DIR (datareader.py:33-53, :369-613)
    bop = os.path.join(root, "bop")
    lm_models = os.path.join(lm, "lm_models", "models")
    src_info = json.load(open(os.path.join(lm_models, "models_info.json")))

    def models(dst, n_ids, symmetric=None):
        os.makedirs(dst, exist_ok=True)
        for i in range(1, min(n_ids, 2) + 1):  # the readers only ever open a model file when asked for a mesh
            shutil.copy(os.path.join(lm_models, "obj_000001.ply" if i == 1 else "obj_000002.ply"), os.path.join(dst, f"obj_{i:06d}.ply"))
        info = {str(i): dict(src_info["1"], diameter=100.0 + i) for i in range(1, n_ids + 1)}
        for i, sym in (symmetric or {}).items():
            info[str(i)].update(sym)
        json.dump(info, open(os.path.join(dst, "models_info.json"), "w"))

    def scene_copy(dst, src_id, grey=False, with_gt=True, depth_scale=None):
        shutil.copytree(os.path.join(lm, "lm_test_all", "test", f"{src_id:06d}"), dst)
        if grey:
            os.makedirs(os.path.join(dst, "gray"))
            for f in sorted(os.listdir(os.path.join(dst, "rgb"))):
                cv2.imwrite(os.path.join(dst, "gray", f), cv2.cvtColor(cv2.imread(os.path.join(dst, "rgb", f)), cv2.COLOR_BGR2GRAY))
            shutil.rmtree(os.path.join(dst, "rgb"))
        if not with_gt:
            os.remove(os.path.join(dst, "scene_gt.json"))
        if depth_scale is not None:
            cam = json.load(open(os.path.join(dst, "scene_camera.json")))
            for k in cam:
                cam[k]["depth_scale"] = depth_scale
            json.dump(cam, open(os.path.join(dst, "scene_camera.json"), "w"))

    def targets(dataset, scene_id, frames):
        json.dump([{"im_id": f, "inst_count": n, "obj_id": o, "scene_id": s} for (s, f, o, n) in frames],
                  open(os.path.join(bop, dataset, "test_targets_bop19.json"), "w"))

    continuous = {"symmetries_continuous": [{"axis": [0, 0, 1], "offset": [0, 0, 0]}]}
    discrete = {"symmetries_discrete": [np.diag([-1.0, -1.0, 1.0, 1.0]).reshape(-1).tolist(), [1, 0, 0, 10.0, 0, -1, 0, 0, 0, 0, -1, 5.0, 0, 0, 0, 1]]}
    models(os.path.join(bop, "lmo", "models"), 15, {8: discrete, 10: continuous})
    scene_copy(os.path.join(bop, "lmo", "lmo_test_bop19", "test", "000002"), 1)
    models(os.path.join(bop, "tudl", "tudl_models", "models"), 3, {2: dict(continuous, **discrete)})
    scene_copy(os.path.join(bop, "tudl", "tudl_test_bop19", "test", "000001"), 2)
    models(os.path.join(bop, "icbin", "icbin_models", "models"), 2)
    scene_copy(os.path.join(bop, "icbin", "icbin_test_bop19", "test", "000001"), 2)
    models(os.path.join(bop, "tless", "models_cad"), 30, {4: {"symmetries_continuous": [{"axis": [0, 1, 0], "offset": [1.0, 2.0, 3.0]}]}})
    scene_copy(os.path.join(bop, "tless", "tless_test_primesense_bop19", "test_primesense", "000002"), 2, depth_scale=0.1)
    models(os.path.join(bop, "hb", "hb_models", "models"), 33, {5: {"symmetries_continuous": [{"axis": [1, 0, 0], "offset": [0, 0, 0]}]}})
    scene_copy(os.path.join(bop, "hb", "hb_test_primesense_bop19", "test_primesense", "000003"), 4, with_gt=False)
    targets("hb", 3, [(3, 0, 4, 1), (3, 1, 4, 2), (3, 1, 9, 1), (5, 0, 1, 1)])
    models(os.path.join(bop, "itodd", "itodd_models", "models"), 28)
    scene_copy(os.path.join(bop, "itodd", "itodd_test_bop19", "test", "000001"), 5, grey=True, with_gt=False)
    targets("itodd", 1, [(1, 0, 5, 3), (1, 1, 2, 1)])
    return {"BOP_DIR": bop, "YCB_VIDEO_DIR": ycb}


def collect(dr, root):
    """Everything the reader classes of module `dr` return on the trees of build(root)."""
    out = {}

    def put(key, v):
        if v is None:
            out[key] = np.array("None")
        elif isinstance(v, str):
            out[key] = np.array(v)
        else:
            out[key] = np.asarray(v)

    def exact(key, a):
        if a is None:
            return put(key, None)
        a = np.asarray(a)
        put(key + ".shape", list(a.shape))
        put(key + ".dtype", str(a.dtype))
        put(key + ".sha1", _sha(a))

    def approx(key, a):
        a = np.asarray(a, dtype=np.float64)
        put(key + ".shape", list(a.shape))
        put(key + ".sample", a[::16, ::16])
        put(key + ".sum", a.sum())
        put(key + ".nonzero", int(np.count_nonzero(a)))

    def mesh(key, m):
        exact(key + ".vertices", np.asarray(m.vertices, dtype=np.float64))
        exact(key + ".faces", np.asarray(m.faces, dtype=np.int64))

    # ---- demo scenes
    for name, kw in (("mustard0", {}), ("mustard0", {"downscale": 0.5}), ("mustard0", {"shorter_side": 120, "zfar": 1.0}), ("bleach0", {})):
        r = dr.YcbineoatReader(os.path.join(root, "demo", name), **kw)
        k = "demo." + name + "." + "_".join(f"{a}{b}" for a, b in kw.items())
        put(k + ".len", len(r))
        put(k + ".id_strs", r.id_strs)
        put(k + ".K", r.K)
        put(k + ".HW", [r.H, r.W])
        put(k + ".downscale", r.downscale)
        put(k + ".video_name", r.get_video_name())
        put(k + ".object", r.videoname_to_object[r.get_video_name()])
        for i in range(len(r)):
            exact(f"{k}.color{i}", r.get_color(i))
            approx(f"{k}.depth{i}", r.get_depth(i))
            approx(f"{k}.xyz{i}", r.get_xyz_map(i).reshape(r.H, -1))
            put(f"{k}.gt_pose{i}", r.get_gt_pose(i))
        exact(k + ".mask0", r.get_mask(0))
        if name == "mustard0" and not kw:
            mesh(k + ".gt_mesh", r.get_gt_mesh())

    # ---- BOP-format scenes
    def bop_reader(key, r, mesh_ids=(), mask_frames=True):
        put(key + ".class", type(r).__name__)
        put(key + ".dataset_name", r.dataset_name)
        put(key + ".ob_ids", list(r.ob_ids))
        put(key + ".id_strs", r.id_strs)
        put(key + ".n_color_files", len(r.color_files))
        put(key + ".video_id", r.get_video_id())
        put(key + ".depth_scale", r.bop_depth_scale)
        put(key + ".K_table", np.stack([r.K_table[s] for s in sorted(r.K_table)]))
        if hasattr(r, "K"):
            put(key + ".K", r.K)
        for ob_id in r.ob_ids:
            put(f"{key}.symmetry_tfs.{ob_id}", r.symmetry_tfs[ob_id])
            put(f"{key}.diameter.{ob_id}", r.get_model_diameter(ob_id))
        put(key + ".geometry_symmetry_info", json.dumps({str(k): v for k, v in r.geometry_symmetry_info_table.items()}, sort_keys=True))
        put(key + ".mesh_file", os.path.relpath(os.path.abspath(r.get_gt_mesh_file(r.ob_ids[0])), root))
        for ob_id in mesh_ids:
            mesh(f"{key}.gt_mesh.{ob_id}", r.get_gt_mesh(ob_id))
        for i in range(len(r.color_files)):
            put(f"{key}.K{i}", r.get_K(i))
            exact(f"{key}.color{i}", r.get_color(i))
            approx(f"{key}.depth{i}", r.get_depth(i))
            approx(f"{key}.xyz{i}", r.get_xyz_map(i).reshape(r.get_depth(i).shape[0], -1))
            ids = r.get_instance_ids_in_image(i)
            put(f"{key}.instance_ids{i}", ids)
            if r.scene_gt is None:
                continue
            for ob_id in sorted(set(int(x) for x in ids)) + [9]:
                msk = r.get_mask(i, ob_id)
                exact(f"{key}.mask{i}.{ob_id}", msk)
                exact(f"{key}.mask_full{i}.{ob_id}", r.get_mask(i, ob_id, type="mask"))
                put(f"{key}.gt_poses{i}.{ob_id}", r.get_gt_poses(i, ob_id))
                put(f"{key}.gt_pose{i}.{ob_id}", r.get_gt_pose(i, ob_id))
                if msk is not None:
                    put(f"{key}.gt_pose_by_mask{i}.{ob_id}", r.get_gt_pose(i, ob_id, mask=msk))

    lm = os.path.join(root, "LINEMOD", "lm_test_all", "test")
    r = dr.LinemodReader(os.path.join(lm, "000001"), split=None)
    bop_reader("lm.000001", r, mesh_ids=(1,))
    import cv2

    other = cv2.imread(os.path.join(lm, "000001", "mask_visib", "000000_000002.png"), -1) > 0
    put("lm.000001.gt_pose_second_instance", r.get_gt_pose(0, 1, mask=other))
    bop_reader("lm.000006", dr.LinemodReader(os.path.join(lm, "000006"), zfar=0.62, split=None), mesh_ids=(6,))
    y = dr.YcbVideoReader(os.path.join(root, "YCB_Video", "test", "000049"), zfar=1.5)
    bop_reader("ycbv.000049", y, mesh_ids=(6, 13))
    put("ycbv.000049.names", [y.ob_id_to_names[i] for i in y.ob_ids])
    put("ycbv.000049.name_to_id", [y.name_to_ob_id[n] for n in sorted(y.name_to_ob_id)])
    put("ycbv.000049.keyframes", [bool(y.is_keyframe(i)) for i in range(len(y.color_files))])
    mesh("ycbv.000049.posecnn_mesh", y.get_gt_mesh(6, get_posecnn_version=True))
    put("ycbv.000049.reconstructed_to_gt", y.get_transform_reconstructed_to_gt_model(6))
    bop = os.path.join(root, "bop")
    for dataset, rel, mesh_ids in (("lmo", "lmo/lmo_test_bop19/test/000002", (1,)), ("tudl", "tudl/tudl_test_bop19/test/000001", (2,)),
                                   ("icbin", "icbin/icbin_test_bop19/test/000001", (1,)),
                                   ("tless", "tless/tless_test_primesense_bop19/test_primesense/000002", ()),
                                   ("hb", "hb/hb_test_primesense_bop19/test_primesense/000003", (1,)),
                                   ("itodd", "itodd/itodd_test_bop19/test/000001", (2,))):
        r = dr.get_bop_reader(os.path.join(bop, rel), zfar=5.0 if dataset == "tless" else np.inf)
        bop_reader("bop." + dataset, r, mesh_ids=mesh_ids)
        put(f"bop.{dataset}.video_dirs", [os.path.relpath(p, bop) for p in dr.get_bop_video_dirs(dataset)])
        if dataset == "hb":
            put("bop.hb.gt_pose", r.get_gt_pose(0, 4))
    put("bop.ycbv.video_dirs", list(dr.get_bop_video_dirs("ycbv")))
    put("bop.list", list(dr.BOP_LIST))
    return out


def compare(got, want, rtol=1e-12, atol=1e-12):
    """Differences between two collect() results as a list of strings (empty: equal).  Checksums, shapes, strings and
    integers must be identical; floating-point entries agree to 1e-12 (an ulp of re-associated unit conversions)."""
    bad = [f"missing: {k}" for k in want if k not in got] + [f"unexpected: {k}" for k in got if k not in want]
    for k in want:
        if k not in got:
            continue
        a, b = np.asarray(got[k]), np.asarray(want[k])
        if a.shape != b.shape:
            bad.append(f"{k}: shape {a.shape} vs {b.shape}")
        elif a.dtype.kind in "fc" or b.dtype.kind in "fc":
            if not np.allclose(a.astype(np.float64), b.astype(np.float64), rtol=rtol, atol=atol, equal_nan=True):
                bad.append(f"{k}: max |diff| {np.nanmax(np.abs(a.astype(np.float64) - b.astype(np.float64))):.3g}")
        elif not np.array_equal(a, b):
            bad.append(f"{k}: {a.tolist() if a.size < 12 else '...'} vs {b.tolist() if b.size < 12 else '...'}")
    return bad
