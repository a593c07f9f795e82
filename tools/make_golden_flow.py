"""Golden vectors for the CONTROL FLOW of the hot path, produced by the reference's own, unmodified methods:

    FoundationPose.register / track_one / generate_random_pose_hypo / guess_translation / get_tf_to_centered_mesh
                                                                                   estimater.py:77-85, :127-268
    PoseRefinePredictor.predict                                                    predict_pose_refine.py:148-239
    ScorePredictor.predict (find_best_among_pairs, the +100, the selection loop)   predict_score.py:160-214
    RefineNet / ScoreNetMultiPair                                                  learning/models/*.py (imported as is)

None of those files can be imported here (`from Utils import *` pulls in pytorch3d, nvdiffrast, kornia, open3d, ...), so
the method sources are extracted with `ast` (tools/make_golden_geometry.py::extract) and executed on the CPU with only
these substitutions — everything else is the reference's code:

  * `nvdiffrast_render` (nvdiffrast, ABSENT here) -> the oracle's rasteriser (oracle.raster.render_crop) called with the
    `bbox2d` windows the reference computes; `kornia.geometry.transform.warp_perspective` (kornia, absent) ->
    oracle.geometry.warp_perspective.  These two primitives therefore stay "parity unpinned"; everything AROUND them is
    the reference's code and is what this fixture pins: both `make_crop_data_batch` functions
    (predict_pose_refine.py:24-89, predict_score.py:56-114: crop windows, `bbox2d_ori` through `tf.inverse()`, x255,
    which image is warped how, BatchPoseData assembly), `PairH5Dataset` / `TripletH5Dataset.transform_batch` +
    `transform_depth_to_xyzmap` (h5_dataset.py:79-127, :137-179, incl. the scorer's depth round trip), batching,
    concatenation order, autocast region, tanh / normaliser / axis-angle / transpose, `trans_delta *= d/2`, the
    egocentric pose update, the iteration feed-back, last_*_update, the scorer's `L = len(A)` call, +100,
    argsort(descending), `poses[ids]`, `best @ T(-model_center)`, pose_last, the < 4 valid pixels early-out,
    track_one's `depth2xyzmap_batch(zfar=inf)` and single-pose refine.
  * `so3_exp_map` (pytorch3d, absent) -> oracle.geometry.so3_exp_map (held to scipy in
    tests/test_oracle_thirdparty_cpu.py).
  * `erode_depth` / `bilateral_filter_depth` (Warp launches) -> the oracle's filters, which
    tests/test_geometry_golden_cpu.py holds to the reference's Warp kernel bodies.
  * `.cuda()` / `device='cuda'` / `torch.set_default_tensor_type` are neutralised (run where we are); the intrinsics enter
    `compute_crop_window_tf_batch` as float32 (its `torch.as_tensor(K)` must meet float32 points);
    `torch.cuda.amp.autocast` disables itself without a CUDA device, so the networks run in fp32 like the oracle.

    python tools/make_golden_flow.py        # needs /root/reference; writes tests/golden/flow_golden.npz (~1 min)

tests/test_flow_golden_cpu.py holds oracle.pipeline.register / track_one — the oracle the GPU parity tests compare the
CUDA path with — to these vectors.
"""
import logging
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
REF = os.environ.get("FPOSE_REFERENCE", "/root/reference")

EXACT_TIES = [True]  # see warp_perspective() in reference_objects()
N_HYP = 5          # hypotheses of the register case (every 50th pose of the 252-pose grid, CPU-sized)
REGISTER_ITERS = 2
TRACK_ITERS = 2


class Cfg(dict):
    __getattr__ = dict.__getitem__


class _TorchProxy(types.ModuleType):
    """`torch` whose factory functions ignore `device=` and whose default-tensor-type switch does nothing."""

    def __getattr__(self, k):
        v = getattr(torch, k)
        if k == "set_default_tensor_type":
            return lambda *a, **kw: None
        if callable(v) and not isinstance(v, (type, types.ModuleType)) and k in ("as_tensor", "tensor", "eye", "zeros", "ones", "arange", "empty"):
            def f(*a, _v=v, **kw):
                kw.pop("device", None)
                return _v(*a, **kw)
            return f
        return v


def extract_class(path, name):
    """Source-exact ClassDef `name` of a reference file, compiled on its own (decorators dropped: BatchPoseData is a
    @dataclass only in name — it defines its own __init__)."""
    import ast

    tree = ast.parse(open(path).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == name)
    cls.decorator_list = []
    for n in ast.walk(cls):
        if isinstance(n, ast.AnnAssign):  # `rgbs: torch.Tensor = None` -> plain class attributes
            n.annotation = ast.Constant(value=None)
        if isinstance(n, ast.FunctionDef):
            n.returns = None
            for a in n.args.args:
                a.annotation = None
    return compile(ast.fix_missing_locations(ast.Module(body=[cls], type_ignores=[])), f"{path}:{name}", "exec")


def scene():
    """The smoke scene: icosphere-2 ellipsoid (320 faces) in front of a textured plane, 640 x 480."""
    from foundationpose_b200 import synth
    from oracle import pipeline

    mesh = synth.make_mesh(2)
    pose = np.eye(4)
    pose[:3, :3] = synth.random_rotation(0)
    pose[:3, 3] = [0.01, -0.01, 0.6]
    rgb, depth, mask = synth.make_scene(mesh.visual.image, pose)
    center = (mesh.vertices.max(0) + mesh.vertices.min(0)) / 2 + np.array([0.004, -0.002, 0.003])  # a non-trivial model_center
    return mesh, pipeline.mesh_tensors(mesh), pose, rgb, depth, mask, synth.DEFAULT_K.copy(), synth.mesh_diameter(mesh.vertices), center


def start_grid():
    from foundationpose_b200 import hypotheses

    return hypotheses.make_rotation_grid()[::50][:N_HYP].astype(np.float32)


def reference_objects(sd_r, sd_s, cfg):
    """(estimator, refiner, scorer) built from the reference's method sources + the substitutions listed above."""
    from make_golden import import_reference_models
    from make_golden_geometry import extract
    from make_golden_geometry import _TorchProxy as _GeomTorchProxy  # noqa: F401  (same neutralisation, kept for reference)
    from oracle import geometry, raster

    RefineNet, ScoreNet = import_reference_models()
    tp = _TorchProxy("torch")
    torch.Tensor.cuda = lambda self, *a, **k: self
    if not torch.cuda.is_available():
        torch.cuda.manual_seed_all = lambda *a, **k: None

    def nvdiffrast_render(K=None, H=None, W=None, ob_in_cams=None, glctx=None, context="cuda", get_normal=False, mesh_tensors=None, mesh=None,
                          projection_mat=None, bbox2d=None, output_size=None, use_light=False, light_color=None, light_dir=None, light_pos=None,
                          w_ambient=0.8, w_diffuse=0.5, extra={}):
        """Utils.py:133-219 by the oracle's rasteriser: colour 0..1 (N,S,S,3), depth (N,S,S), no normals; extra['xyz_map']."""
        assert use_light and not get_normal and tuple(output_size) == (160, 160) and bbox2d is not None
        cols, xyzs = [], []
        for pose, box in zip(ob_in_cams.numpy(), bbox2d.numpy()):
            rgb, xyz, _ = raster.render_crop(pose, mesh_tensors, K, tuple(np.float32(v) for v in box), w_ambient=w_ambient, w_diffuse=w_diffuse)
            cols.append(rgb)
            xyzs.append(xyz)
        xyz = torch.from_numpy(np.stack(xyzs))
        extra["xyz_map"] = xyz
        return torch.from_numpy(np.stack(cols)), xyz[..., 2].clone(), None

    def warp_perspective(src, M, dsize, mode="bilinear", align_corners=False, **kw):
        assert align_corners is False and not kw
        if EXACT_TIES[0] and mode == "nearest" and tuple(dsize) != (160, 160):
            # The scorer's crop -> full-resolution warp (h5_dataset.py:158).  The crop window has integer edges, so the
            # first crop row / column maps back to EXACTLY -0.5: a rounding tie that kornia's op sequence decides by the
            # last bit of a 3x3 LU inverse (implementation-defined: CPU LAPACK and cuSOLVER differ; on this CPU the column
            # vanishes for this scene).  Default here: the tie as exact arithmetic resolves it (oracle.geometry.
            # unwarp_nearest, what the CUDA kernel evaluates); the op-sequence result is recorded next to it.
            Mn = M.float().numpy()
            w, h = np.rint(Mn[:, 0, 0] * 160), np.rint(Mn[:, 1, 1] * 160)
            f32 = np.float32
            win = dict(left=np.rint(Mn[:, 0, 2]).astype(f32), top=np.rint(Mn[:, 1, 2]).astype(f32),
                       sx=((f32(1) / w.astype(f32)).astype(f32) * f32(160)).astype(f32), sy=((f32(1) / h.astype(f32)).astype(f32) * f32(160)).astype(f32))
            return geometry.unwarp_nearest(src.float().contiguous(), win, tuple(dsize))
        return geometry.warp_perspective(src.float().contiguous(), M.float(), tuple(dsize), mode)

    kornia = types.SimpleNamespace(geometry=types.SimpleNamespace(transform=types.SimpleNamespace(warp_perspective=warp_perspective)))

    def filt(fn):
        def wrapped(depth, radius=2, device="cuda", **kw):
            is_t = torch.is_tensor(depth)
            out = fn(depth.numpy() if is_t else depth, radius=radius, **kw)
            return torch.from_numpy(out) if is_t else out
        return wrapped

    base = {"np": np, "torch": tp, "logging": logging, "os": os, "so3_exp_map": geometry.so3_exp_map, "kornia": kornia,
            "nvdiffrast_render": nvdiffrast_render,
            "erode_depth": filt(geometry.erode_depth), "bilateral_filter_depth": filt(geometry.bilateral_filter_depth),
            "dr": types.SimpleNamespace(RasterizeCudaContext=lambda *a, **k: "glctx")}
    for name in ("set_seed", "depth2xyzmap", "depth2xyzmap_batch", "egocentric_delta_pose_to_pose", "compute_crop_window_tf_batch", "transform_pts"):
        exec(extract(os.path.join(REF, "Utils.py"), name), base)
    _ccw = base["compute_crop_window_tf_batch"]
    # `K = torch.as_tensor(K)` (Utils.py:610) must meet float32 points: the drivers' float64 intrinsics enter as float32
    # here, exactly as in tools/make_golden_geometry.py, whose vectors pin the oracle's crop window bit for bit
    base["compute_crop_window_tf_batch"] = lambda *a, K=None, **kw: _ccw(*a, K=np.asarray(K, dtype=np.float32), **kw)
    exec(extract_class(os.path.join(REF, "learning/datasets/pose_dataset.py"), "BatchPoseData"), base)
    base["PoseRefinePairH5Dataset"] = base["TripletH5Dataset"] = object  # annotations of the two make_crop_data_batch

    def dataset(cls, cfg):
        ns = dict(base)
        for name in ("transform_batch", "transform_depth_to_xyzmap"):
            exec(extract(os.path.join(REF, "learning/datasets/h5_dataset.py"), name, cls=cls), ns)
        d = type("Ref" + cls, (), {"transform_batch": ns["transform_batch"], "transform_depth_to_xyzmap": ns["transform_depth_to_xyzmap"]})()
        d.cfg = cfg
        return d

    ns_r = dict(base)
    exec(extract(os.path.join(REF, "learning/training/predict_pose_refine.py"), "make_crop_data_batch"), ns_r)
    exec(extract(os.path.join(REF, "learning/training/predict_pose_refine.py"), "predict", cls="PoseRefinePredictor"), ns_r)
    ns_s = dict(base)
    exec(extract(os.path.join(REF, "learning/training/predict_score.py"), "make_crop_data_batch"), ns_s)
    exec(extract(os.path.join(REF, "learning/training/predict_score.py"), "predict", cls="ScorePredictor"), ns_s)
    ns_e = dict(base)
    methods = {}
    for name in ("register", "track_one", "generate_random_pose_hypo", "guess_translation", "get_tf_to_centered_mesh", "compute_add_err_to_gt_pose"):
        exec(extract(os.path.join(REF, "estimater.py"), name, cls="FoundationPose"), ns_e)
        methods[name] = ns_e[name]

    net_cfg = Cfg(use_BN=True, rot_rep="axis_angle")
    model_r = RefineNet(cfg=net_cfg, c_in=6).eval()
    model_r.load_state_dict(sd_r, strict=True)
    model_s = ScoreNet(cfg=net_cfg, c_in=6).eval()
    model_s.load_state_dict(sd_s, strict=True)
    refiner = type("RefPoseRefinePredictor", (), {"predict": ns_r["predict"]})()
    refiner.cfg, refiner.amp, refiner.model, refiner.dataset = cfg, True, model_r, dataset("PairH5Dataset", cfg)
    refiner.make_crop_data_batch = staticmethod(ns_r["make_crop_data_batch"])
    scorer = type("RefScorePredictor", (), {"predict": ns_s["predict"]})()
    scorer.cfg, scorer.amp, scorer.model, scorer.dataset = cfg, True, model_s, dataset("TripletH5Dataset", cfg)
    scorer.make_crop_data_batch = staticmethod(ns_s["make_crop_data_batch"])
    est = type("RefFoundationPose", (), methods)()
    est.refiner, est.scorer, est.glctx, est.debug, est.debug_dir, est.pose_last = refiner, scorer, None, 0, "/tmp", None
    return est, refiner, scorer


def main():
    from foundationpose_b200.weights import DEFAULT_CFG, random_state_dict

    torch.set_num_threads(os.cpu_count())
    mesh, mt, gt, rgb, depth, mask, K, d, center = scene()
    sd_r, sd_s = random_state_dict("refine", 0), random_state_dict("score", 0)
    # OmegaConf's ListConfig compares equal to a torch.Size (`rgb_rs.shape[-2:] != cfg['input_resize']`,
    # predict_pose_refine.py:66); a plain list would not, a tuple does
    cfg = Cfg(DEFAULT_CFG, input_resize=(160, 160))
    est, refiner, scorer = reference_objects(sd_r, sd_s, cfg)
    est.mesh, est.mesh_tensors, est.diameter, est.model_center = mesh, mt, d, center
    est.rot_grid = torch.from_numpy(start_grid())
    out = {}
    with torch.inference_mode():
        # ---- register (estimater.py:159-240)
        best = est.register(K=K, rgb=rgb, depth=depth, ob_mask=mask, iteration=REGISTER_ITERS)
        out.update(reg_best_pose=best, reg_pose_last=est.pose_last.numpy(), reg_best_id=np.int64(est.best_id), reg_poses_sorted=est.poses.numpy(),
                   reg_scores_sorted=est.scores.numpy(), reg_last_trans=refiner.last_trans_update.numpy(), reg_last_rot=refiner.last_rot_update.numpy())
        s = est.scores.numpy()
        print(f"register: best id {int(est.best_id)}, scores {s}, top-2 margin {s[0] - s[1]:.4f}")
        # the same register with the inverse warp's tie decided by kornia's op sequence on THIS machine (sensitivity only)
        EXACT_TIES[0] = False
        est.register(K=K, rgb=rgb, depth=depth, ob_mask=mask, iteration=REGISTER_ITERS)
        out.update(opseq_best_id=np.int64(est.best_id), opseq_scores_sorted=est.scores.numpy())
        print(f"   op-sequence ties: best id {int(est.best_id)}, scores {est.scores.numpy()}")
        EXACT_TIES[0] = True
        est.register(K=K, rgb=rgb, depth=depth, ob_mask=mask, iteration=REGISTER_ITERS)
        # ---- the early-out (estimater.py:181-189): a mask with 3 valid pixels, and an empty one
        tiny = np.zeros_like(mask)
        ys, xs = np.nonzero(mask)
        tiny[ys[:3], xs[:3]] = True
        keep = (est.pose_last.clone(), est.best_id)
        out["early_out_3px"] = est.register(K=K, rgb=rgb, depth=depth, ob_mask=tiny, iteration=REGISTER_ITERS)
        out["early_out_empty"] = est.register(K=K, rgb=rgb, depth=depth, ob_mask=np.zeros_like(mask), iteration=REGISTER_ITERS)
        assert torch.equal(est.pose_last, keep[0]) and est.best_id == keep[1]  # the early return leaves the state alone
        # ---- track_one (estimater.py:250-268) on the next frame of a short sequence, twice (pose fed back)
        from foundationpose_b200 import synth

        seq = synth.track_sequence(3, gt)
        frames = [synth.make_scene(mesh.visual.image, p, seed=11 + i) for i, p in enumerate(seq[1:])]
        for i, (rgb_i, depth_i, _) in enumerate(frames):
            out[f"track_pose{i}"] = est.track_one(rgb=rgb_i, depth=depth_i, K=K, iteration=TRACK_ITERS)
            out[f"track_pose_last{i}"] = est.pose_last.numpy().copy()
            out[f"track_last_trans{i}"] = refiner.last_trans_update.numpy().copy()
        out["track_gt"] = np.stack(seq[1:])
    # ---- the crops themselves, straight from the reference's two make_crop_data_batch + dataset.transform_batch
    import hashlib

    from oracle import geometry

    sha = lambda t: np.frombuffer(hashlib.sha1(np.ascontiguousarray(t.numpy()).tobytes()).digest(), dtype=np.uint8).copy()
    depth_f = geometry.bilateral_filter_depth(geometry.erode_depth(depth))
    poses = start_grid().copy()
    poses[:, :3, 3] = geometry.guess_translation(depth_f, mask, K)
    xyz_map = geometry.depth2xyzmap(depth_f, K)
    as_t = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32)
    with torch.inference_mode():
        pr = refiner.make_crop_data_batch(cfg.input_resize, as_t(poses), mesh, as_t(rgb), as_t(depth_f), K, crop_ratio=cfg["crop_ratio"], xyz_map=as_t(xyz_map),
                                          cfg=cfg, glctx=None, mesh_tensors=mt, dataset=refiner.dataset, mesh_diameter=d)
        ps = scorer.make_crop_data_batch(cfg.input_resize, as_t(poses), mesh, as_t(rgb), as_t(depth_f), K, crop_ratio=cfg["crop_ratio"], cfg=cfg, glctx=None,
                                         mesh_tensors=mt, dataset=scorer.dataset, mesh_diameter=d)
    out.update(crop_poses=poses, crop_refine_rgbB_sha1=sha(pr.rgbBs), crop_refine_xyzB_sha1=sha(pr.xyz_mapBs), crop_score_rgbB_sha1=sha(ps.rgbBs),
               crop_score_xyzB_sha1=sha(ps.xyz_mapBs), crop_refine_A=torch.cat([pr.rgbAs, pr.xyz_mapAs], 1).numpy(),
               crop_score_xyzA=ps.xyz_mapAs.numpy())
    # ---- track-before-register raises (estimater.py:251-253)
    est2, _, _ = reference_objects(sd_r, sd_s, cfg)
    try:
        est2.track_one(rgb=rgb, depth=depth, K=K, iteration=1)
        raised = False
    except RuntimeError:
        raised = True
    out["track_before_register_raises"] = np.array(raised)
    dst = os.path.join(ROOT, "tests", "golden", "flow_golden.npz")
    np.savez_compressed(dst, **out)
    print(f"wrote {dst}: {len(out)} entries")


def headline():
    """The reference's own `register` over the WHOLE headline configuration — 252 start poses x 5 refine iterations +
    scoring + ranking on the scene of tools/make_golden_register.py — written to
    tests/golden/register_252x5_reference_flow.npz (~15 min on 8 cores: the crops come from the Python rasteriser).
    tests/test_flow_golden_cpu.py compares the committed oracle golden the GPU test uses (register_252x5.npz) with it."""
    import time

    import make_golden_register as mgr
    from foundationpose_b200.weights import DEFAULT_CFG, random_state_dict

    torch.set_num_threads(os.cpu_count())
    mesh, mt, gt, rgb, depth, mask, K, d = mgr.scene()
    gold = dict(np.load(os.path.join(ROOT, "tests", "golden", "register_252x5.npz")))
    cfg = Cfg(DEFAULT_CFG, input_resize=(160, 160))
    est, refiner, scorer = reference_objects(random_state_dict("refine", 0), random_state_dict("score", 0), cfg)
    est.mesh, est.mesh_tensors, est.diameter, est.model_center = mesh, mt, d, np.zeros(3)
    est.rot_grid = torch.from_numpy(gold["start"].copy())  # register() overwrites the translation column itself
    seen = {}
    inner = scorer.predict

    def recording_predict(*a, **kw):  # the scorer's input poses and its UNSORTED scores (register only keeps them sorted)
        scores, vis = inner(*a, **kw)
        seen["poses"], seen["scores"] = np.asarray(kw["ob_in_cams"]).copy(), scores.numpy().copy()
        return scores, vis

    scorer.predict = recording_predict
    t0 = time.time()
    with torch.inference_mode():
        best = est.register(K=K, rgb=rgb, depth=depth, ob_mask=mask, iteration=5)
    print(f"reference register over 252 x 5: {time.time() - t0:.0f} s; best id {int(est.best_id)}")
    out = dict(best_pose=best, best_id=np.int64(est.best_id), poses=seen["poses"], scores=seen["scores"], scores_sorted=est.scores.numpy(),
               last_trans=refiner.last_trans_update.numpy(), last_rot=refiner.last_rot_update.numpy())
    dst = os.path.join(ROOT, "tests", "golden", "register_252x5_reference_flow.npz")
    np.savez_compressed(dst, **out)
    print(f"wrote {dst}")
    dp = np.abs(out["poses"] - gold["poses"][5]).max()
    ds = out["scores"] - gold["scores"]
    print(f"vs the oracle golden: final poses max |diff| {dp:.2e}; scores offset {ds.mean():+.4f}, rank-relevant {np.abs(ds - ds.mean()).max():.4f}; "
          f"best {int(est.best_id)} vs {int(gold['best'][0])}; argsort equal: {np.array_equal(np.argsort(-out['scores'], kind='stable'), gold['ids'])}")


def track():
    """The reference's own `track_one` over the 49 tracked frames (and the 5-frame chain) of tools/make_golden_track.py
    -> tests/golden/track_seq_reference_flow.npz; compared with the committed oracle golden (track_seq.npz) by
    tests/test_flow_golden_cpu.py."""
    from foundationpose_b200 import synth
    from foundationpose_b200.weights import DEFAULT_CFG, random_state_dict

    torch.set_num_threads(os.cpu_count())
    gold = dict(np.load(os.path.join(ROOT, "tests", "golden", "track_seq.npz")))
    mesh = synth.make_mesh(3)
    from oracle import pipeline

    mt, d, K = pipeline.mesh_tensors(mesh), synth.mesh_diameter(mesh.vertices), synth.DEFAULT_K
    cfg = Cfg(DEFAULT_CFG, input_resize=(160, 160))
    est, refiner, scorer = reference_objects(random_state_dict("refine", 0), random_state_dict("score", 0), cfg)
    est.mesh, est.mesh_tensors, est.diameter, est.model_center = mesh, mt, d, np.zeros(3)
    gt = gold["gt"]
    frame = lambda i: synth.make_scene(mesh.visual.image, gt[i], seed=1 + i)[:2]
    pose_out, lt = [], []
    with torch.inference_mode():
        for i in range(1, len(gt)):
            rgb, depth = frame(i)
            est.pose_last = torch.from_numpy(gold["pose_in"][i - 1].copy())
            pose_out.append(est.track_one(rgb=rgb, depth=depth, K=K, iteration=2))
            lt.append(refiner.last_trans_update.numpy()[0].copy())
        est.pose_last = torch.from_numpy(gold["chain"][0].copy())
        chain = [gold["chain"][0]]
        for i in range(1, 6):
            rgb, depth = frame(i)
            chain.append(est.track_one(rgb=rgb, depth=depth, K=K, iteration=2))
    out = dict(pose_out=np.stack(pose_out), last_trans=np.stack(lt), chain=np.stack(chain))
    dst = os.path.join(ROOT, "tests", "golden", "track_seq_reference_flow.npz")
    np.savez_compressed(dst, **out)
    print(f"wrote {dst}; vs the oracle golden: poses max |diff| {np.abs(out['pose_out'] - gold['pose_out']).max():.2e}, "
          f"last_trans {np.abs(out['last_trans'] - gold['last_trans']).max():.2e}, chain {np.abs(out['chain'] - gold['chain']).max():.2e}")


if __name__ == "__main__":
    if "--headline" in sys.argv:
        headline()
    elif "--track" in sys.argv:
        track()
    else:
        main()
