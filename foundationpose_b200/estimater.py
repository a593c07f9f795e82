"""Host-side mirror of the reference's estimator and predictor surfaces, running on libfpose.so.

Same names, argument orders and error behaviour as
  estimater.py:18-268                              FoundationPose {register, track_one, reset_object, ...}
  learning/training/predict_pose_refine.py:92-239  PoseRefinePredictor {predict, last_trans_update, ...}
  learning/training/predict_score.py:117-226       ScorePredictor {predict}
so a driver written against the reference (run_demo.py:38-63) only changes its import.

Differences that are deliberate: no nvdiffrast context is needed (`glctx` is accepted
and ignored) and visualisation (`get_vis`, debug >= 2 dumps) is not produced.  `predict()` honours the
per-call `mesh` / `mesh_tensors` / `mesh_diameter` / `xyz_map` arguments of the reference: a mesh that is not
the one already in the context is uploaded first, a caller-supplied xyz map replaces the derived one.
"""
import logging
import os

import numpy as np
import torch

from . import hypotheses, meshprep, synth, weights
from .engine import Engine


class _Cfg(dict):
    __getattr__ = dict.__getitem__

    def __setattr__(self, k, v):
        self[k] = v


class _ModelHandle:
    """Stands in for the `.model` attribute the reference estimator moves between devices
    (estimater.py:97-100); the packed weights live inside the fp_ctx."""

    def to(self, *a, **k):
        return self

    def cuda(self):
        return self

    def eval(self):
        return self


_shared_engine = None


def get_engine():
    """One fp_ctx per process/device, shared by the scorer and the refiner (they share mesh + frame)."""
    global _shared_engine
    if _shared_engine is None:
        _shared_engine = Engine()
    return _shared_engine


def _load_cfg_and_weights(run_name, kind, state_dict, cfg):
    """Config + weights the way the reference predictors find them (predict_pose_refine.py:96-131,
    predict_score.py:120-143): weights/<run>/config.yml next to model_best.pth, then the reference's
    backward-compatibility defaults for missing keys.  Without a checkpoint tree (this repository ships none)
    the released-config values of weights.DEFAULT_CFG and the seeded stand-in weights are used."""
    path = weights.find_reference_weights(run_name) if state_dict is None else None
    if path is not None:
        c = _Cfg(weights.load_reference_config(os.path.join(os.path.dirname(path), "config.yml"), kind))
        state_dict = weights.load_checkpoint(path)
        logging.info(f"Using pretrained model from {path}")
    else:
        c = _Cfg(weights.DEFAULT_CFG)
        if state_dict is None:
            logging.info(f"weights/{run_name}/model_best.pth not found: using the seeded synthetic {kind} weights")
            state_dict = weights.random_state_dict(kind, seed=0)
    if cfg:
        unknown = set(cfg) - set(weights.DEFAULT_CFG) - {"ckpt_dir", "enable_amp", "use_mask", "n_view", "normal_uint8"}
        if unknown:
            raise ValueError(f"unsupported config keys for the {kind} predictor: {sorted(unknown)}")
        c.update(cfg)
    if not c["normalize_xyz"] or c["rot_rep"] != "axis_angle" or c["trans_rep"] != "tracknet" or c["use_normal"]:
        raise NotImplementedError("engine supports the released configs: normalize_xyz, tracknet, axis_angle, no normals")
    if list(c["input_resize"]) != [160, 160] or int(c["c_in"]) != 6:
        raise NotImplementedError("engine supports input_resize = [160, 160] and c_in = 6 (the released configs)")
    return c, state_dict


def _mesh_arrays(mesh_tensors):
    """Accepts this package's make_mesh_tensors() dict or the reference's (Utils.py:104-130: torch tensors under
    'pos', 'faces', 'vnormals', 'uv', 'tex' [1,H,W,3 float 0..1] or 'vertex_color')."""
    g = lambda k: mesh_tensors.get(k)
    to_np = lambda t: t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)
    out = dict(pos=to_np(g("pos")), faces=to_np(g("faces")), normals=to_np(g("normals") if g("normals") is not None else g("vnormals")))
    if g("uv") is not None and g("tex") is not None:
        tex = to_np(g("tex"))
        if tex.ndim == 4:
            tex = tex[0]
        if tex.dtype != np.uint8:  # the reference keeps the texture as float 0..1
            tex = np.clip(np.rint(tex * 255.0), 0, 255).astype(np.uint8)
        out["uv"], out["tex"] = to_np(g("uv")), tex
    else:
        vc = g("vcolor") if g("vcolor") is not None else g("vertex_color")
        out["vcolor"] = to_np(vc)
    return out


def _sync_mesh(engine, mesh, mesh_tensors, mesh_diameter):
    """predict()'s per-call mesh arguments (predict_pose_refine.py:171-172, predict_score.py:176-177): upload the
    mesh unless it is the one the context already holds."""
    src = mesh_tensors if mesh_tensors is not None else mesh
    if src is None:
        if engine.diameter is None:
            raise ValueError("predict(): no mesh in the context and none passed (mesh / mesh_tensors)")
        return
    d = float(mesh_diameter) if mesh_diameter is not None else engine.diameter
    if getattr(engine, "_mesh_src", None) is src and (d is None or d == engine.diameter):
        return
    mt = _mesh_arrays(mesh_tensors) if mesh_tensors is not None else make_mesh_tensors(mesh)
    if d is None:
        d = synth.mesh_diameter(mt["pos"])
    engine.set_mesh(mt["pos"], mt["normals"], mt["faces"], d, uv=mt.get("uv"), tex=mt.get("tex"), vertex_colors=mt.get("vcolor"))
    engine._mesh_src = src


class PoseRefinePredictor:
    def __init__(self, engine=None, state_dict=None, cfg=None):
        self.amp = True
        self.run_name = "2023-10-28-18-33-37"
        self.engine = engine or get_engine()
        self.cfg, sd = _load_cfg_and_weights(self.run_name, "refine", state_dict, cfg)
        self.engine.load_network("refine", sd)
        self.engine.set_config("refine", self.cfg["crop_ratio"], self.cfg["rot_normalizer"])
        self.model = _ModelHandle()
        self.dataset = None
        self.last_trans_update = None
        self.last_rot_update = None

    @torch.inference_mode()
    def predict(self, rgb, depth, K, ob_in_cams, xyz_map=None, normal_map=None, get_vis=False, mesh=None, mesh_tensors=None,
                glctx=None, mesh_diameter=None, iteration=5, _frame_ready=False):
        """@rgb: (H,W,3) uint8; @ob_in_cams: (N,4,4).  Returns ((N,4,4) cuda tensor, None)."""
        e = self.engine
        _sync_mesh(e, mesh, mesh_tensors, mesh_diameter)
        if not _frame_ready:
            e.set_frame(rgb, depth, K, filter_depth=False, zfar=float("inf"))
            if xyz_map is not None:
                e.set_xyz_map(xyz_map)
        poses, lt, lr = e.refine(ob_in_cams, iteration)
        self.last_trans_update = lt
        self.last_rot_update = lr
        return poses, None


class ScorePredictor:
    def __init__(self, amp=True, engine=None, state_dict=None, cfg=None):
        self.amp = amp
        self.run_name = "2024-01-11-20-02-45"
        self.engine = engine or get_engine()
        self.cfg, sd = _load_cfg_and_weights(self.run_name, "score", state_dict, cfg)
        self.engine.load_network("score", sd)
        self.engine.set_config("score", self.cfg["crop_ratio"])
        self.model = _ModelHandle()
        self.dataset = None

    @torch.inference_mode()
    def predict(self, rgb, depth, K, ob_in_cams, normal_map=None, get_vis=False, mesh=None, mesh_tensors=None, glctx=None,
                mesh_diameter=None, _frame_ready=False):
        """Returns ((N,) cuda tensor of scores = logits + 100, None)."""
        e = self.engine
        _sync_mesh(e, mesh, mesh_tensors, mesh_diameter)
        if not _frame_ready:
            e.set_frame(rgb, depth, K, filter_depth=False)
        scores, _ = e.score(ob_in_cams)
        return scores, None


def make_mesh_tensors(mesh):
    """Utils.py:104-130 on a trimesh-like object -> host arrays for fp_set_mesh."""
    out = dict(pos=np.asarray(mesh.vertices, dtype=np.float32), faces=np.asarray(mesh.faces, dtype=np.int32),
               normals=np.asarray(mesh.vertex_normals, dtype=np.float32))
    uv = getattr(mesh.visual, "uv", None)
    img = getattr(mesh.visual, "image", None)
    if img is None and getattr(mesh.visual, "material", None) is not None:
        img = np.asarray(mesh.visual.material.image.convert("RGB"))
    if uv is not None and img is not None:
        uv = np.asarray(uv, dtype=np.float32).copy()
        uv[:, 1] = 1 - uv[:, 1]
        out["uv"] = uv
        out["tex"] = np.ascontiguousarray(np.asarray(img)[..., :3], dtype=np.uint8)
    else:
        vc = getattr(mesh.visual, "vertex_colors", None)
        if vc is None:
            vc = np.tile(np.array([128, 128, 128]).reshape(1, 3), (len(mesh.vertices), 1))
        out["vcolor"] = np.asarray(vc, dtype=np.float32)[..., :3] / 255.0
    return out


class FoundationPose:
    def __init__(self, model_pts, model_normals, symmetry_tfs=None, mesh=None, scorer=None, refiner=None, glctx=None, debug=0,
                 debug_dir="/tmp/fpose_b200_debug"):
        self.gt_pose = None
        self.ignore_normal_flip = True
        self.debug = debug
        self.debug_dir = debug_dir
        os.makedirs(debug_dir, exist_ok=True)
        self.engine = (refiner.engine if refiner is not None else (scorer.engine if scorer is not None else get_engine()))
        self.reset_object(model_pts, model_normals, symmetry_tfs=symmetry_tfs, mesh=mesh)
        self.make_rotation_grid(min_n_views=40, inplane_step=60)
        self.glctx = glctx
        self.scorer = scorer if scorer is not None else ScorePredictor(engine=self.engine)
        self.refiner = refiner if refiner is not None else PoseRefinePredictor(engine=self.engine)
        self.pose_last = None  # used for tracking; per the centred mesh
        # True: read (tx, ty, tz, n_valid) back before the 252 x K loop so that an empty / depth-less mask returns
        # without running it (the reference's control flow, one extra sync); False: sync-free, the loop runs and its
        # result is discarded in that case
        self.strict_early_out = bool(debug)

    def reset_object(self, model_pts, model_normals, symmetry_tfs=None, mesh=None):
        max_xyz = mesh.vertices.max(axis=0)
        min_xyz = mesh.vertices.min(axis=0)
        self.model_center = (min_xyz + max_xyz) / 2
        self.mesh_ori = mesh.copy()
        mesh = mesh.copy()
        mesh.vertices = mesh.vertices - self.model_center.reshape(1, 3)
        self.diameter = synth.mesh_diameter(mesh.vertices)
        self.vox_size = max(self.diameter / 20.0, 0.003)
        self.dist_bin = self.vox_size / 2
        self.angle_bin = 20  # deg
        # estimater.py:59-64: voxel-down-sampled model points / normals (used by callers for ADD-style metrics)
        pts, nrm = meshprep.voxel_down_sample(mesh.vertices, self.vox_size, normals=model_normals)
        self.max_xyz = pts.max(axis=0)
        self.min_xyz = pts.min(axis=0)
        self.pts = torch.tensor(pts, dtype=torch.float32, device="cuda")
        self.normals = torch.nn.functional.normalize(torch.tensor(nrm, dtype=torch.float32, device="cuda"), dim=-1)
        self.mesh_path = None  # the reference exports a temporary .obj for its debug tooling; not needed here
        self.mesh = mesh
        self.mesh_tensors = make_mesh_tensors(mesh)
        mt = self.mesh_tensors
        self.engine.set_mesh(mt["pos"], mt["normals"], mt["faces"], self.diameter, uv=mt.get("uv"), tex=mt.get("tex"),
                             vertex_colors=mt.get("vcolor"))
        self.engine._mesh_src = self.mesh_tensors
        if symmetry_tfs is None:
            self.symmetry_tfs = np.eye(4, dtype=np.float32)[None]
        else:
            self.symmetry_tfs = np.asarray(symmetry_tfs, dtype=np.float32)

    def get_tf_to_centered_mesh(self):
        tf = torch.eye(4, dtype=torch.float32, device="cuda")
        tf[:3, 3] = -torch.as_tensor(self.model_center, device="cuda", dtype=torch.float32)
        return tf

    def to_device(self, s="cuda:0"):
        """estimater.py:88-102 moves tensors / modules / the raster context to `s`.  Here the packed weights, the
        mesh and the frame live in the fp_ctx, which is bound to ONE device for its lifetime: moving to the
        context's own device is a no-op, anything else must be a new Engine on that device (see
        foundationpose_b200.replicas for one estimator per GPU)."""
        dev = torch.device(s)
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        if dev.type != "cuda" or idx != self.engine.device_index:
            raise RuntimeError(f"to_device({s!r}): this estimator's fp_ctx lives on cuda:{self.engine.device_index}; "
                               "create the Engine / FoundationPose under torch.cuda.device(...) of the target GPU instead")
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v):
                self.__dict__[k] = v.to(s)
        return self

    def make_rotation_grid(self, min_n_views=40, inplane_step=60):
        rot_grid = hypotheses.make_rotation_grid(min_n_views, inplane_step, self.symmetry_tfs)
        self.rot_grid = torch.as_tensor(rot_grid, device="cuda", dtype=torch.float32)
        self._rot_grid_host = torch.from_numpy(rot_grid.copy()).pin_memory()

    def generate_random_pose_hypo(self, K, rgb, depth, mask, scene_pts=None):
        ob_in_cams = self.rot_grid.clone()
        center = self.guess_translation(depth=depth, mask=mask, K=K)
        ob_in_cams[:, :3, 3] = torch.tensor(center, device="cuda", dtype=torch.float32).reshape(1, 3)
        return ob_in_cams

    def guess_translation(self, depth, mask, K):
        return hypotheses.guess_translation(depth, mask, K)

    def register(self, K, rgb, depth, ob_mask, ob_id=None, glctx=None, iteration=5):
        """Compute the object pose in the frame (estimater.py:159-240). Returns (4,4) numpy."""
        e = self.engine
        # erode_depth + bilateral_filter_depth + depth2xyzmap on the device (estimater.py:173-174, :214), then the
        # translation guess and the start poses, also on the device: nothing synchronises until the result is read
        e.set_frame(rgb, depth, K, filter_depth=True, zfar=float("inf"))
        poses, info = e.start_poses(ob_mask, self.rot_grid)
        if self.strict_early_out:
            # estimater.py:181-189 returns BEFORE the hypothesis loop when fewer than 4 valid masked pixels remain;
            # knowing that on the host costs one 16-byte read-back (a synchronisation the sync-free path avoids)
            head = info.cpu().numpy()
            if head[3] < 4:
                logging.info("valid too small, return")
                pose = np.eye(4)
                pose[:3, 3] = head[:3]
                return pose
        self.H, self.W = depth.shape[:2]
        self.K = K
        self.ob_id = ob_id
        self.ob_mask = ob_mask
        poses, _ = self.refiner.predict(mesh=self.mesh, mesh_tensors=self.mesh_tensors, rgb=rgb, depth=depth, K=K,
                                        ob_in_cams=poses, normal_map=None, xyz_map=None, glctx=self.glctx,
                                        mesh_diameter=self.diameter, iteration=iteration, _frame_ready=True)
        scores, _ = self.scorer.predict(mesh=self.mesh, rgb=rgb, depth=depth, K=K, ob_in_cams=poses, normal_map=None,
                                        mesh_tensors=self.mesh_tensors, glctx=self.glctx, mesh_diameter=self.diameter,
                                        _frame_ready=True)
        ids = torch.as_tensor(scores).argsort(descending=True)
        scores = scores[ids]
        poses = poses[ids]
        best_pose = poses[0] @ self.get_tf_to_centered_mesh()
        prev = (self.pose_last, getattr(self, "best_id", None), getattr(self, "poses", None), getattr(self, "scores", None))
        self.pose_last = poses[0]
        self.best_id = ids[0]
        self.poses = poses
        self.scores = scores
        # the only host synchronisation of register(): (tx, ty, tz, n_valid) and the best pose in one read-back
        out = torch.cat([info, best_pose.reshape(-1)]).cpu().numpy()
        if out[3] < 4:
            # estimater.py:183-189: too few valid pixels -> identity rotation, guessed translation
            logging.info("valid too small, return")
            pose = np.eye(4)
            pose[:3, 3] = out[:3]
            self.pose_last, self.best_id, self.poses, self.scores = prev  # the reference returns before touching them
            return pose
        return out[4:].reshape(4, 4).copy()

    def track_one(self, rgb, depth, K, iteration, extra={}):
        if self.pose_last is None:
            logging.info("Please init pose by register first")
            raise RuntimeError
        e = self.engine
        _sync_mesh(e, self.mesh, self.mesh_tensors, self.diameter)
        if torch.is_tensor(rgb) or torch.is_tensor(depth):
            # device-resident frame: enqueue the stages one by one (estimater.py:255-264)
            e.set_frame(rgb, depth, K, filter_depth=True, zfar=float("inf"))
            pose, _ = self.refiner.predict(mesh=self.mesh, mesh_tensors=self.mesh_tensors, rgb=rgb, depth=depth, K=K,
                                           ob_in_cams=self.pose_last.reshape(1, 4, 4), normal_map=None, xyz_map=None,
                                           mesh_diameter=self.diameter, glctx=self.glctx, iteration=iteration, _frame_ready=True)
            self.pose_last = pose
            return (pose @ self.get_tf_to_centered_mesh()).data.cpu().numpy().reshape(4, 4)
        # host frame (what the drivers pass): the whole frame is one CUDA-graph launch (fp_track)
        pose_dev, pose_host = e.track(rgb, depth, K, self.pose_last.reshape(4, 4), iteration)
        self.pose_last = pose_dev.reshape(1, 4, 4)
        self.refiner.last_trans_update = self.refiner.last_rot_update = None
        out = pose_host.astype(np.float64)
        out[:3, 3] -= out[:3, :3] @ np.asarray(self.model_center, dtype=np.float64)  # pose @ T(-model_center), estimater.py:268
        return out.astype(np.float32)
