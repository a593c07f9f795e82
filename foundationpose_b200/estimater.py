"""Host-side mirror of the reference's estimator and predictor surfaces, running on libfpose.so.

Same names, argument orders and error behaviour as
  estimater.py:18-268                              FoundationPose {register, track_one, reset_object, ...}
  learning/training/predict_pose_refine.py:92-239  PoseRefinePredictor {predict, last_trans_update, ...}
  learning/training/predict_score.py:117-226       ScorePredictor {predict}
so a driver written against the reference (run_demo.py:38-63) only changes its import.

Differences that are deliberate: no nvdiffrast context is needed (`glctx` is accepted
and ignored), visualisation (`get_vis`, debug >= 2 dumps) is not produced, and `predict()` re-derives
the xyz map on the device from `depth` instead of uploading the caller's copy.
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
    c = _Cfg(weights.DEFAULT_CFG)
    if cfg:
        c.update(cfg)
    if state_dict is None:
        path = weights.find_reference_weights(run_name)
        if path is not None:
            state_dict = weights.load_checkpoint(path)
            logging.info(f"Using pretrained model from {path}")
        else:
            logging.info(f"weights/{run_name}/model_best.pth not found: using the seeded synthetic {kind} weights")
            state_dict = weights.random_state_dict(kind, seed=0)
    if not c["normalize_xyz"] or c["rot_rep"] != "axis_angle" or c["trans_rep"] != "tracknet" or c["use_normal"]:
        raise NotImplementedError("engine supports the released configs: normalize_xyz, tracknet, axis_angle, no normals")
    return c, state_dict


class PoseRefinePredictor:
    def __init__(self, engine=None, state_dict=None, cfg=None):
        self.amp = True
        self.run_name = "2023-10-28-18-33-37"
        self.engine = engine or get_engine()
        self.cfg, sd = _load_cfg_and_weights(self.run_name, "refine", state_dict, cfg)
        self.engine.load_network("refine", sd)
        self.engine.set_config(self.cfg["crop_ratio"], self.cfg["rot_normalizer"])
        self.model = _ModelHandle()
        self.dataset = None
        self.last_trans_update = None
        self.last_rot_update = None

    @torch.inference_mode()
    def predict(self, rgb, depth, K, ob_in_cams, xyz_map=None, normal_map=None, get_vis=False, mesh=None, mesh_tensors=None,
                glctx=None, mesh_diameter=None, iteration=5, _frame_ready=False):
        """@rgb: (H,W,3) uint8; @ob_in_cams: (N,4,4).  Returns ((N,4,4) cuda tensor, None)."""
        e = self.engine
        if not _frame_ready:
            e.set_frame(rgb, depth, K, filter_depth=False, zfar=self.cfg["zfar"] if np.isfinite(self.cfg["zfar"]) else float("inf"))
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
        self.model = _ModelHandle()
        self.dataset = None

    @torch.inference_mode()
    def predict(self, rgb, depth, K, ob_in_cams, normal_map=None, get_vis=False, mesh=None, mesh_tensors=None, glctx=None,
                mesh_diameter=None, _frame_ready=False):
        """Returns ((N,) cuda tensor of scores = logits + 100, None)."""
        e = self.engine
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
        if symmetry_tfs is None:
            self.symmetry_tfs = np.eye(4, dtype=np.float32)[None]
        else:
            self.symmetry_tfs = np.asarray(symmetry_tfs, dtype=np.float32)

    def get_tf_to_centered_mesh(self):
        tf = torch.eye(4, dtype=torch.float32, device="cuda")
        tf[:3, 3] = -torch.as_tensor(self.model_center, device="cuda", dtype=torch.float32)
        return tf

    def to_device(self, s="cuda:0"):
        return self  # weights, mesh and frame live in the fp_ctx of the current device

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
        e.set_frame(rgb, depth, K, filter_depth=True, zfar=float("inf"))
        pose, _ = self.refiner.predict(mesh=self.mesh, mesh_tensors=self.mesh_tensors, rgb=rgb, depth=depth, K=K,
                                       ob_in_cams=self.pose_last.reshape(1, 4, 4), normal_map=None, xyz_map=None,
                                       mesh_diameter=self.diameter, glctx=self.glctx, iteration=iteration, _frame_ready=True)
        self.pose_last = pose
        return (pose @ self.get_tf_to_centered_mesh()).data.cpu().numpy().reshape(4, 4)
