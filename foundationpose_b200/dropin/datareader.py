"""`from datareader import *` (run_demo.py:11): the reader of the reference's demo / YCBInEOAT scene layout
(datareader.py:57-152):

    <video_dir>/cam_K.txt            3x3 intrinsics
    <video_dir>/rgb/*.png            colour frames (the sorted file stems are the frame ids)
    <video_dir>/depth/*.png          uint16 depth in millimetres
    <video_dir>/masks/*.png          object mask of (at least) the first frame
    <video_dir>/annotated_poses/*    optional ground-truth poses, one 4x4 text file per frame

and the BOP-format readers its dataset drivers use (run_linemod.py:90-112, run_ycb_video.py:85-118; datareader.py:155-531):

    <scene>/rgb|gray/<frame:06d>.png          colour
    <scene>/depth/<frame:06d>.png             uint16, metres = value * 1e-3 * scene_camera[frame].depth_scale
    <scene>/mask_visib/<frame:06d>_<k:06d>.png   visible mask of the k-th annotation of the frame
    <scene>/scene_camera.json, scene_gt.json  {frame: {cam_K, depth_scale}}, {frame: [{obj_id, cam_R_m2c, cam_t_m2c [mm]}]}
    <models>/obj_<id:06d>.ply (+ models_info.json: diameter [mm], symmetries)   model units: millimetres

The BOP readers live in `bop.py` (one scene-directory class driven by a per-dataset table) and are re-exported here
under the reference's names: `BopBaseReader`, `LinemodOcclusionReader`, `LinemodReader`, `YcbVideoReader`, `TlessReader`,
`HomebrewedReader`, `ItoddReader`, `IcbinReader`, `TudlReader`, `get_bop_reader`, `get_bop_video_dirs`, `BOP_LIST`, `BOP_DIR`.
"""
from Utils import *  # noqa: F401,F403
import os

import cv2
import numpy as np

from Utils import depth2xyzmap, glob, imageio, logging, trimesh


_YCBINEOAT_OBJECTS = {  # video name -> YCB model directory (datareader.py:78-88)
    "021_bleach_cleanser": ("bleach0", "bleach_hard_00_03_chaitanya"),
    "003_cracker_box": ("cracker_box_reorient", "cracker_box_yalehand0"),
    "006_mustard_bottle": ("mustard0", "mustard_easy_00_02"),
    "004_sugar_box": ("sugar_box1", "sugar_box_yalehand0"),
    "005_tomato_soup_can": ("tomato_soup_can_yalehand0",),
}


class YcbineoatReader:
    """Demo / YCBInEOAT scene directory.  Images are resized (nearest neighbour) by `downscale`, or so that the
    shorter side becomes `shorter_side`; K is scaled with them."""

    def __init__(self, video_dir, downscale=1, shorter_side=None, zfar=np.inf):
        self.video_dir, self.zfar = video_dir, zfar
        self.color_files = sorted(glob.glob(os.path.join(video_dir, "rgb", "*.png")))
        self.id_strs = [os.path.basename(f)[:-4] for f in self.color_files]
        self.gt_pose_files = sorted(glob.glob(os.path.join(video_dir, "annotated_poses", "*")))
        h0, w0 = cv2.imread(self.color_files[0]).shape[:2]
        self.downscale = downscale if shorter_side is None else shorter_side / min(h0, w0)
        self.H, self.W = int(h0 * self.downscale), int(w0 * self.downscale)
        self.K = np.loadtxt(os.path.join(video_dir, "cam_K.txt")).reshape(3, 3)
        self.K[:2] *= self.downscale
        self.videoname_to_object = {video: model for model, videos in _YCBINEOAT_OBJECTS.items() for video in videos}

    def __len__(self):
        return len(self.color_files)

    def get_video_name(self):
        return self.video_dir.split("/")[-1]

    def _fit(self, img):
        return cv2.resize(img, (self.W, self.H), interpolation=cv2.INTER_NEAREST)

    def _sibling(self, i, folder):
        """The file of frame i in another folder of the scene (same file name)."""
        return self.color_files[i].replace("rgb", folder)

    def get_color(self, i):
        return self._fit(imageio.imread(self.color_files[i])[..., :3])

    def get_depth(self, i):
        metres = self._fit(cv2.imread(self._sibling(i, "depth"), -1) / 1e3)
        metres[np.logical_or(metres < 0.001, metres >= self.zfar)] = 0
        return metres

    def get_xyz_map(self, i):
        return depth2xyzmap(self.get_depth(i), self.K)

    def get_mask(self, i):
        raw = cv2.imread(self._sibling(i, "masks"), -1)
        if raw.ndim == 3:  # colour mask: the first channel that is not empty
            filled = [c for c in range(3) if raw[..., c].any()]
            raw = raw[..., filled[0]] if filled else raw[..., 0]
        return self._fit(raw).astype(bool).astype(np.uint8)

    def get_occ_mask(self, i):
        """Union of the hand masks (left and right), when the scene has them."""
        hidden = np.zeros((self.H, self.W), dtype=bool)
        for folder in ("masks_hand", "masks_hand_right"):
            path = self._sibling(i, folder)
            if os.path.exists(path):
                hidden |= self._fit((cv2.imread(path, -1) > 0).astype(np.uint8)) > 0
        return hidden.astype(np.uint8)

    def get_gt_pose(self, i):
        try:
            return np.loadtxt(self.gt_pose_files[i]).reshape(4, 4)
        except Exception:
            logging.info("GT pose not found, return None")
            return None

    def get_gt_mesh(self):
        model = self.videoname_to_object[self.get_video_name()]
        return trimesh.load(os.path.join(os.getenv("YCB_VIDEO_DIR"), "models", model, "textured_simple.obj"))


from bop import *  # noqa: E402,F401,F403  (the BOP-format readers)
from bop import BOP_DIR, BOP_LIST  # noqa: E402,F401
