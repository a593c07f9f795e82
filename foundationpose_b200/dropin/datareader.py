"""`from datareader import *` (run_demo.py:11): the reader of the reference's demo / YCBInEOAT scene layout
(datareader.py:57-152):

    <video_dir>/cam_K.txt            3x3 intrinsics
    <video_dir>/rgb/*.png            colour frames (the sorted file stems are the frame ids)
    <video_dir>/depth/*.png          uint16 depth in millimetres
    <video_dir>/masks/*.png          object mask of (at least) the first frame
    <video_dir>/annotated_poses/*    optional ground-truth poses, one 4x4 text file per frame

The BOP readers of the reference (datareader.py:155-660) are dataset plumbing outside the hot path and are not provided.
"""
from Utils import *  # noqa: F401,F403
import os

import cv2
import numpy as np

from Utils import depth2xyzmap, glob, imageio, logging, trimesh


class YcbineoatReader:
    def __init__(self, video_dir, downscale=1, shorter_side=None, zfar=np.inf):
        self.video_dir = video_dir
        self.downscale = downscale
        self.zfar = zfar
        self.color_files = sorted(glob.glob(f"{self.video_dir}/rgb/*.png"))
        self.K = np.loadtxt(f"{video_dir}/cam_K.txt").reshape(3, 3)
        self.id_strs = [os.path.basename(f).replace(".png", "") for f in self.color_files]
        self.H, self.W = cv2.imread(self.color_files[0]).shape[:2]
        if shorter_side is not None:
            self.downscale = shorter_side / min(self.H, self.W)
        self.H = int(self.H * self.downscale)
        self.W = int(self.W * self.downscale)
        self.K[:2] *= self.downscale
        self.gt_pose_files = sorted(glob.glob(f"{self.video_dir}/annotated_poses/*"))
        self.videoname_to_object = {
            "bleach0": "021_bleach_cleanser", "bleach_hard_00_03_chaitanya": "021_bleach_cleanser",
            "cracker_box_reorient": "003_cracker_box", "cracker_box_yalehand0": "003_cracker_box",
            "mustard0": "006_mustard_bottle", "mustard_easy_00_02": "006_mustard_bottle",
            "sugar_box1": "004_sugar_box", "sugar_box_yalehand0": "004_sugar_box",
            "tomato_soup_can_yalehand0": "005_tomato_soup_can",
        }

    def get_video_name(self):
        return self.video_dir.split("/")[-1]

    def __len__(self):
        return len(self.color_files)

    def _resize(self, img):
        return cv2.resize(img, (self.W, self.H), interpolation=cv2.INTER_NEAREST)

    def get_gt_pose(self, i):
        try:
            return np.loadtxt(self.gt_pose_files[i]).reshape(4, 4)
        except Exception:
            logging.info("GT pose not found, return None")
            return None

    def get_color(self, i):
        return self._resize(imageio.imread(self.color_files[i])[..., :3])

    def get_mask(self, i):
        mask = cv2.imread(self.color_files[i].replace("rgb", "masks"), -1)
        if mask.ndim == 3:  # first non-empty channel
            for c in range(3):
                if mask[..., c].sum() > 0:
                    mask = mask[..., c]
                    break
        return self._resize(mask).astype(bool).astype(np.uint8)

    def get_depth(self, i):
        depth = self._resize(cv2.imread(self.color_files[i].replace("rgb", "depth"), -1) / 1e3)
        depth[(depth < 0.001) | (depth >= self.zfar)] = 0
        return depth

    def get_xyz_map(self, i):
        return depth2xyzmap(self.get_depth(i), self.K)

    def get_occ_mask(self, i):
        occ = np.zeros((self.H, self.W), dtype=bool)
        for sub in ("masks_hand", "masks_hand_right"):
            f = self.color_files[i].replace("rgb", sub)
            if os.path.exists(f):
                occ |= self._resize((cv2.imread(f, -1) > 0).astype(np.uint8)).astype(bool)
        return occ.astype(np.uint8)

    def get_gt_mesh(self):
        ob_name = self.videoname_to_object[self.get_video_name()]
        return trimesh.load(f"{os.getenv('YCB_VIDEO_DIR')}/models/{ob_name}/textured_simple.obj")
