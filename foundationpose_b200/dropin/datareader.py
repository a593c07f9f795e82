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

`BopBaseReader`, `LinemodOcclusionReader`, `LinemodReader`, `YcbVideoReader` keep the reference's method names, argument
orders and directory conventions; the remaining BOP datasets (T-LESS, HB, ITODD, IC-BIN, TUD-L: datareader.py:533-613)
differ only in their object count and model directory and are generated from one table.
"""
from Utils import *  # noqa: F401,F403
import copy
import json
import os

import cv2
import numpy as np

from Utils import depth2xyzmap, euler_matrix, glob, imageio, logging, symmetry_tfs_from_info, trimesh

BOP_LIST = ["lmo", "tless", "ycbv", "hb", "tudl", "icbin", "itodd"]
BOP_DIR = os.getenv("BOP_DIR")


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


def get_bop_reader(video_dir, zfar=np.inf):
    """Reader class by substring of the scene path (datareader.py:17-33)."""
    if "ycbv" in video_dir or "YCB" in video_dir:
        return YcbVideoReader(video_dir, zfar=zfar)
    if "lmo" in video_dir or "LINEMOD-O" in video_dir:
        return LinemodOcclusionReader(video_dir, zfar=zfar)
    if "tless" in video_dir or "TLESS" in video_dir:
        return TlessReader(video_dir, zfar=zfar)
    for key, cls in (("hb", "HomebrewedReader"), ("tudl", "TudlReader"), ("icbin", "IcbinReader"), ("itodd", "ItoddReader")):
        if key in video_dir:
            return globals()[cls](video_dir, zfar=zfar)
    raise RuntimeError(f"no BOP reader for {video_dir}")


def get_bop_video_dirs(dataset):
    """Scene directories of a BOP'19 test split under $BOP_DIR (datareader.py:36-53)."""
    sub = {"ycbv": "ycbv/test", "lmo": "lmo/lmo_test_bop19/test", "tless": "tless/tless_test_primesense_bop19/test_primesense",
           "hb": "hb/hb_test_primesense_bop19/test_primesense", "tudl": "tudl/tudl_test_bop19/test", "icbin": "icbin/icbin_test_bop19/test",
           "itodd": "itodd/itodd_test_bop19/test"}
    if dataset not in sub:
        raise RuntimeError(f"unknown BOP dataset {dataset}")
    return sorted(glob.glob(f"{BOP_DIR}/{sub[dataset]}/*"))


class BopBaseReader:
    """One BOP scene directory (datareader.py:155-366).  Frame ids are the file stems of the colour images."""

    def __init__(self, base_dir, zfar=np.inf, resize=1):
        self.base_dir = base_dir
        self.resize = resize
        self.dataset_name = None
        self.zfar = zfar
        self.color_files = sorted(glob.glob(f"{base_dir}/rgb/*")) or sorted(glob.glob(f"{base_dir}/gray/*"))
        with open(f"{base_dir}/scene_camera.json") as fh:
            cams = json.load(fh)
        self.K_table = {}
        self.bop_depth_scale = 1.0
        for k, v in cams.items():
            self.K_table[f"{int(k):06d}"] = np.array(v["cam_K"], dtype=float).reshape(3, 3)
            self.bop_depth_scale = v["depth_scale"]
        self.scene_gt = None
        self.scene_ob_ids_dict = None
        gt_file = f"{base_dir}/scene_gt.json"
        if os.path.exists(gt_file):
            with open(gt_file) as fh:
                self.scene_gt = json.load(fh)
            assert len(self.scene_gt) == len(self.color_files), "scene_gt.json does not cover every frame"
        self.make_id_strs()

    # ---- ids
    def make_id_strs(self):
        self.id_strs = [os.path.basename(f).split(".")[0] for f in self.color_files]

    def get_video_id(self):
        return int(self.base_dir.rstrip("/").split("/")[-1])

    def get_video_dir(self):
        return self.get_video_id()

    def make_scene_ob_ids_dict(self):
        """Targets of the BOP'19 challenge file: frame id -> object ids (repeated per instance)."""
        self.scene_ob_ids_dict = {}
        with open(f"{BOP_DIR}/{self.dataset_name}/test_targets_bop19.json") as fh:
            for d in json.load(fh):
                if d["scene_id"] == self.get_video_id():
                    self.scene_ob_ids_dict.setdefault(f"{d['im_id']:06d}", []).extend([d["obj_id"]] * d["inst_count"])

    def get_instance_ids_in_image(self, i_frame):
        if self.scene_gt is not None:
            return np.asarray([a["obj_id"] for a in self.scene_gt[str(int(self.id_strs[i_frame]))]])
        if self.scene_ob_ids_dict is not None:
            return np.array(self.scene_ob_ids_dict[self.id_strs[i_frame]])
        mask_dir = os.path.dirname(self.color_files[0]).replace("rgb", "mask_visib")
        files = sorted(glob.glob(f"{mask_dir}/{self.id_strs[i_frame]}_*.png"))
        return np.asarray([int(os.path.basename(f).split(".")[0].split("_")[1]) for f in files])

    # ---- frame data
    def get_K(self, i_frame):
        K = self.K_table[self.id_strs[i_frame]]
        if self.resize != 1:  # a scaled COPY (scaling the table entry in place would compound over calls)
            K = K.copy()
            K[:2, :2] *= self.resize
        return K

    def get_color(self, i):
        color = imageio.imread(self.color_files[i])
        if color.ndim == 2:
            color = np.tile(color[..., None], (1, 1, 3))
        if self.resize != 1:
            color = cv2.resize(color, fx=self.resize, fy=self.resize, dsize=None)
        return color

    def get_depth(self, i, filled=False):
        if filled:
            f = self.color_files[i].replace("rgb", "depth_filled")
            depth = cv2.imread(f"{os.path.dirname(f)}/0{os.path.basename(f)}", -1) / 1e3
        else:
            f = self.color_files[i].replace("rgb", "depth").replace("gray", "depth")
            depth = cv2.imread(f, -1) * 1e-3 * self.bop_depth_scale
        if self.resize != 1:
            depth = cv2.resize(depth, fx=self.resize, fy=self.resize, dsize=None, interpolation=cv2.INTER_NEAREST)
        depth[(depth < 0.001) | (depth > self.zfar)] = 0
        return depth

    def get_xyz_map(self, i):
        return depth2xyzmap(self.get_depth(i), self.get_K(i))

    def get_mask(self, i_frame, ob_id, type="mask_visib"):
        """`mask_visib` (visible part) or `mask` (whole projected model) of the FIRST annotation of `ob_id`."""
        if self.scene_gt is None:
            raise RuntimeError("get_mask needs scene_gt.json")
        name = int(self.id_strs[i_frame])
        pos = 0
        for a in self.scene_gt[str(name)]:
            if a["obj_id"] == ob_id:
                break
            pos += 1
        f = f"{self.base_dir}/{type}/{name:06d}_{pos:06d}.png"
        if not os.path.exists(f):
            logging.info(f"{f} not found")
            return None
        mask = cv2.imread(f, -1)
        if self.resize != 1:
            mask = cv2.resize(mask, fx=self.resize, fy=self.resize, dsize=None, interpolation=cv2.INTER_NEAREST)
        return mask > 0

    # ---- models
    def get_gt_mesh_file(self, ob_id):
        raise RuntimeError("You should override this")

    def get_gt_mesh(self, ob_id):
        mesh = trimesh.load(self.get_gt_mesh_file(ob_id))
        mesh.vertices *= 1e-3
        return mesh

    def _models_info(self):
        with open(f"{os.path.dirname(self.get_gt_mesh_file(self.ob_ids[0]))}/models_info.json") as fh:
            return json.load(fh)

    def get_model_diameter(self, ob_id):
        return self._models_info()[str(ob_id)]["diameter"] / 1e3

    def load_symmetry_tfs(self):
        info = self._models_info()
        self.symmetry_tfs = {}
        self.symmetry_info_table = {}
        for ob_id in self.ob_ids:
            self.symmetry_info_table[ob_id] = info[str(ob_id)]
            self.symmetry_tfs[ob_id] = symmetry_tfs_from_info(info[str(ob_id)], rot_angle_discrete=5)
        self.geometry_symmetry_info_table = copy.deepcopy(self.symmetry_info_table)

    # ---- ground truth
    @staticmethod
    def _pose_of(a):
        T = np.eye(4)
        T[:3, :3] = np.array(a["cam_R_m2c"], dtype=float).reshape(3, 3)
        T[:3, 3] = np.array(a["cam_t_m2c"], dtype=float) / 1e3
        return T

    def get_gt_poses(self, i_frame, ob_id):
        anns = self.scene_gt[str(int(self.id_strs[i_frame]))]
        return np.asarray([self._pose_of(a) for a in anns if a["obj_id"] == ob_id]).reshape(-1, 4, 4)

    def get_gt_pose(self, i_frame, ob_id, mask=None, use_my_correction=False):
        """First annotation of `ob_id`; with several instances and a `mask`, the one whose visible mask overlaps it most."""
        best, best_iou = np.eye(4), -np.inf
        for k, a in enumerate(self.scene_gt[str(int(self.id_strs[i_frame]))]):
            if a["obj_id"] != ob_id:
                continue
            if mask is None:
                best = self._pose_of(a)
                break
            gt_mask = cv2.imread(f"{self.base_dir}/mask_visib/{self.id_strs[i_frame]}_{k:06d}.png", -1).astype(bool)
            union = (gt_mask | mask.astype(bool)).sum()
            iou = float((gt_mask & mask.astype(bool)).sum()) / max(union, 1)
            if iou > best_iou:
                best_iou, best = iou, self._pose_of(a)
        if use_my_correction and "ycb" in self.base_dir.lower() and "train_real" in self.color_files[i_frame]:
            if ob_id == 1 and self.get_video_id() in (12, 13, 14, 17, 24):
                best = best @ self.symmetry_tfs[ob_id][1]
        return best


class LinemodOcclusionReader(BopBaseReader):
    def __init__(self, base_dir, zfar=np.inf):
        super().__init__(base_dir, zfar=zfar)
        self.dataset_name = "lmo"
        self.K = list(self.K_table.values())[0]
        self.ob_ids = [1, 5, 6, 8, 9, 10, 11, 12]
        self.ob_id_to_names = dict(zip(range(1, 16), ["ape", "benchvise", "bowl", "camera", "water_pour", "cat", "cup", "driller", "duck",
                                                     "eggbox", "glue", "holepuncher", "iron", "lamp", "phone"]))
        self.load_symmetry_tfs()

    def get_gt_mesh_file(self, ob_id):
        return f"{BOP_DIR}/{self.dataset_name}/models/obj_{ob_id:06d}.ply"


class LinemodReader(LinemodOcclusionReader):
    """LINEMOD: one scene directory per object; the models live in the nearest ancestor that has `lm_models/`."""

    def __init__(self, base_dir, zfar=np.inf, split=None):
        BopBaseReader.__init__(self, base_dir, zfar=zfar)
        self.dataset_name = "lm"
        self.K = list(self.K_table.values())[0]
        self.ob_id_to_names = dict(zip(range(1, 16), ["ape", "benchvise", "bowl", "camera", "water_pour", "cat", "cup", "driller", "duck",
                                                     "eggbox", "glue", "holepuncher", "iron", "lamp", "phone"]))
        if split is not None:  # file with one frame id per line, next to the scene directory
            with open(f"{self.base_dir}/{split}.txt") as fh:
                ids = [int(x) for x in fh.read().split()]
            self.color_files = [f"{self.base_dir}/rgb/{i:06d}.png" for i in ids]
            self.make_id_strs()
        self.ob_ids = [i for i in range(1, 16) if i not in (3, 7)]  # bowl and cup are not evaluated
        self.load_symmetry_tfs()

    def get_gt_mesh_file(self, ob_id):
        root = os.path.abspath(self.base_dir)
        while not os.path.exists(f"{root}/lm_models"):
            parent = os.path.dirname(root)
            if parent == root:
                raise FileNotFoundError(f"no lm_models/ directory above {self.base_dir}")
            root = parent
        return f"{root}/lm_models/models/obj_{ob_id:06d}.ply"

    def get_reconstructed_mesh(self, ob_id, ref_view_dir):
        return trimesh.load(os.path.abspath(f"{ref_view_dir}/ob_{ob_id:07d}/model/model.obj"))


class YcbVideoReader(BopBaseReader):
    """YCB-Video in BOP layout: <root>/test/<scene>/..., <root>/ycbv_models/models/obj_*.ply, <root>/models/<name>/,
    <root>/keyframe.txt (`<scene:04d>/<frame:06d>` per line; absent for the BOP re-release)."""

    def __init__(self, base_dir, zfar=np.inf):
        super().__init__(base_dir, zfar=zfar)
        self.dataset_name = "ycbv"
        self.K = list(self.K_table.values())[0]
        self.ob_ids = list(range(1, 22))
        names = sorted(os.listdir(f"{os.getenv('YCB_VIDEO_DIR')}/models/"))
        self.ob_id_to_names = {ob_id: names[i] for i, ob_id in enumerate(self.ob_ids)}
        self.name_to_ob_id = {v: k for k, v in self.ob_id_to_names.items()}
        self.keyframe_lines = None
        if "BOP" not in self.base_dir:
            with open(f"{self.base_dir}/../../keyframe.txt") as fh:
                self.keyframe_lines = fh.read().splitlines()
        self.load_symmetry_tfs()
        flips = [euler_matrix(rx, ry, rz) for rz in (0, np.pi) for rx in (0, np.pi) for ry in (0, np.pi)]
        for ob_id in self.ob_ids:  # geometric (texture-less) symmetries: cans, the bowl, boxes
            if ob_id in (1, 4, 6, 18):
                self.geometry_symmetry_info_table[ob_id] = {"symmetries_continuous": [{"axis": [0, 0, 1], "offset": [0, 0, 0]}],
                                                            "symmetries_discrete": euler_matrix(0, np.pi, 0).reshape(1, 4, 4).tolist()}
            elif ob_id == 13:
                self.geometry_symmetry_info_table[ob_id] = {"symmetries_continuous": [{"axis": [0, 0, 1], "offset": [0, 0, 0]}]}
            elif ob_id in (2, 3, 9, 21):
                self.geometry_symmetry_info_table[ob_id] = {"symmetries_discrete": np.asarray(flips).reshape(-1, 4, 4).tolist()}

    def get_gt_mesh_file(self, ob_id):
        f = f"{self.base_dir}/../../ycbv_models/models/obj_{ob_id:06d}.ply"
        return os.path.abspath(f) if "BOP" in self.base_dir else f

    def get_gt_mesh(self, ob_id, get_posecnn_version=False):
        if get_posecnn_version:
            return trimesh.load(f"{os.getenv('YCB_VIDEO_DIR')}/models/{self.ob_id_to_names[ob_id]}/textured_simple.obj")
        mesh_file = self.get_gt_mesh_file(ob_id)
        mesh = trimesh.load(mesh_file, process=False)
        mesh.vertices *= 1e-3
        tex_file = mesh_file.replace(".ply", ".png")
        if os.path.exists(tex_file) and getattr(mesh.visual, "uv", None) is not None:
            from PIL import Image

            im = Image.open(tex_file)
            mesh.visual = trimesh.visual.texture.TextureVisuals(uv=mesh.visual.uv, image=im,
                                                                material=trimesh.visual.texture.SimpleMaterial(image=im))
        return mesh

    def get_reconstructed_mesh(self, ob_id, ref_view_dir):
        return trimesh.load(os.path.abspath(f"{ref_view_dir}/ob_{ob_id:07d}/model/model.obj"))

    def get_transform_reconstructed_to_gt_model(self, ob_id):
        return np.eye(4)

    def is_keyframe(self, i):
        if self.keyframe_lines is None:
            return True
        frame_id = int(os.path.basename(self.color_files[i]).split(".")[0])
        return f"{self.get_video_id():04d}/{frame_id:06d}" in self.keyframe_lines


def _simple_bop_reader(name, dataset_name, n_objects, models_subdir, targets=False, doc=""):
    """The five BOP readers that differ only by object count and model directory (datareader.py:533-613)."""

    class _Reader(BopBaseReader):
        def __init__(self, base_dir, zfar=np.inf):
            super().__init__(base_dir, zfar=zfar)
            self.dataset_name = dataset_name
            self.ob_ids = list(range(1, n_objects + 1))
            self.load_symmetry_tfs()
            if targets:  # no scene_gt.json in the test split: object ids per frame come from the challenge's target list
                self.make_scene_ob_ids_dict()

        def get_gt_mesh_file(self, ob_id):
            return f"{self.base_dir}/../../../{models_subdir}/obj_{ob_id:06d}.ply"

    _Reader.__name__ = _Reader.__qualname__ = name
    _Reader.__doc__ = doc
    return _Reader


_TlessBase = _simple_bop_reader("TlessReader", "tless", 30, "models_cad")


class TlessReader(_TlessBase):
    """T-LESS: texture-less CAD models, rendered with a uniform grey (datareader.py:547-551)."""

    def get_gt_mesh(self, ob_id):
        mesh = trimesh.load(self.get_gt_mesh_file(ob_id))
        mesh.vertices *= 1e-3
        grey = np.tile(np.array([[200, 200, 200, 255]], dtype=np.uint8), (len(mesh.vertices), 1))
        mesh.visual = trimesh.visual.ColorVisuals(vertex_colors=grey) if hasattr(trimesh.visual, "ColorVisuals") else mesh.visual
        if getattr(mesh.visual, "vertex_colors", None) is None or len(mesh.visual.vertex_colors) != len(mesh.vertices):
            mesh.visual.vertex_colors = grey
        return mesh


_HbBase = _simple_bop_reader("HomebrewedReader", "hb", 33, "hb_models/models", targets=True)


class HomebrewedReader(_HbBase):
    """HomebrewedDB: the test split has no public ground truth (datareader.py:568-570)."""

    def get_gt_pose(self, i_frame, ob_id, use_my_correction=False):
        logging.info("WARN HomeBrewed doesn't have GT pose")
        return np.eye(4)


ItoddReader = _simple_bop_reader("ItoddReader", "itodd", 28, "itodd_models/models", targets=True, doc="ITODD (grey-scale frames under gray/)")
IcbinReader = _simple_bop_reader("IcbinReader", "icbin", 2, "icbin_models/models", doc="IC-BIN")
TudlReader = _simple_bop_reader("TudlReader", "tudl", 3, "tudl_models/models", doc="TUD-L")

