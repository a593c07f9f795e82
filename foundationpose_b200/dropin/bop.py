"""BOP-format scene readers behind the names of the reference's `datareader.py` (:17-53, :155-613), which its dataset
drivers use (`run_linemod.py:90-112`, `run_ycb_video.py:85-118`).

Design: ONE scene class, `BopBaseReader`, does all the work from three pieces of per-frame metadata parsed once
(`scene_camera.json`, `scene_gt.json`, the sorted colour files); a dataset is a row of `_DATASETS` (object ids, where
the models live relative to the scene, whether the test split has ground truth) turned into a subclass by
`_dataset_reader`.  Only LINEMOD (models found by walking up to `lm_models/`) and YCB-Video (key frames, model
names, geometric symmetries) add behaviour of their own.  Units: files are millimetres, everything returned is metres.
"""
import json
import os

import cv2
import numpy as np

from Utils import depth2xyzmap, euler_matrix, glob, imageio, logging, symmetry_tfs_from_info, trimesh

__all__ = ["BOP_LIST", "BOP_DIR", "get_bop_reader", "get_bop_video_dirs", "BopBaseReader", "LinemodOcclusionReader", "LinemodReader",
           "YcbVideoReader", "TlessReader", "HomebrewedReader", "ItoddReader", "IcbinReader", "TudlReader"]

BOP_LIST = ["lmo", "tless", "ycbv", "hb", "tudl", "icbin", "itodd"]
BOP_DIR = os.getenv("BOP_DIR")
_LM_NAMES = "ape benchvise bowl camera water_pour cat cup driller duck eggbox glue holepuncher iron lamp phone".split()


def _png(path, flags=-1):
    img = cv2.imread(path, flags)
    if img is None:
        raise FileNotFoundError(path)
    return img


def _rigid(annotation):
    """scene_gt.json entry -> 4x4 object-in-camera transform in metres."""
    T = np.eye(4)
    T[:3, :3] = np.asarray(annotation["cam_R_m2c"], dtype=float).reshape(3, 3)
    T[:3, 3] = np.asarray(annotation["cam_t_m2c"], dtype=float) * 1e-3
    return T


class BopBaseReader:
    ob_ids = ()

    def __init__(self, base_dir, zfar=np.inf, resize=1):
        self.base_dir, self.zfar, self.resize = base_dir, zfar, resize
        self.dataset_name = None
        self.scene_ob_ids_dict = None
        colour = glob.glob(os.path.join(base_dir, "rgb", "*"))
        self.color_files = sorted(colour if colour else glob.glob(os.path.join(base_dir, "gray", "*")))
        with open(os.path.join(base_dir, "scene_camera.json")) as fh:
            cameras = json.load(fh)
        self.K_table = {"%06d" % int(k): np.asarray(v["cam_K"], dtype=float).reshape(3, 3) for k, v in cameras.items()}
        self.bop_depth_scale = next(reversed(cameras.values()))["depth_scale"] if cameras else 1.0
        gt_path = os.path.join(base_dir, "scene_gt.json")
        self.scene_gt = None
        if os.path.isfile(gt_path):
            with open(gt_path) as fh:
                self.scene_gt = json.load(fh)
            if len(self.scene_gt) != len(self.color_files):
                raise AssertionError("scene_gt.json and the colour frames disagree in length")
        self.make_id_strs()

    # ------------------------------------------------------------------ frame bookkeeping
    def make_id_strs(self):
        self.id_strs = [os.path.splitext(os.path.basename(f))[0] for f in self.color_files]

    def _annotations(self, i_frame):
        return self.scene_gt[str(int(self.id_strs[i_frame]))]

    def _scaled(self, img, nearest=False):
        if self.resize == 1:
            return img
        return cv2.resize(img, None, fx=self.resize, fy=self.resize, interpolation=cv2.INTER_NEAREST if nearest else cv2.INTER_LINEAR)

    def get_video_id(self):
        return int(os.path.basename(os.path.normpath(self.base_dir)))

    get_video_dir = get_video_id

    def make_scene_ob_ids_dict(self):
        """BOP'19 target list of this scene: frame id -> object ids, one entry per instance."""
        table = {}
        with open(os.path.join(BOP_DIR, self.dataset_name, "test_targets_bop19.json")) as fh:
            for t in json.load(fh):
                if t["scene_id"] == self.get_video_id():
                    table.setdefault("%06d" % t["im_id"], []).extend([t["obj_id"]] * t["inst_count"])
        self.scene_ob_ids_dict = table

    def get_instance_ids_in_image(self, i_frame):
        if self.scene_gt is not None:
            ids = [a["obj_id"] for a in self._annotations(i_frame)]
        elif self.scene_ob_ids_dict is not None:
            ids = self.scene_ob_ids_dict[self.id_strs[i_frame]]
        else:  # neither ground truth nor targets: the annotation slots present as mask files
            pattern = os.path.join(os.path.dirname(self.color_files[0]).replace("rgb", "mask_visib"), self.id_strs[i_frame] + "_*.png")
            ids = [int(os.path.basename(f)[:-4].split("_")[1]) for f in sorted(glob.glob(pattern))]
        return np.asarray(ids)

    # ------------------------------------------------------------------ frame data
    def get_K(self, i_frame):
        K = self.K_table[self.id_strs[i_frame]]
        if self.resize != 1:  # a scaled COPY (scaling the table entry in place would compound over calls)
            K = K.copy()
            K[:2, :2] *= self.resize
        return K

    def get_color(self, i):
        img = imageio.imread(self.color_files[i])
        if img.ndim == 2:  # grey-scale datasets
            img = np.repeat(img[:, :, None], 3, axis=2)
        return self._scaled(img)

    def get_depth(self, i, filled=False):
        src = self.color_files[i]
        if filled:
            d, b = os.path.split(src.replace("rgb", "depth_filled"))
            depth = _png(os.path.join(d, "0" + b)) / 1e3
        else:
            depth = _png(src.replace("rgb", "depth").replace("gray", "depth")) * (1e-3 * self.bop_depth_scale)
        depth = self._scaled(depth, nearest=True)
        depth[np.logical_or(depth < 0.001, depth > self.zfar)] = 0
        return depth

    def get_xyz_map(self, i):
        return depth2xyzmap(self.get_depth(i), self.get_K(i))

    def get_mask(self, i_frame, ob_id, type="mask_visib"):
        """Boolean mask of the FIRST annotation of `ob_id` (`mask_visib`: visible part, `mask`: whole projection);
        None when the file does not exist."""
        if self.scene_gt is None:
            raise RuntimeError("get_mask needs scene_gt.json")
        anns = self._annotations(i_frame)
        slot = next((k for k, a in enumerate(anns) if a["obj_id"] == ob_id), len(anns))
        path = os.path.join(self.base_dir, type, "%06d_%06d.png" % (int(self.id_strs[i_frame]), slot))
        if not os.path.exists(path):
            logging.info(f"{path} not found")
            return None
        return self._scaled(_png(path), nearest=True) > 0

    # ------------------------------------------------------------------ models
    def get_gt_mesh_file(self, ob_id):
        raise RuntimeError("You should override this")

    def get_gt_mesh(self, ob_id):
        mesh = trimesh.load(self.get_gt_mesh_file(ob_id))
        mesh.vertices *= 1e-3
        return mesh

    def _models_info(self):
        with open(os.path.join(os.path.dirname(self.get_gt_mesh_file(self.ob_ids[0])), "models_info.json")) as fh:
            return json.load(fh)

    def get_model_diameter(self, ob_id):
        return self._models_info()[str(ob_id)]["diameter"] * 1e-3

    def load_symmetry_tfs(self):
        info = self._models_info()
        self.symmetry_info_table = {i: info[str(i)] for i in self.ob_ids}
        self.symmetry_tfs = {i: symmetry_tfs_from_info(e, rot_angle_discrete=5) for i, e in self.symmetry_info_table.items()}
        self.geometry_symmetry_info_table = json.loads(json.dumps(self.symmetry_info_table), object_hook=lambda d: {(int(k) if k.isdigit() else k): v for k, v in d.items()})

    # ------------------------------------------------------------------ ground truth
    def get_gt_poses(self, i_frame, ob_id):
        return np.asarray([_rigid(a) for a in self._annotations(i_frame) if a["obj_id"] == ob_id]).reshape(-1, 4, 4)

    def get_gt_pose(self, i_frame, ob_id, mask=None, use_my_correction=False):
        """Pose of `ob_id`: its first annotation, or — several instances and a detection `mask` given — the annotation
        whose visible mask has the largest IoU with it.  Identity when the object is not annotated in the frame."""
        candidates = [(k, a) for k, a in enumerate(self._annotations(i_frame)) if a["obj_id"] == ob_id]
        pose = np.eye(4)
        if candidates and mask is None:
            pose = _rigid(candidates[0][1])
        elif candidates:
            det = np.asarray(mask).astype(bool)
            scored = []
            for k, a in candidates:
                vis = _png(os.path.join(self.base_dir, "mask_visib", "%s_%06d.png" % (self.id_strs[i_frame], k))).astype(bool)
                scored.append((np.logical_and(vis, det).sum() / max(np.logical_or(vis, det).sum(), 1), -k, a))
            pose = _rigid(max(scored, key=lambda s: s[:2])[2])
        if use_my_correction and "ycb" in self.base_dir.lower() and "train_real" in self.color_files[i_frame]:
            if ob_id == 1 and self.get_video_id() in (12, 13, 14, 17, 24):  # the reference's fix-up of five training videos
                pose = pose @ self.symmetry_tfs[ob_id][1]
        return pose


# ---------------------------------------------------------------------------------------------------------------------
# datasets
# ---------------------------------------------------------------------------------------------------------------------
# name -> (class name, object ids, model directory relative to the scene directory, frame ids come from the target list)
_DATASETS = {
    "tless": ("TlessReader", range(1, 31), "../../../models_cad", False),
    "hb": ("HomebrewedReader", range(1, 34), "../../../hb_models/models", True),
    "itodd": ("ItoddReader", range(1, 29), "../../../itodd_models/models", True),
    "icbin": ("IcbinReader", range(1, 3), "../../../icbin_models/models", False),
    "tudl": ("TudlReader", range(1, 4), "../../../tudl_models/models", False),
}


def _dataset_reader(dataset):
    cls_name, ids, models, from_targets = _DATASETS[dataset]

    def init(self, base_dir, zfar=np.inf):
        BopBaseReader.__init__(self, base_dir, zfar=zfar)
        self.dataset_name = dataset
        self.ob_ids = list(ids)
        self.load_symmetry_tfs()
        if from_targets:
            self.make_scene_ob_ids_dict()

    def mesh_file(self, ob_id):
        return "%s/%s/obj_%06d.ply" % (self.base_dir, models, ob_id)

    return type(cls_name, (BopBaseReader,), {"__init__": init, "get_gt_mesh_file": mesh_file, "__doc__": f"BOP '{dataset}' scenes"})


_Tless, _Hb = _dataset_reader("tless"), _dataset_reader("hb")
ItoddReader, IcbinReader, TudlReader = (_dataset_reader(n) for n in ("itodd", "icbin", "tudl"))


class TlessReader(_Tless):
    """T-LESS: texture-less CAD models; they are given one uniform grey (the reference paints a pure-colour texture)."""

    def get_gt_mesh(self, ob_id):
        mesh = BopBaseReader.get_gt_mesh(self, ob_id)
        grey = np.full((len(mesh.vertices), 4), 200, dtype=np.uint8)
        grey[:, 3] = 255
        try:
            mesh.visual = trimesh.visual.ColorVisuals(vertex_colors=grey)
        except TypeError:  # the real trimesh wants the mesh as first argument
            mesh.visual = trimesh.visual.ColorVisuals(mesh, vertex_colors=grey)
        return mesh


class HomebrewedReader(_Hb):
    """HomebrewedDB: no public ground truth for the test split."""

    def get_gt_pose(self, i_frame, ob_id, use_my_correction=False):
        logging.info("WARN HomeBrewed doesn't have GT pose")
        return np.eye(4)


class LinemodOcclusionReader(BopBaseReader):
    _evaluated = (1, 5, 6, 8, 9, 10, 11, 12)

    def __init__(self, base_dir, zfar=np.inf):
        BopBaseReader.__init__(self, base_dir, zfar=zfar)
        self._finish("lmo")

    def _finish(self, dataset):
        self.dataset_name = dataset
        self.K = next(iter(self.K_table.values()))
        self.ob_id_to_names = {i + 1: n for i, n in enumerate(_LM_NAMES)}
        self.ob_ids = list(self._evaluated)
        self.load_symmetry_tfs()

    def get_gt_mesh_file(self, ob_id):
        return "%s/%s/models/obj_%06d.ply" % (BOP_DIR, self.dataset_name, ob_id)


class LinemodReader(LinemodOcclusionReader):
    """LINEMOD: one scene per object; models under the nearest ancestor directory that contains `lm_models/`.
    `split`: name of a text file of frame ids next to the scene's images (the reference reads it from a fixed path of
    its authors' machine)."""

    _evaluated = tuple(i for i in range(1, 16) if i not in (3, 7))  # bowl and cup are left out

    def __init__(self, base_dir, zfar=np.inf, split=None):
        BopBaseReader.__init__(self, base_dir, zfar=zfar)
        if split is not None:
            with open(os.path.join(base_dir, split + ".txt")) as fh:
                self.color_files = [os.path.join(base_dir, "rgb", "%06d.png" % int(tok)) for tok in fh.read().split()]
            self.make_id_strs()
        self._finish("lm")

    def get_gt_mesh_file(self, ob_id):
        here = os.path.abspath(self.base_dir)
        while not os.path.isdir(os.path.join(here, "lm_models")):
            up = os.path.dirname(here)
            if up == here:
                raise FileNotFoundError(f"no lm_models/ directory above {self.base_dir}")
            here = up
        return "%s/lm_models/models/obj_%06d.ply" % (here, ob_id)

    def get_reconstructed_mesh(self, ob_id, ref_view_dir):
        return trimesh.load(os.path.abspath("%s/ob_%07d/model/model.obj" % (ref_view_dir, ob_id)))


class YcbVideoReader(BopBaseReader):
    """YCB-Video: <root>/test/<scene>/, <root>/ycbv_models/models/obj_*.ply, <root>/models/<name>/ ($YCB_VIDEO_DIR) and
    <root>/keyframe.txt with `<scene:04d>/<frame:06d>` lines (absent in the BOP re-release, where every frame counts)."""

    # shapes without texture cues: z-axis cylinders (with / without the top-bottom flip) and boxes
    _CYLINDERS_FLIP, _CYLINDERS, _BOXES = (1, 4, 6, 18), (13,), (2, 3, 9, 21)

    def __init__(self, base_dir, zfar=np.inf):
        BopBaseReader.__init__(self, base_dir, zfar=zfar)
        self.dataset_name = "ycbv"
        self.K = next(iter(self.K_table.values()))
        self.ob_ids = list(range(1, 22))
        model_names = sorted(os.listdir(os.path.join(os.getenv("YCB_VIDEO_DIR"), "models")))
        self.ob_id_to_names = dict(zip(self.ob_ids, model_names))
        self.name_to_ob_id = {n: i for i, n in self.ob_id_to_names.items()}
        self.keyframe_lines = None
        if "BOP" not in self.base_dir:
            with open(os.path.join(self.base_dir, "..", "..", "keyframe.txt")) as fh:
                self.keyframe_lines = fh.read().splitlines()
        self.load_symmetry_tfs()
        z_axis = [{"axis": [0, 0, 1], "offset": [0, 0, 0]}]
        half_turns = [euler_matrix(rx, ry, rz) for rz in (0, np.pi) for rx in (0, np.pi) for ry in (0, np.pi)]
        for i in self._CYLINDERS_FLIP:
            self.geometry_symmetry_info_table[i] = {"symmetries_continuous": z_axis, "symmetries_discrete": euler_matrix(0, np.pi, 0).reshape(1, 4, 4).tolist()}
        for i in self._CYLINDERS:
            self.geometry_symmetry_info_table[i] = {"symmetries_continuous": z_axis}
        for i in self._BOXES:
            self.geometry_symmetry_info_table[i] = {"symmetries_discrete": np.asarray(half_turns).reshape(-1, 4, 4).tolist()}

    def get_gt_mesh_file(self, ob_id):
        path = "%s/../../ycbv_models/models/obj_%06d.ply" % (self.base_dir, ob_id)
        return os.path.abspath(path) if "BOP" in self.base_dir else path

    def get_gt_mesh(self, ob_id, get_posecnn_version=False):
        if get_posecnn_version:
            return trimesh.load(os.path.join(os.getenv("YCB_VIDEO_DIR"), "models", self.ob_id_to_names[ob_id], "textured_simple.obj"))
        ply = self.get_gt_mesh_file(ob_id)
        mesh = trimesh.load(ply, process=False)
        mesh.vertices *= 1e-3
        png = ply[:-4] + ".png"
        uv = getattr(mesh.visual, "uv", None)
        if uv is not None and os.path.exists(png):  # BOP ships the texture next to the model
            from PIL import Image

            tex = Image.open(png)
            mesh.visual = trimesh.visual.texture.TextureVisuals(uv=uv, image=tex, material=trimesh.visual.texture.SimpleMaterial(image=tex))
        return mesh

    def get_reconstructed_mesh(self, ob_id, ref_view_dir):
        return trimesh.load(os.path.abspath("%s/ob_%07d/model/model.obj" % (ref_view_dir, ob_id)))

    def get_transform_reconstructed_to_gt_model(self, ob_id):
        return np.eye(4)

    def is_keyframe(self, i):
        if self.keyframe_lines is None:
            return True
        return "%04d/%06d" % (self.get_video_id(), int(self.id_strs[i])) in self.keyframe_lines


# ---------------------------------------------------------------------------------------------------------------------
# lookup by path / dataset name
# ---------------------------------------------------------------------------------------------------------------------
_BY_PATH = (("ycbv", YcbVideoReader), ("YCB", YcbVideoReader), ("lmo", LinemodOcclusionReader), ("LINEMOD-O", LinemodOcclusionReader),
            ("tless", TlessReader), ("TLESS", TlessReader), ("hb", HomebrewedReader), ("tudl", TudlReader), ("icbin", IcbinReader),
            ("itodd", ItoddReader))
_TEST_SPLITS = {"ycbv": "ycbv/test", "lmo": "lmo/lmo_test_bop19/test", "tless": "tless/tless_test_primesense_bop19/test_primesense",
                "hb": "hb/hb_test_primesense_bop19/test_primesense", "tudl": "tudl/tudl_test_bop19/test", "icbin": "icbin/icbin_test_bop19/test",
                "itodd": "itodd/itodd_test_bop19/test"}


def get_bop_reader(video_dir, zfar=np.inf):
    """The reader whose dataset key occurs in the scene path (first match in the reference's order)."""
    for key, cls in _BY_PATH:
        if key in video_dir:
            return cls(video_dir, zfar=zfar)
    raise RuntimeError(f"no BOP reader for {video_dir}")


def get_bop_video_dirs(dataset):
    """Scene directories of the BOP'19 test split of `dataset` under $BOP_DIR."""
    if dataset not in _TEST_SPLITS:
        raise RuntimeError(f"unknown BOP dataset {dataset}")
    return sorted(glob.glob(os.path.join(BOP_DIR, _TEST_SPLITS[dataset], "*")))
