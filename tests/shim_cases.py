"""Cases for the helper functions of the drop-in `Utils` module (not a test module; imported by
tools/make_golden_shim.py and tests/test_dropin_golden_cpu.py).

`collect(ns)` calls the helpers by NAME out of a namespace — either the reference's own function bodies (extracted
unmodified from /root/reference/Utils.py with `ast`, tools/make_golden_shim.py) or the drop-in module — on seeded inputs
and returns {name: array}.  Functions covered (reference lines): to_homo :511, transform_pts :529, project_3d_to_2d :667,
draw_xyz_axis :675, draw_posed_3d_box :713, symmetry_tfs_from_info :806, make_yaml_dumpable :996 (+ NestDict :60),
depth2xyzmap :399, depth2xyzmap_batch :420, compute_mesh_diameter :559 (model_pts branch), set_seed :222.
"""
import random

import numpy as np
import torch
import yaml

K = np.array([[615.0, 0.0, 320.0], [0.0, 615.0, 240.0], [0.0, 0.0, 1.0]])


def _pose(seed, t):
    from scipy.spatial.transform import Rotation

    T = np.eye(4)
    T[:3, :3] = Rotation.random(random_state=seed).as_matrix()
    T[:3, 3] = t
    return T


def _image(seed):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 255, size=(30, 40, 3), dtype=np.uint8)
    return np.ascontiguousarray(np.kron(img, np.ones((16, 16, 1), dtype=np.uint8)))  # 480 x 640, blocky


def collect(ns):
    out = {}
    g = lambda name: ns[name] if isinstance(ns, dict) else getattr(ns, name)
    rng = np.random.default_rng(0)

    # ---- points
    pts = rng.normal(size=(7, 3))
    out["to_homo.3"] = g("to_homo")(pts)
    out["to_homo.2"] = g("to_homo")(pts[:, :2])
    T = _pose(1, [0.1, -0.2, 0.7])
    out["transform_pts.np"] = g("transform_pts")(pts, T)
    Tb = np.stack([_pose(s, [0.0, 0.1 * s, 0.5]) for s in range(3)])
    out["transform_pts.np_batch"] = g("transform_pts")(np.stack([pts, pts + 1, pts - 1]), Tb)
    out["transform_pts.torch"] = g("transform_pts")(torch.as_tensor(pts, dtype=torch.float32), torch.as_tensor(T, dtype=torch.float32)).numpy()
    # the render window of the predictors: 2-D corner points through inverse 3x3 crop transforms (predict_pose_refine.py:44-45)
    tf = torch.tensor([[[2.0, 0.0, -100.0], [0.0, 2.0, -60.0], [0.0, 0.0, 1.0]], [[1.6, 0.0, -320.0], [0.0, 1.6, -160.0], [0.0, 0.0, 1.0]]])
    corners = torch.tensor([[0.0, 0.0], [159.0, 159.0]]).reshape(1, 2, 2).expand(2, -1, -1)
    out["transform_pts.bbox2d_ori"] = g("transform_pts")(corners, tf.inverse()[:, None]).reshape(-1, 4).numpy()

    # ---- pose visualisation (run_demo.py:71-74)
    pose = _pose(2, [0.03, -0.02, 0.6])
    for i, p in enumerate(np.array([[0, 0, 0, 1], [0.1, 0, 0, 1], [0, -0.05, 0.02, 1.0]])):
        out[f"project_3d_to_2d.{i}"] = g("project_3d_to_2d")(p, K, pose)
    bbox = np.array([[-0.05, -0.03, -0.09], [0.05, 0.03, 0.09]])
    img = _image(3)
    out["draw_posed_3d_box"] = g("draw_posed_3d_box")(K, img=img.copy(), ob_in_cam=pose, bbox=bbox)
    out["draw_posed_3d_box.red_thick"] = g("draw_posed_3d_box")(K, img=img.copy(), ob_in_cam=_pose(5, [-0.1, 0.05, 0.5]), bbox=bbox, line_color=(255, 0, 0), linewidth=4)
    out["draw_xyz_axis"] = g("draw_xyz_axis")(img.copy(), ob_in_cam=pose, scale=0.1, K=K, thickness=3, transparency=0, is_input_rgb=True)
    out["draw_xyz_axis.bgr_transparent"] = g("draw_xyz_axis")(img.copy(), ob_in_cam=pose, scale=0.07, K=K, thickness=2, transparency=0.3, is_input_rgb=False)
    # the combination run_demo.py draws: box first, then the axes on top
    vis = g("draw_posed_3d_box")(K, img=img.copy(), ob_in_cam=pose, bbox=bbox)
    out["run_demo_vis"] = g("draw_xyz_axis")(vis, ob_in_cam=pose, scale=0.1, K=K, thickness=3, transparency=0, is_input_rgb=True)

    # ---- symmetry tables (BOP models_info.json entries)
    half_z = np.diag([-1.0, -1.0, 1.0, 1.0])
    moved = np.eye(4)
    moved[:3, 3] = [10.0, -20.0, 30.0]  # millimetres in the file
    infos = {
        "none": {"diameter": 100.0},
        "discrete": {"symmetries_discrete": [half_z.reshape(-1).tolist(), moved.reshape(-1).tolist()]},
        "cont_x": {"symmetries_continuous": [{"axis": [1, 0, 0], "offset": [0, 0, 0]}]},
        "cont_y_offset": {"symmetries_continuous": [{"axis": [0, 1, 0], "offset": [0.01, 0.02, 0.03]}]},
        "cont_z": {"symmetries_continuous": [{"axis": [0, 0, 1], "offset": [0, 0, 0]}]},
        "both": {"symmetries_discrete": [half_z.reshape(-1).tolist()], "symmetries_continuous": [{"axis": [0, 0, 1], "offset": [0, 0, 0]}]},
    }
    for name, info in infos.items():
        out[f"symmetry_tfs.{name}"] = g("symmetry_tfs_from_info")(info)
    out["symmetry_tfs.cont_z_60"] = g("symmetry_tfs_from_info")(infos["cont_z"], rot_angle_discrete=60)

    # ---- result dictionaries of the dataset drivers (run_linemod.py:150-153: NestDict of 4x4 poses -> yaml)
    res = g("NestDict")()
    res[1][6]["000000"][6] = _pose(7, [0, 0, 1.0])
    res[1][6]["000001"][6] = _pose(8, [0, 0, 1.1])
    res["meta"]["count"] = np.int64(2)
    res["meta"]["score"] = np.float64(0.25)
    res["meta"]["name"] = np.str_("lm")
    res["meta"]["list"] = [np.arange(3), {"a": np.float64(1.5)}]
    res["meta"]["plain"] = {"x": 1, "y": "text"}
    dumped = g("make_yaml_dumpable")(res)
    out["make_yaml_dumpable.yaml"] = np.array(yaml.safe_dump(dumped, sort_keys=True))
    out["make_yaml_dumpable.array"] = np.array(g("make_yaml_dumpable")(np.arange(6).reshape(2, 3)))

    # ---- depth back-projection (the drop-in keeps its own copy for callers of Utils.depth2xyzmap)
    depth = rng.uniform(0.0, 2.0, size=(24, 32))
    depth[rng.uniform(size=depth.shape) < 0.2] = 0.0005
    Ks = np.array([[60.0, 0.0, 16.0], [0.0, 62.0, 12.0], [0.0, 0.0, 1.0]])
    out["depth2xyzmap"] = g("depth2xyzmap")(depth, Ks)
    out["depth2xyzmap.uvs"] = g("depth2xyzmap")(depth, Ks, uvs=np.array([[3, 4], [10.4, 7.6], [31, 23]]))
    d_t = torch.as_tensor(np.stack([depth, depth[::-1].copy()]), dtype=torch.float32)
    K_t = torch.as_tensor(np.stack([Ks, Ks * [[1.1], [0.9], [1.0]]]), dtype=torch.float32)
    out["depth2xyzmap_batch.inf"] = g("depth2xyzmap_batch")(d_t, K_t, zfar=np.inf).numpy()
    out["depth2xyzmap_batch.zfar"] = g("depth2xyzmap_batch")(d_t, K_t, zfar=1.5).numpy()

    # ---- model diameter as reset_object asks for it (estimater.py:54: n_sample = 10000 >= the vertex count)
    from foundationpose_b200 import synth

    verts = synth.make_mesh(3, tex_size=16).vertices
    out["compute_mesh_diameter"] = np.float64(g("compute_mesh_diameter")(model_pts=verts, n_sample=10000))

    # ---- set_seed: the three generators the reference seeds
    g("set_seed")(123)
    out["set_seed.numpy"] = np.random.rand(3)
    out["set_seed.random"] = np.array([random.random() for _ in range(3)])
    out["set_seed.torch"] = torch.rand(3).numpy()
    out["set_seed.cudnn"] = np.array([bool(torch.backends.cudnn.deterministic), bool(torch.backends.cudnn.benchmark)])
    return out
