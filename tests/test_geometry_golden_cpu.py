"""CPU: the oracle's geometry restatements (oracle/geometry.py, oracle/raster.py) and the product's host-side
guess_translation against vectors produced by the REFERENCE'S OWN FUNCTION BODIES (tools/make_golden_geometry.py
extracts them from Utils.py / estimater.py / h5_dataset.py with `ast` and runs them on the CPU).

Bars: crop-window transforms bit-exact (the edges are rounded integers; the scale is one fp32 division);
back-projection, pose composition and xyz normalisation 1e-6 abs (same fp32 operations);
guess_translation 1e-12 (fp64 on both sides); raster pixel mapping 2e-3 px (fp32 chain vs the fp64 reference)."""
import os

import numpy as np
import torch

from foundationpose_b200 import hypotheses
from oracle import geometry, raster

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "geometry_golden.npz"))


def test_crop_window_matches_reference_compute_crop_window_tf_batch():
    win, tf = geometry.crop_window(G["cw_poses"], G["cw_K"], float(G["cw_diameter"]))
    ref = G["cw_tf"]
    # window edges recovered from the reference transform: tf = [[sx, 0, -left*sx], [0, sy, -top*sy], [0, 0, 1]]
    np.testing.assert_array_equal(tf[:, 0, 0], ref[:, 0, 0])
    np.testing.assert_array_equal(tf[:, 1, 1], ref[:, 1, 1])
    np.testing.assert_array_equal(np.round(-ref[:, 0, 2] / ref[:, 0, 0]), win["left"])
    np.testing.assert_array_equal(np.round(-ref[:, 1, 2] / ref[:, 1, 1]), win["top"])
    np.testing.assert_allclose(tf, ref, rtol=0, atol=2e-4)  # the offset column is a product; 1 ulp at |x| ~ 1e3
    # 'box_3d' windows are square up to the independent rounding of the four edges
    assert (np.abs((win["right"] - win["left"]) - (win["bottom"] - win["top"])) <= 1).all()


def test_depth2xyzmap_matches_reference():
    np.testing.assert_allclose(geometry.depth2xyzmap(G["dx_depth"], G["dx_K"]), G["dx_xyz"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(geometry.depth2xyzmap(G["dx_depth"], G["dx_K"], zfar=np.inf), G["dx_xyz_inf"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(geometry.depth2xyzmap(G["dx_depth"], G["dx_K"], zfar=1.5), G["dx_xyz_zfar15"], rtol=0, atol=1e-6)
    assert (G["dx_xyz"][G["dx_depth"] < 0.001] == 0).all()


def test_pose_composition_matches_reference_egocentric_delta():
    """pose_update = so3_exp_map (pytorch3d, unpinned) + egocentric_delta_pose_to_pose (pinned here): feed the
    reference's rotation deltas through the same composition."""
    A = torch.from_numpy(G["pu_A"])
    out = torch.eye(4)[None].repeat(len(A), 1, 1)
    out[:, :3, 3] = A[:, :3, 3] + torch.from_numpy(G["pu_trans_delta"])
    out[:, :3, :3] = torch.from_numpy(G["pu_rot_delta"]) @ A[:, :3, :3]
    np.testing.assert_allclose(out.numpy(), G["pu_B"], rtol=0, atol=1e-6)
    # and the oracle's full update with a zero rotation vector reproduces the translation part exactly
    zero = torch.zeros(len(A), 3)
    upd, td, _ = geometry.pose_update(A, torch.from_numpy(G["pu_trans_delta"]) / (0.19 / 2), zero, 0.19, 0.349)
    np.testing.assert_allclose(upd[:, :3, 3].numpy(), G["pu_B"][:, :3, 3], rtol=0, atol=1e-6)


def test_guess_translation_matches_reference_method():
    for i in range(4):
        depth = np.zeros_like(G["gt_depth"]) if G["gt_zero_depth_case"][i] else G["gt_depth"]
        t = hypotheses.guess_translation(depth, G["gt_masks"][i], G["gt_K"])
        np.testing.assert_allclose(t, G["gt_t"][i], rtol=0, atol=1e-12)
    assert np.abs(G["gt_t"][0]).sum() > 0 and (G["gt_t"][2] == 0).all() and (G["gt_t"][3] == 0).all()
    assert (G["gt_t"][0] != G["gt_t"][1]).any()  # odd vs even count: np.median averages the two middle values


def test_normalise_xyz_matches_reference_transform_depth_to_xyzmap():
    for src, dst in (("nx_xyzA", "nx_outA"), ("nx_xyzB", "nx_outB")):
        got = geometry.normalise_xyz(torch.from_numpy(G[src]), torch.from_numpy(G["nx_t"]), float(G["nx_diameter"]), 0.001)
        np.testing.assert_allclose(got.numpy(), G[dst], rtol=0, atol=1e-6)
    assert (G["nx_outA"][:, :, :3] == 0).all()  # z < 0.001 -> 0 after the shift, as in the reference


def test_raster_pixel_mapping_matches_reference_projection_and_bbox_crop():
    """Where a camera-frame point lands in the 160 x 160 crop: the reference's OpenGL projection matrix followed by the
    bbox2d clip-space crop (Utils.py:159-181) vs the direct K-projection + window scaling of oracle/raster.py (and of
    xform_vertex() in csrc/fp_crop.cu)."""
    S = 160
    umin, vmin, umax, vmax = [np.float32(x) for x in G["pj_bbox"]]
    rsx, rsy = np.float32(S) / (umax - umin), np.float32(S) / (vmax - vmin)
    X, Y, Z, iz, xi, yi = raster._project(np.eye(4), G["pj_pts_cam"].astype(np.float32), G["pj_K"], umin, vmin, rsx, rsy)
    np.testing.assert_allclose(xi / 256.0, G["pj_px"], rtol=0, atol=2e-3 + 1 / 512)
    np.testing.assert_allclose(yi / 256.0, S - G["pj_py_from_bottom"], rtol=0, atol=2e-3 + 1 / 512)


def test_scorer_normalise_xyz_matches_reference_triplet_transform():
    """TripletH5Dataset.transform_depth_to_xyzmap (h5_dataset.py:137-170): invalid = z < 0.1 for the scorer."""
    gotA = geometry.normalise_xyz(torch.from_numpy(G["ns_xyzA"]), torch.from_numpy(G["nx_t"]), float(G["nx_diameter"]), 0.1)
    gotB = geometry.normalise_xyz(torch.from_numpy(G["nx_xyzB"]), torch.from_numpy(G["nx_t"]), float(G["nx_diameter"]), 0.1)
    np.testing.assert_allclose(gotA.numpy(), G["ns_outA"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(gotB.numpy(), G["ns_outB"], rtol=0, atol=1e-6)
    assert (G["ns_outA"][:, :, 10:12] == 0).all() and (G["nx_outA"][:, :, 10:12] != 0).any()  # 0.05 m: scorer-only invalid


def test_depth_filters_match_reference_warp_kernel_bodies():
    """erode_depth_kernel / bilateral_filter_depth_kernel (Utils.py:304-384) executed thread by thread as plain Python
    (fp64 scalars; Warp and the oracle compute in fp32, hence 2e-6)."""
    er = geometry.erode_depth(G["df_depth"])
    np.testing.assert_allclose(er, G["df_eroded"], rtol=0, atol=0)  # erosion only selects values: exact
    bl = geometry.bilateral_filter_depth(G["df_eroded"])
    np.testing.assert_allclose(bl, G["df_bilateral"], rtol=0, atol=2e-6)
    assert (G["df_eroded"] == 0).sum() > (G["df_depth"] == 0).sum()  # the fixture really erodes something


def test_make_mesh_tensors_matches_reference():
    """Utils.py:104-130 vs the product's host-side estimater.make_mesh_tensors (pure numpy part; no GPU involved):
    v-flipped uv, RGB texture (the 1/255 scale is applied inside the shading kernel), vertex colours / 255."""
    from foundationpose_b200 import synth
    from foundationpose_b200.estimater import make_mesh_tensors

    m = synth.SimpleMesh(G["mt_vertices"], G["mt_faces"], G["mt_normals"], uv=G["mt_uv_in"], texture=G["mt_tex_in"])
    mt = make_mesh_tensors(m)
    np.testing.assert_array_equal(mt["pos"], G["mt_pos"])
    np.testing.assert_array_equal(mt["faces"], G["mt_faces_out"])
    np.testing.assert_array_equal(mt["normals"], G["mt_vnormals"])
    np.testing.assert_array_equal(mt["uv"], G["mt_uv"])
    np.testing.assert_allclose(mt["tex"].astype(np.float32) / 255.0, G["mt_tex"][0], rtol=0, atol=1e-7)
    mc = synth.SimpleMesh(G["mt_vertices"], G["mt_faces"], G["mt_normals"], vertex_colors=G["mt_vcolor_in"])
    np.testing.assert_allclose(make_mesh_tensors(mc)["vcolor"], G["mt_vertex_color"], rtol=0, atol=1e-7)
