"""The call sequence of the reference's run_demo.py (:26-79) written against the drop-in module names — used by
tests/test_dropin_gpu.py when the reference's own file is not available on the box.

    PYTHONPATH=foundationpose_b200/dropin:. python examples/run_demo_dropin.py --mesh_file <obj> --test_scene_dir <dir>
"""
from estimater import *  # noqa: F401,F403
from datareader import *  # noqa: F401,F403
import argparse

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh_file", required=True)
    ap.add_argument("--test_scene_dir", required=True)
    ap.add_argument("--est_refine_iter", type=int, default=5)
    ap.add_argument("--track_refine_iter", type=int, default=2)
    ap.add_argument("--debug", type=int, default=0)
    ap.add_argument("--debug_dir", default="debug")
    a = ap.parse_args()
    set_logging_format()
    set_seed(0)
    mesh = trimesh.load(a.mesh_file)
    os.makedirs(f"{a.debug_dir}/track_vis", exist_ok=True)
    os.makedirs(f"{a.debug_dir}/ob_in_cam", exist_ok=True)
    to_origin, extents = trimesh.bounds.oriented_bounds(mesh)
    bbox = np.stack([-extents / 2, extents / 2], axis=0).reshape(2, 3)
    est = FoundationPose(model_pts=mesh.vertices, model_normals=mesh.vertex_normals, mesh=mesh, scorer=ScorePredictor(),
                         refiner=PoseRefinePredictor(), debug_dir=a.debug_dir, debug=a.debug, glctx=dr.RasterizeCudaContext())
    reader = YcbineoatReader(video_dir=a.test_scene_dir, shorter_side=None, zfar=np.inf)
    for i in range(len(reader.color_files)):
        color, depth = reader.get_color(i), reader.get_depth(i)
        if i == 0:
            pose = est.register(K=reader.K, rgb=color, depth=depth, ob_mask=reader.get_mask(0).astype(bool), iteration=a.est_refine_iter)
        else:
            pose = est.track_one(rgb=color, depth=depth, K=reader.K, iteration=a.track_refine_iter)
        np.savetxt(f"{a.debug_dir}/ob_in_cam/{reader.id_strs[i]}.txt", pose.reshape(4, 4))
        if a.debug >= 1:
            center_pose = pose @ np.linalg.inv(to_origin)
            vis = draw_posed_3d_box(reader.K, img=color, ob_in_cam=center_pose, bbox=bbox)
            vis = draw_xyz_axis(color, ob_in_cam=center_pose, scale=0.1, K=reader.K, thickness=3, transparency=0, is_input_rgb=True)
            if a.debug >= 2:
                imageio.imwrite(f"{a.debug_dir}/track_vis/{reader.id_strs[i]}.png", vis)
