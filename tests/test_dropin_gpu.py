"""GPU: the reference's own run_demo.py, UNMODIFIED, on top of foundationpose_b200/dropin and a synthetic scene in
the reference's demo-data layout.  The script is staged by __graft_entry__.build() into oracle/_ref/ (git-ignored,
never committed) because the GPU box has no /root/reference; without it the tests are skipped."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "foundationpose_b200", "dropin")


def _driver():
    for cand in ("/root/reference/run_demo.py", os.path.join(ROOT, "oracle", "_ref", "run_demo.py")):
        if os.path.exists(cand):
            return cand, True
    pytest.skip("reference driver not staged (run __graft_entry__.build() where /root/reference exists)")


@pytest.mark.parametrize("debug", [0, 2])
def test_run_demo_unmodified(tmp_path, debug):
    from foundationpose_b200 import synth

    scene = str(tmp_path / "demo_data" / "synth0")
    mesh, gt = synth.write_demo_scene(scene, n_frames=4, subdivisions=3)
    script, is_reference = _driver()
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([DROPIN, ROOT, env.get("PYTHONPATH", "")])
    env["QT_QPA_PLATFORM"] = "offscreen"
    dbg = str(tmp_path / "debug")
    cmd = [sys.executable, script, "--mesh_file", scene + "/mesh/textured_simple.obj", "--test_scene_dir", scene,
           "--est_refine_iter", "5", "--track_refine_iter", "2", "--debug", str(debug), "--debug_dir", dbg]
    if debug >= 1:
        # cv2.imshow needs a display: the headless OpenCV build raises.  A sitecustomize on PYTHONPATH turns imshow /
        # waitKey into no-ops for this process only; the driver itself stays byte-identical.
        sc = tmp_path / "site"
        sc.mkdir()
        (sc / "sitecustomize.py").write_text("import cv2\ncv2.imshow = lambda *a, **k: None\ncv2.waitKey = lambda *a, **k: -1\n")
        env["PYTHONPATH"] = str(sc) + os.pathsep + env["PYTHONPATH"]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert out.returncode == 0, (out.stdout + out.stderr)[-4000:]
    print(f"driver: {script} (reference's own file: {is_reference})")
    poses = []
    for i in range(4):
        f = os.path.join(dbg, "ob_in_cam", f"{i:06d}.txt")
        assert os.path.exists(f), f"{f} missing\n" + (out.stdout + out.stderr)[-2000:]
        p = np.loadtxt(f).reshape(4, 4)
        assert np.isfinite(p).all() and abs(np.linalg.det(p[:3, :3]) - 1) < 1e-3
        poses.append(p)
    # random-init weights: no accuracy claim, but the seeded stand-in moves a pose by millimetres per pass, so the
    # tracked object stays near where the first frame's mask put it
    assert all(0.4 < p[2, 3] < 0.8 for p in poses), [p[:3, 3] for p in poses]
    if debug >= 2:
        assert os.path.exists(os.path.join(dbg, "track_vis", "000003.png"))


def _dataset_driver(name):
    for cand in ("/root/reference/" + name, os.path.join(ROOT, "oracle", "_ref", name)):
        if os.path.exists(cand):
            return cand
    return None


def _load_result(path):
    import yaml

    with open(path) as fh:
        return yaml.safe_load(fh)


def test_run_linemod_unmodified(tmp_path):
    """the reference's LINEMOD driver (run_linemod.py: 13 objects, `reset_object` per object, one
    `register` per frame, results to linemod_res.yml), UNMODIFIED, over a synthetic dataset in its directory layout.
    Every pose it writes must be the pose the native API returns for the same reader inputs."""
    from foundationpose_b200 import synth

    script = _dataset_driver("run_linemod.py")
    if script is None:
        pytest.skip("reference driver not staged (run __graft_entry__.build() where /root/reference exists)")
    root = str(tmp_path / "LINEMOD")
    gts = synth.write_bop_dataset(root, "lm", n_frames=1)
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([DROPIN, ROOT, env.get("PYTHONPATH", "")])
    dbg = str(tmp_path / "debug")
    out = subprocess.run([sys.executable, script, "--linemod_dir", root, "--debug_dir", dbg], env=env, capture_output=True, text=True,
                         timeout=900, cwd=str(tmp_path))
    assert out.returncode == 0, (out.stdout + out.stderr)[-4000:]
    res = _load_result(os.path.join(dbg, "linemod_res.yml"))
    assert sorted(res.keys()) == [1, 2, 4, 5, 6, 8, 9, 10, 11, 12, 13, 14, 15]
    for (scene, frame, ob), gt in gts.items():
        p = np.array(res[scene][frame][ob])
        assert p.shape == (4, 4) and np.isfinite(p).all() and abs(np.linalg.det(p[:3, :3]) - 1) < 1e-3
        assert np.linalg.norm(p[:3, 3] - gt[:3, 3]) < 0.08, (scene, p[:3, 3], gt[:3, 3])  # stand-in weights: stays near the mask
    # the same call through the native API, in this process
    sys.path[:0] = [DROPIN, ROOT]
    try:
        import datareader
        from foundationpose_b200.estimater import FoundationPose

        trimesh = datareader.trimesh  # the real package or the stand-in, whichever `Utils` resolved

        box = trimesh.primitives.Box(extents=np.ones(3), transform=np.eye(4)).to_mesh()
        est = FoundationPose(model_pts=box.vertices.copy(), model_normals=box.vertex_normals.copy(), mesh=box, debug_dir=str(tmp_path / "dbg2"))
        for ob in (6, 9):
            r = datareader.LinemodReader(f"{root}/lm_test_all/test/{ob:06d}", split=None)
            mesh = r.get_gt_mesh(ob)
            # the driver's sequence (run_linemod.py:100-112): one estimator, `reset_object` per object.  Like the
            # reference's, `reset_object` keeps the rotation grid built at construction (estimater.py:40-41 vs :43-85):
            # the per-object symmetries do not thin the 252 start poses on this route
            est.reset_object(model_pts=mesh.vertices.copy(), model_normals=mesh.vertex_normals.copy(), symmetry_tfs=r.symmetry_tfs[ob], mesh=mesh)
            pose = est.register(K=r.K, rgb=r.get_color(0), depth=r.get_depth(0), ob_mask=r.get_mask(0, ob), ob_id=ob)
            np.testing.assert_allclose(np.array(res[ob]["000000"][ob]), pose, atol=1e-5)
            assert len(est.rot_grid) == 252
        # constructed WITH the symmetry (obj 6: half turn about z in models_info.json) the grid is clustered under it
        r = datareader.LinemodReader(f"{root}/lm_test_all/test/000006", split=None)
        mesh = r.get_gt_mesh(6)
        sym = FoundationPose(model_pts=mesh.vertices.copy(), model_normals=mesh.vertex_normals.copy(), symmetry_tfs=r.symmetry_tfs[6], mesh=mesh,
                             debug_dir=str(tmp_path / "dbg3"))
        assert len(sym.rot_grid) < 252
    finally:
        del sys.path[:2]


def test_run_ycb_video_unmodified(tmp_path):
    """The YCB-Video driver (run_ycb_video.py: 21 objects x the scenes that contain them, key frames only,
    zfar = 1.5), UNMODIFIED, over a synthetic dataset of three one-object scenes."""
    from foundationpose_b200 import synth

    script = _dataset_driver("run_ycb_video.py")
    if script is None:
        pytest.skip("reference driver not staged (run __graft_entry__.build() where /root/reference exists)")
    root = str(tmp_path / "YCB_Video")
    gts = synth.write_bop_dataset(root, "ycbv", n_frames=2, scene_objects={48: 1, 49: 6, 50: 13})
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([DROPIN, ROOT, env.get("PYTHONPATH", "")])
    dbg = str(tmp_path / "debug")
    out = subprocess.run([sys.executable, script, "--ycbv_dir", root, "--debug_dir", dbg], env=env, capture_output=True, text=True,
                         timeout=900, cwd=str(tmp_path))
    assert out.returncode == 0, (out.stdout + out.stderr)[-4000:]
    res = _load_result(os.path.join(dbg, "ycbv_res.yml"))
    assert sorted(res.keys()) == [48, 49, 50]
    n = 0
    for (scene, frame, ob), gt in gts.items():
        p = np.array(res[scene][frame][ob])
        assert p.shape == (4, 4) and np.isfinite(p).all() and abs(np.linalg.det(p[:3, :3]) - 1) < 1e-3
        assert np.linalg.norm(p[:3, 3] - gt[:3, 3]) < 0.08
        n += 1
    assert n == 6


def test_linemod_over_replicas_matches_the_sequential_driver(tmp_path):
    """examples/run_linemod_replicas.py (frames of an object spread over the GPUs by ReplicaPool) writes the same
    linemod_res.yml as the reference's sequential driver loop does through one estimator."""
    import torch

    from foundationpose_b200 import synth

    root = str(tmp_path / "LINEMOD")
    synth.write_bop_dataset(root, "lm", n_frames=2)
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([DROPIN, ROOT, env.get("PYTHONPATH", "")])
    out_dirs = []
    for tag, script, extra in (("pool", os.path.join(ROOT, "examples", "run_linemod_replicas.py"), ["--gpus", str(min(2, torch.cuda.device_count()))]),
                               ("seq", _dataset_driver("run_linemod.py"), [])):
        if script is None:
            pytest.skip("reference driver not staged")
        dbg = str(tmp_path / ("debug_" + tag))
        out = subprocess.run([sys.executable, script, "--linemod_dir", root, "--debug_dir", dbg] + extra, env=env, capture_output=True,
                             text=True, timeout=900, cwd=str(tmp_path))
        assert out.returncode == 0, (out.stdout + out.stderr)[-4000:]
        out_dirs.append(dbg)
    a, b = (_load_result(os.path.join(d, "linemod_res.yml")) for d in out_dirs)
    assert sorted(a.keys()) == sorted(b.keys())
    for vid in a:
        for frame in a[vid]:
            for ob in a[vid][frame]:
                np.testing.assert_allclose(np.array(a[vid][frame][ob]), np.array(b[vid][frame][ob]), atol=1e-5)

