"""CPU: the drop-in module tree (foundationpose_b200/dropin) resolves every name the reference's UNMODIFIED
run_demo.py uses, the trimesh / imageio stand-ins round-trip the demo-scene files, and the reader parses them."""
import ast
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "foundationpose_b200", "dropin")
REF_DEMO = "/root/reference/run_demo.py"


def _env():
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([DROPIN, ROOT, env.get("PYTHONPATH", "")])
    return env


@pytest.mark.skipif(not os.path.exists(REF_DEMO), reason="reference tree not present (GPU box)")
def test_every_name_run_demo_uses_resolves():
    """Static check against the reference's own driver: all unqualified names and first-level attributes
    (`trimesh.load`, `dr.RasterizeCudaContext`, `np.stack`, ...) exist after its two star-imports."""
    import builtins

    tree = ast.parse(open(REF_DEMO).read())
    assigned, used, attrs = set(), set(), set()
    for node in ast.walk(tree):
        if isinstance(node, ast.Name):
            (assigned if isinstance(node.ctx, ast.Store) else used).add(node.id)
        elif isinstance(node, ast.Attribute) and isinstance(node.value, ast.Name):
            attrs.add((node.value.id, node.attr))
    need = sorted(n for n in used - assigned - set(dir(builtins)) if n != "__file__")
    assert {"trimesh", "dr", "np", "cv2", "imageio", "logging", "set_seed", "YcbineoatReader", "FoundationPose"} <= set(need)
    mod_attrs = sorted((m, a) for (m, a) in attrs if m in need and m not in ("args", "parser", "o3d"))
    code = ("from estimater import *\nfrom datareader import *\nimport argparse\n"
            f"missing = [n for n in {need!r} if n not in globals()]\n"
            f"missing += [f'{{m}}.{{a}}' for (m, a) in {mod_attrs!r} if m in globals() and not hasattr(globals()[m], a)]\n"
            "missing += [] if hasattr(trimesh.bounds, 'oriented_bounds') else ['trimesh.bounds.oriented_bounds']\n"
            "print('MISSING', missing)\n")
    out = subprocess.run([sys.executable, "-c", code], env=_env(), capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "MISSING []" in out.stdout, out.stdout[-2000:]


def test_demo_scene_round_trip(tmp_path):
    code = f"""
import numpy as np
from estimater import *
from datareader import *
from foundationpose_b200 import synth
mesh0, gt = synth.write_demo_scene({str(tmp_path)!r}, n_frames=2, subdivisions=2)
mesh = trimesh.load({str(tmp_path)!r} + '/mesh/textured_simple.obj')
assert np.abs(mesh.vertices[mesh.faces] - mesh0.vertices[mesh0.faces]).max() < 1e-6
assert np.abs(mesh.visual.uv[mesh.faces] - mesh0.visual.uv[mesh0.faces]).max() < 1e-6
assert np.abs(mesh.vertex_normals[mesh.faces] - mesh0.vertex_normals[mesh0.faces]).max() < 1e-6
img = np.asarray(mesh.visual.material.image.convert('RGB'))
assert (img == mesh0.visual.image).all()
mt, mt0 = make_mesh_tensors(mesh), make_mesh_tensors(mesh0)
assert (mt['tex'] == mt0['tex']).all() and mt['uv'].shape == (len(mesh.vertices), 2)
to_origin, extents = trimesh.bounds.oriented_bounds(mesh)
assert np.allclose(sorted(extents), sorted(2 * synth.RADII), rtol=0.03), extents
assert np.allclose(to_origin[:3, :3] @ to_origin[:3, :3].T, np.eye(3), atol=1e-9)
reader = YcbineoatReader(video_dir={str(tmp_path)!r}, shorter_side=None, zfar=np.inf)
assert len(reader.color_files) == 2 and reader.id_strs == ['000000', '000001'] and reader.K.shape == (3, 3)
color, depth, mask = reader.get_color(0), reader.get_depth(0), reader.get_mask(0).astype(bool)
rgb0, depth0, mask0 = synth.make_scene(mesh0.visual.image, gt[0], seed=1)
assert color.dtype == np.uint8 and (color == rgb0).all()
assert np.abs(depth - depth0).max() <= 0.00051 and (mask == mask0).all()
assert np.allclose(reader.get_gt_pose(1), gt[1])
vis = draw_posed_3d_box(reader.K, img=color.copy(), ob_in_cam=gt[0], bbox=np.stack([-extents / 2, extents / 2]))
vis = draw_xyz_axis(vis, ob_in_cam=gt[0], scale=0.1, K=reader.K, thickness=3, transparency=0, is_input_rgb=True)
assert vis.shape == color.shape and (vis != color).any()
xyz = depth2xyzmap(depth, reader.K)
assert xyz.shape == (480, 640, 3) and abs(xyz[240, 320, 2] - depth[240, 320]) < 1e-6
print('OK')
"""
    out = subprocess.run([sys.executable, "-c", code], env=_env(), capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "OK" in out.stdout, (out.stdout + out.stderr)[-3000:]
