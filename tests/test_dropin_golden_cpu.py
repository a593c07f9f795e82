"""CPU: the helper functions of the drop-in `Utils` module (foundationpose_b200/dropin/Utils.py) against golden vectors
produced by the REFERENCE's own function bodies (tools/make_golden_shim.py: unmodified source extracted from
/root/reference/Utils.py with `ast`): to_homo, transform_pts (3-D and the predictors' 2-D window corners),
project_3d_to_2d, draw_xyz_axis and draw_posed_3d_box down to the anti-aliased pixel, symmetry_tfs_from_info,
NestDict + make_yaml_dumpable, depth2xyzmap(_batch), compute_mesh_diameter, set_seed."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "foundationpose_b200", "dropin")
GOLDEN = os.path.join(ROOT, "tests", "golden", "shim_golden.npz")


def test_dropin_utils_match_the_reference_function_bodies():
    code = f"""
import os, sys
import numpy as np
sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})
import shim_cases
import Utils                                    # the drop-in (first on PYTHONPATH)
assert os.path.dirname(os.path.abspath(Utils.__file__)) == {DROPIN!r}, Utils.__file__
got = shim_cases.collect(Utils)
want = dict(np.load({GOLDEN!r}))
bad = [k for k in want if k not in got]
for k in want:
    if k not in got:
        continue
    a, b = np.asarray(got[k]), np.asarray(want[k])
    if a.shape != b.shape:
        bad.append(f'{{k}}: shape {{a.shape}} vs {{b.shape}}')
    elif a.dtype.kind in 'US' or b.dtype.kind in 'US':
        if str(a) != str(b):
            bad.append(f'{{k}}: text differs')
    elif a.dtype.kind in 'ui' and b.dtype.kind in 'ui':
        if not np.array_equal(a, b):                                   # images, pixel coordinates: identical
            bad.append(f'{{k}}: {{int((a != b).sum())}} entries differ')
    elif not np.allclose(a.astype(np.float64), b.astype(np.float64), rtol=1e-12, atol=1e-12):
        bad.append(f'{{k}}: max |diff| {{np.abs(a.astype(np.float64) - b.astype(np.float64)).max():.3g}}')
print('ENTRIES', len(want))
print('DIFFERENCES', bad)
"""
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([DROPIN, ROOT, env.get("PYTHONPATH", "")])
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, (out.stdout + out.stderr)[-3000:]
    assert "DIFFERENCES []" in out.stdout, out.stdout[-3000:]
    assert "ENTRIES 32" in out.stdout, out.stdout[-500:]
