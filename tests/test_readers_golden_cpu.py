"""CPU: the drop-in dataset readers (foundationpose_b200/dropin/datareader.py, bop.py) against golden vectors produced
by the REFERENCE's own unmodified `datareader.py` (tools/make_golden_readers.py): the demo / YCBInEOAT reader
(datareader.py:57-152) and every BOP reader (:155-613: LINEMOD, LINEMOD-O, YCB-Video, T-LESS, HB, ITODD, IC-BIN, TUD-L,
the path dispatch and the test-split lookup) over the synthetic trees of tests/reader_cases.py — 1 090 recorded values:
ids, intrinsics, images and masks (checksums), metric depth / xyz maps, ground-truth poses incl. the multi-instance IoU
selection, symmetry tables, model files, key frames."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "foundationpose_b200", "dropin")
GOLDEN = os.path.join(ROOT, "tests", "golden", "readers_golden.npz")


def test_dropin_readers_match_the_reference_readers(tmp_path):
    code = f"""
import os, sys
import numpy as np
sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})
import reader_cases
root = {str(tmp_path)!r}
os.environ.update(reader_cases.build(root))    # BOP_DIR is read when the reader module is imported
import datareader                               # the drop-in (first on PYTHONPATH)
assert os.path.dirname(os.path.abspath(datareader.__file__)) == {DROPIN!r}, datareader.__file__
got = reader_cases.collect(datareader, root)
want = dict(np.load({GOLDEN!r}))
bad = reader_cases.compare(got, want)
print('ENTRIES', len(want))
print('DIFFERENCES', bad[:20])
"""
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([DROPIN, ROOT, env.get("PYTHONPATH", "")])
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, (out.stdout + out.stderr)[-3000:]
    assert "DIFFERENCES []" in out.stdout, out.stdout[-3000:]
    assert "ENTRIES 1090" in out.stdout, out.stdout[-500:]
