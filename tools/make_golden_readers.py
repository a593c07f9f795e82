"""Golden vectors for the dataset readers of the drop-in tree, produced by the REFERENCE's own unmodified
`datareader.py` (:57-613).

The file starts with `from Utils import *`, and the reference's `Utils.py` cannot be imported here (open3d, trimesh,
nvdiffrast, ... are absent).  The reader classes themselves only need cv2 / imageio / trimesh / json / glob and the two
helpers `depth2xyzmap` and `symmetry_tfs_from_info`, so the reference file is loaded AS IS with the drop-in `Utils` on the
path (whose `depth2xyzmap` is itself pinned to the reference body, tests/test_dropin_golden_cpu.py), driven over the
synthetic trees of tests/reader_cases.py, and everything its classes return is written to
tests/golden/readers_golden.npz.  tests/test_readers_golden_cpu.py holds the drop-in readers
(foundationpose_b200/dropin/datareader.py, bop.py) to it.

    python tools/make_golden_readers.py            # needs /root/reference; writes the fixture and compares the drop-in
"""
import importlib.util
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "foundationpose_b200", "dropin"))
REF = os.environ.get("FPOSE_REFERENCE_DIR", "/root/reference")


def main():
    import reader_cases

    root = tempfile.mkdtemp(prefix="fpose_readers_")
    try:
        os.environ.update(reader_cases.build(root))
        spec = importlib.util.spec_from_file_location("reference_datareader", os.path.join(REF, "datareader.py"))
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)  # the reference's file, unmodified
        want = reader_cases.collect(ref, root)
        dst = os.path.join(ROOT, "tests", "golden", "readers_golden.npz")
        np.savez_compressed(dst, **want)
        print(f"wrote {dst}: {len(want)} entries, {os.path.getsize(dst) / 1024:.0f} KiB")
        import datareader as shim  # the drop-in, same process, same trees

        got = reader_cases.collect(shim, root)
        bad = reader_cases.compare(got, want)
        print("drop-in vs reference:", "identical" if not bad else f"{len(bad)} differences")
        for line in bad[:40]:
            print("  ", line)
    finally:
        shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    main()
