"""Golden vectors for the predictors' configuration defaults, produced by the REFERENCE's own statements.

`PoseRefinePredictor.__init__` (learning/training/predict_pose_refine.py:92-146) and `ScorePredictor.__init__`
(learning/training/predict_score.py:117-158) load `weights/<run>/config.yml` and then patch missing keys with
backward-compatibility defaults — a run of `if '<key>' not in self.cfg: ...` statements in the middle of constructors that
cannot be executed here (OmegaConf, datasets, CUDA models).  This script extracts exactly those `if` statements from the
two constructors with `ast` (every top-level `If` of `__init__` whose test mentions `self.cfg`), executes them on a plain
dict standing in for the OmegaConf node, and writes the resulting dictionaries for a set of partial configs to
tests/golden/predictor_defaults.json.  tests/test_host_logic_cpu.py holds `weights.load_reference_config` to them.

    python tools/make_golden_config.py      # needs /root/reference
"""
import ast
import json
import os
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("FPOSE_REFERENCE", "/root/reference")

CASES = {
    "empty": {},
    "released_like": {"use_BN": True, "c_in": 6, "normalize_xyz": True, "crop_ratio": 1.2, "zfar": "inf", "rot_rep": "axis_angle", "trans_rep": "tracknet"},
    "crop_ratio_null": {"crop_ratio": None, "zfar": "Inf", "c_in": 6},
    "scorer_like": {"crop_ratio": 1.1, "use_BN": True, "c_in": 6, "normalize_xyz": True},
    "old_run": {"use_normal": True, "n_view": 2, "zfar": 2.5},
}


def default_statements(path, cls):
    tree = ast.parse(open(path).read())
    init = next(f for c in tree.body if isinstance(c, ast.ClassDef) and c.name == cls for f in c.body
                if isinstance(f, ast.FunctionDef) and f.name == "__init__")
    ifs = [n for n in init.body if isinstance(n, ast.If) and "self.cfg" in ast.unparse(n.test)]
    return compile(ast.fix_missing_locations(ast.Module(body=ifs, type_ignores=[])), f"{path}:{cls}.__init__ defaults", "exec"), len(ifs)


def main():
    out = {}
    for kind, rel, cls in (("refine", "learning/training/predict_pose_refine.py", "PoseRefinePredictor"),
                           ("score", "learning/training/predict_score.py", "ScorePredictor")):
        code, n = default_statements(os.path.join(REF, rel), cls)
        print(f"{cls}: {n} default statements")
        for name, cfg in CASES.items():
            self = types.SimpleNamespace(cfg=dict(cfg))
            exec(code, {"self": self, "np": np})
            out[f"{kind}.{name}"] = {"in": cfg, "out": {k: ("inf" if isinstance(v, float) and np.isinf(v) else v) for k, v in self.cfg.items()}}
    dst = os.path.join(ROOT, "tests", "golden", "predictor_defaults.json")
    with open(dst, "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print(f"wrote {dst}: {len(out)} cases")


if __name__ == "__main__":
    main()
