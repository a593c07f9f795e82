"""Generate tests/golden/nets_golden.npz by running the REFERENCE's own, unmodified network classes
(/root/reference/learning/models/*.py) on seeded inputs.  Runs only in the build container (the
reference tree is not present on the GPU box); the fixture it writes is committed.

    python tools/make_golden.py

What is pinned:
  * foundationpose_b200.weights.random_state_dict() loads *strictly* into RefineNet /
    ScoreNetMultiPair (key layout + shapes are the reference's);
  * the fp32 CPU outputs of the reference modules for seeded crops -> oracle/nets.py must match.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def import_reference_models():
    # `from Utils import *` at the top of the model files needs heavy third-party deps; the model
    # code itself uses nothing from it, so an empty stand-in module is enough.
    sys.modules.setdefault("Utils", types.ModuleType("Utils"))
    sys.path.insert(0, os.path.join(REF, "learning", "models"))
    import refine_network  # noqa
    import score_network  # noqa

    return refine_network.RefineNet, score_network.ScoreNetMultiPair


class Cfg(dict):
    def __getattr__(self, k):
        return self[k]


def seeded_crops(n, seed):
    g = torch.Generator().manual_seed(seed)
    A = torch.rand(n, 6, 160, 160, generator=g)
    B = torch.rand(n, 6, 160, 160, generator=g)
    # xyz channels roughly in [-1, 1] with a zeroed background, like the real normalised crops
    for T in (A, B):
        T[:, 3:] = (T[:, 3:] - 0.5) * 2
        T[:, 3:, :30] = 0
    return A, B


def main():
    from foundationpose_b200.weights import random_state_dict

    RefineNet, ScoreNet = import_reference_models()
    cfg = Cfg(use_BN=True, rot_rep="axis_angle")
    out = {}
    torch.set_num_threads(os.cpu_count())

    sd_r = random_state_dict("refine", seed=0)
    m = RefineNet(cfg=cfg, c_in=6).eval()
    missing = m.load_state_dict(sd_r, strict=True)
    print("refine strict load:", missing)
    A, B = seeded_crops(2, 123)
    with torch.inference_mode():
        o = m(A, B)
    out["refine_trans"] = o["trans"].numpy()
    out["refine_rot"] = o["rot"].numpy()

    sd_s = random_state_dict("score", seed=0)
    ms = ScoreNet(cfg=cfg, c_in=6).eval()
    print("score strict load:", ms.load_state_dict(sd_s, strict=True))
    A, B = seeded_crops(3, 456)
    with torch.inference_mode():
        feats = ms.extract_feat(A, B)
        logits = ms(A, B, L=3)["score_logit"]
    out["score_feats"] = feats.numpy()
    out["score_logits"] = logits.numpy()
    # checksums of the seeded weights so a torch-version drift of the generator is detected
    out["refine_wsum"] = np.array([float(sum(v.double().abs().sum() for k, v in sd_r.items() if v.dtype.is_floating_point))])
    out["score_wsum"] = np.array([float(sum(v.double().abs().sum() for k, v in sd_s.items() if v.dtype.is_floating_point))])
    path = os.path.join(ROOT, "tests", "golden", "nets_golden.npz")
    np.savez(path, **out)
    for k, v in out.items():
        print(k, v.shape, np.abs(v).max())
    print("wrote", path)


if __name__ == "__main__":
    main()
